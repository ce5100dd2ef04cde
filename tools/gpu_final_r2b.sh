#!/bin/bash
# round-2 evidence on ONE GPU, second half: bench lines + ncu captures exported to CSV on the box (the .ncu-rep files of
# four captures exceed what gpurun copies back)
mkdir -p gpurun_out
P=gpurun_out/r2f
timeout 600 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
for w in h2o_16 c20h42; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 5 > ${P}_bench_$w.json 2> ${P}_bench_$w.err; done
MYQC_OUTPUT_MODE=compose timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > ${P}_bench_h2o_64_compose.json 2> ${P}_bench_h2o_64_compose.err; echo "compose bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_h2o64_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > ${P}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
cap() { tag=$1; regex=$2; shift; shift
  env "$@" timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$regex" -c 10 -o /tmp/cap_$tag -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e $EXTRA > ${P}_ncu_full_$tag.log 2>&1; echo "ncu full $tag rc=$?"
  ncu -i /tmp/cap_$tag.ncu-rep --page raw --csv > ${P}_${tag}_raw.csv 2>/dev/null
  rm -f /tmp/cap_$tag.ncu-rep
}
EXTRA="" cap h2o64 'eri_class|fill_zero' MYQC_X=0
EXTRA="" cap h2o64_compose 'eri_class|compose' MYQC_OUTPUT_MODE=compose
for w in h2o_16 c20h42; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_${w}_launches.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > ${P}_ncu_launches_$w.log 2>&1; echo "ncu launches $w rc=$?"
EXTRA="--workload $w" cap $w 'eri_class|fill_zero' MYQC_X=0
done
du -sh gpurun_out
