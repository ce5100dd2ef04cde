#!/bin/bash
# SM partition by shared-memory exhaustion: zero fill on a few SMs of its own next to the class kernels (staging mode),
# then the partition mode end to end (fill || class, scatter pass)
mkdir -p gpurun_out
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_${w}_$tag.json 2> gpurun_out/r2q_bench_${w}_$tag.err
  python - gpurun_out/r2q_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for n in 16 32 48 64; do
run fillonly_$n h2o_64 MYQC_EXP_CORUN=2 MYQC_FILL_SMS=$n
run corun_$n h2o_64 MYQC_EXP_CORUN=1 MYQC_FILL_SMS=$n
done
MYQC_OUTPUT_MODE=partition timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "CO2 or h2o_8 or shard or c4h10" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2q_pytest.log
for n in 24 32 40 48; do
run part_$n h2o_64 MYQC_OUTPUT_MODE=partition MYQC_FILL_SMS=$n
done
run part_32 h2o_16 MYQC_OUTPUT_MODE=partition
run part_32 c20h42 MYQC_OUTPUT_MODE=partition
