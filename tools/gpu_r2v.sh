#!/bin/bash
# copy-engine zero fill: (1) stand-alone probe of memset / D2D-copy fills next to FP64 and scatter kernels,
# (2) the library with MYQC_FILL_ENGINE x MYQC_VSHARDS (fill of piece k+1 under the class kernels of piece k)
mkdir -p gpurun_out
timeout 300 tools/probes/ce_fill_probe 16 > gpurun_out/r2v_ce_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2v_ce_probe.txt
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench_${w}_$tag.json 2> gpurun_out/r2v_bench_${w}_$tag.err
  python - gpurun_out/r2v_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run kernel_v1 h2o_64 MYQC_X=0
run memset_v1 h2o_64 MYQC_FILL_ENGINE=memset
run memset_v2 h2o_64 MYQC_FILL_ENGINE=memset MYQC_VSHARDS=2
run memset_v4 h2o_64 MYQC_FILL_ENGINE=memset MYQC_VSHARDS=4
run memset_v8 h2o_64 MYQC_FILL_ENGINE=memset MYQC_VSHARDS=8
run copy_s1_v1 h2o_64 MYQC_FILL_ENGINE=copy
run copy_s4_v1 h2o_64 MYQC_FILL_ENGINE=copy MYQC_FILL_STREAMS=4
run copy_s4_v2 h2o_64 MYQC_FILL_ENGINE=copy MYQC_FILL_STREAMS=4 MYQC_VSHARDS=2
run copy_s4_v4 h2o_64 MYQC_FILL_ENGINE=copy MYQC_FILL_STREAMS=4 MYQC_VSHARDS=4
run copy_s4_v8 h2o_64 MYQC_FILL_ENGINE=copy MYQC_FILL_STREAMS=4 MYQC_VSHARDS=8
run copy_s4_v16 h2o_64 MYQC_FILL_ENGINE=copy MYQC_FILL_STREAMS=4 MYQC_VSHARDS=16
run kernel_v4 h2o_64 MYQC_VSHARDS=4
