#!/bin/bash
# multi-GPU evidence: usage gpu_multi_r2.sh N   (one box with N GPUs)
N=$1
mkdir -p gpurun_out
P=gpurun_out/r2m${N}
nvidia-smi -L | head -n 8
if [ "$N" = "2" ]; then
  # the in-process multi-device path (myqc_eri_packed / myqc_eri_dense / int2e with ngpu > 1) on real devices
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "all_devices or slabs or shards or drop_in or executable" > ${P}_pytest_multidev.log 2>&1; echo "pytest multi-device rc=$?"; tail -n 3 ${P}_pytest_multidev.log
  timeout 300 python - > ${P}_inprocess.log 2>&1 <<'PY'
import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import myqc_b200 as Q
from bench import build_system
s, zm = build_system("h2o_16")
a = Q.eri_packed(s, ngpu=1)
t0 = time.perf_counter(); b = Q.eri_packed(s, ngpu=2); t1 = time.perf_counter()
print("h2o_16 in-process ngpu=2 vs ngpu=1: max abs diff %.3e, zero pattern equal %s, %.1f ms" % (np.abs(a - b).max(), np.array_equal(a == 0, b == 0), 1e3 * (t1 - t0)))
PY
  cat ${P}_inprocess.log | tail -n 2
fi
run() { w=$1; tag=$2; shift; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-cpu-baseline > ${P}_bench_${w}_$tag.json 2> ${P}_bench_${w}_$tag.err
  python - ${P}_bench_${w}_$tag.json "$w $tag N=$N" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "value %.4g"%d["value"], "per rank", ["%.3f"%r["ms"] for r in d["per_rank"]], "e2e ms", (d["e2e"] or {}).get("ms_per_step"))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run h2o_64 scatter MYQC_X=0
run c20h42 scatter MYQC_X=0
if [ "$N" = "8" ]; then run h2o_64 compose MYQC_OUTPUT_MODE=compose; fi
