#!/bin/bash
# new defaults (memset fill, 256-byte transfer chunks, streaming host zeros): parity + the three bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2x_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 5 > gpurun_out/r2x_bench_h2o_64.json 2> gpurun_out/r2x_bench_h2o_64.err; echo "bench rc=$?"
for w in h2o_16 c20h42; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 3 > gpurun_out/r2x_bench_$w.json 2> gpurun_out/r2x_bench_$w.err; done
for w in h2o_64 h2o_16 c20h42; do python - gpurun_out/r2x_bench_$w.json $w <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "e2e %.1f ms"%d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done
