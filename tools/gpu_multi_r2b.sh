#!/bin/bash
# multi-GPU bench lines only: usage gpu_multi_r2b.sh N tag [workloads...]
N=$1; tag=$2; shift; shift
mkdir -p gpurun_out
P=gpurun_out/r2m${N}
for w in "$@"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-cpu-baseline > ${P}_bench_${w}_$tag.json 2> ${P}_bench_${w}_$tag.err
python - ${P}_bench_${w}_$tag.json "$w $tag N=$N" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "value %.4g"%d["value"], "per rank", ["%.3f"%r["ms"] for r in d["per_rank"]], "e2e ms", (d["e2e"] or {}).get("ms_per_step"))
    print("   rank0 kernels:", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "serial", d["roofline"]["serialised_launch_sum_ms"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done
