#!/bin/bash
# usage: tools/exp_env.sh tag "ENV=.. ENV=.." ...   -- one short bench run per environment variant
tag=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$envs" gpurun_out/${tag}_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1] or "(default)", "| step %.2f ms |"%d["ms_per_step"], " ".join("%s %.2f"%(k["kernel"].replace("eri_class",""),k["ms"]) for k in d["kernels"]), "| sum %.2f"%sum(k["ms"] for k in d["kernels"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
