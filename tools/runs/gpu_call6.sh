#!/bin/bash
# experiment: deferred near/mid Boys pass for the light classes (variant libraries from tools/build_variants.sh)
mkdir -p gpurun_out
for v in base defer1 defer2; do
  if [ "$v" = base ]; then unset MYQC_LIB; else export MYQC_LIB=$PWD/myqc_b200/csrc/variants/libmyqc_eri_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c6_var_$v.json 2> gpurun_out/c6_var_$v.err
  python - "$v" gpurun_out/c6_var_$v.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "| step %.2f ms |"%d["ms_per_step"], " ".join("%s %.2f"%(k["kernel"].replace("eri_class",""),k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
export MYQC_LIB=$PWD/myqc_b200/csrc/variants/libmyqc_eri_defer2.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/c6_pytest_defer2.log 2>&1; echo "parity(defer2) rc=$?"; tail -n 3 gpurun_out/c6_pytest_defer2.log
for w in h2o_16 c20h42; do timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('defer2', d['config']['workload'], 'ms/step %.4f'%d['ms_per_step'], ' '.join('%s %.3f'%(k['kernel'].replace('eri_class',''),k['ms']) for k in d['kernels']))"; done
