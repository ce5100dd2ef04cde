#!/bin/bash
# why does the primitive split slow the (SP SP|SP SP) slices down?  ncu of one slice, split on / off, C20H42
mkdir -p gpurun_out
for m in split nosplit; do
  if [ $m = nosplit ]; then export MYQC_SPLIT_MAXLG=0; else unset MYQC_SPLIT_MAXLG; fi
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:eri_class -s 6 -c 1 -o gpurun_out/r2s_c20h42_22_$m -f python bench.py --workload c20h42 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2s_ncu_$m.log 2>&1; echo "ncu $m rc=$?"
done
