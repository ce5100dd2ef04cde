#!/bin/bash
# (1) fill kernel variants against cudaMemsetAsync, (2) sweep of the sparse device->host route
mkdir -p gpurun_out
timeout 200 tools/probes/ce_fill_probe 16 variants > gpurun_out/r2w_fill_variants.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2w_fill_variants.txt
timeout 600 python tools/exp_e2e_sweep.py h2o_64 > gpurun_out/r2w_e2e_sweep.txt 2> gpurun_out/r2w_e2e_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r2w_e2e_sweep.txt; grep "myqc trace\] *s" gpurun_out/r2w_e2e_sweep.err
