#!/bin/bash
# new shard cuts: parity file (shard tests), then every shard of the 2/4/8-way splits replayed on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2j_pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2j_pytest_parity.log
PP_MODES=auto timeout 500 python tools/exp_shard_times.py h2o_64 8 4 2 > gpurun_out/r2j_shard_times.txt 2> gpurun_out/r2j_shard_times.err; echo "rc=$?"; grep "^shard\|^---" gpurun_out/r2j_shard_times.txt | cut -c1-200
PP_MODES=auto timeout 300 python tools/exp_shard_times.py c20h42 8 4 2 > gpurun_out/r2j_shard_times_c20h42.txt 2>> gpurun_out/r2j_shard_times.err; grep "^shard\|^---" gpurun_out/r2j_shard_times_c20h42.txt | cut -c1-200
