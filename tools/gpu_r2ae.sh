#!/bin/bash
mkdir -p gpurun_out
PP_MODES=auto timeout 500 python tools/exp_shard_times.py h2o_64 8 4 2 > gpurun_out/r2k_shard_times.txt 2> gpurun_out/r2k_shard_times.err; echo "rc=$?"; grep "^shard\|^---" gpurun_out/r2k_shard_times.txt | cut -c1-175
