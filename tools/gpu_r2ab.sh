#!/bin/bash
# HEAD check: default bench line (refreshes profiles/r2g_bench.json), smoke
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2h_smoke.log
timeout 600 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench.json"))
r=d["roofline"]
print("ms/step %.4f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e %.1f ms"%d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "frac %.3f"%r["frac"], "traffic %.4g"%r["traffic"], "cpu", d["cpu_baseline"]["value"])
print(d["config"]); print(r["traffic_source"])
PY
