"""End-to-end timing of the host-buffer C-ABI call (plan build, H2D, kernels, device -> pinned host) with
the library's own breakdown (MYQC_TRACE=1 on stderr).  usage: bench_e2e.py [workload] [reps]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import myqc_b200 as Q
from myqc_b200 import molecules

INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
name = sys.argv[1] if len(sys.argv) > 1 else "h2o_64"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with tempfile.TemporaryDirectory() as d:
    s = Q.make_job(d, molecules.zmat(name), INP)
host = torch.empty(s.nunique, dtype=torch.float64, pin_memory=True)
h = host.numpy()
Q.eri_packed_shard(s, h)  # warm-up
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); Q.eri_packed_shard(s, h); ts.append(time.perf_counter() - t0)
print(f"{name}: e2e {1e3 * min(ts):.1f} ms best, {1e3 * sum(ts) / len(ts):.1f} ms mean over {reps}; "
      f"{s.nunique / min(ts):.4g} unique ERIs/s; checksum {float(h.sum()):.12f}; "
      f"MYQC_SPARSE_D2H={os.environ.get('MYQC_SPARSE_D2H', '')} MYQC_HOST_THREADS={os.environ.get('MYQC_HOST_THREADS', '')}", flush=True)
