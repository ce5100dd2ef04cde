"""Sweep of the sparse device->host route of the host-buffer call on one process (pinned buffer allocated once):
chunk size, host zeroing threads, and the two halves alone (MYQC_XFER_NOHOST / MYQC_XFER_NOPUSH leave the result
incomplete and are measurement hooks only).  usage: exp_e2e_sweep.py [workload]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import myqc_b200 as Q
from myqc_b200 import molecules

INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
name = sys.argv[1] if len(sys.argv) > 1 else "h2o_64"
with tempfile.TemporaryDirectory() as d:
    s = Q.make_job(d, molecules.zmat(name), INP)
host = torch.empty(s.nunique, dtype=torch.float64, pin_memory=True)
h = host.numpy()
Q.eri_packed_shard(s, h)  # warm-up: plan, device slice, first touch of the pinned pages
ref = float(h.sum())
KEYS = ("MYQC_XFER_CHUNK", "MYQC_HOST_THREADS", "MYQC_XFER_NOHOST", "MYQC_XFER_NOPUSH", "MYQC_SPARSE_D2H", "MYQC_TRACE")
def run(tag, reps=3, **env):
    for k in KEYS: os.environ.pop(k, None)
    for k, v in env.items(): os.environ[k] = str(v)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); Q.eri_packed_shard(s, h); ts.append(time.perf_counter() - t0)
    ok = "" if ("MYQC_XFER_NOHOST" in env or "MYQC_XFER_NOPUSH" in env) else f" checksum {'same' if float(h.sum()) == ref else 'DIFFERENT'}"
    print(f"{tag:40s} best {1e3 * min(ts):7.1f} ms  mean {1e3 * sum(ts) / len(ts):7.1f} ms  d2h {Q.last_d2h_bytes() / 1e9:6.2f} GB{ok}", flush=True)
KEYS = KEYS + ("MYQC_HOST_MEMSET",)
run("chunk 2048 B, 8 thr, streaming (default)")
run("chunk 2048 B, 8 thr, memset", MYQC_HOST_MEMSET=1)
for c in (256, 128, 64, 32):
    for t in (4, 8, 12, 16):
        run(f"chunk {8 * c} B, {t} thr, streaming", MYQC_XFER_CHUNK=c, MYQC_HOST_THREADS=t)
for c in (256, 64, 32):
    run(f"chunk {8 * c} B, push only", MYQC_XFER_CHUNK=c, MYQC_XFER_NOHOST=1)
    for t in (4, 8, 16):
        run(f"chunk {8 * c} B, host zeros only, {t} thr", MYQC_XFER_CHUNK=c, MYQC_XFER_NOPUSH=1, MYQC_HOST_THREADS=t)
run("chunk 512 B, 16 thr, traced", reps=1, MYQC_XFER_CHUNK=64, MYQC_HOST_THREADS=16, MYQC_TRACE=1)
h[:] = 1.0
run("chunk 512 B after poisoning", reps=1, MYQC_XFER_CHUNK=64)
h[:] = 1.0
run("chunk 256 B after poisoning", reps=1, MYQC_XFER_CHUNK=32)
