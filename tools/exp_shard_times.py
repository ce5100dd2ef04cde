"""Per-shard timings on ONE GPU: every shard of an N-way split of a workload is planned and run on device 0 (shards are
independent), with the per-launch CUDA-event times of the serialised pass and the time of the production execute().
Shows which class of which shard the cut model under- or over-estimates, and the (SP SP|SP SP) kernel choice per piece.
usage: exp_shard_times.py [workload] [nshards ...]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import myqc_b200 as Q
from myqc_b200 import molecules

INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
name = sys.argv[1] if len(sys.argv) > 1 else "h2o_64"
splits = [int(x) for x in sys.argv[2:]] or [8]
with tempfile.TemporaryDirectory() as d:
    s = Q.make_job(d, molecules.zmat(name), INP)
stream = torch.cuda.current_stream().cuda_stream
for nsh in splits:
    for mode in (os.environ.get("PP_MODES", "auto,warp,slices").split(",")):
        if mode == "auto": os.environ.pop("MYQC_PP_KERNEL", None)
        else: os.environ["MYQC_PP_KERNEL"] = mode
        print(f"--- {name}, {nsh} shards, (SP SP|SP SP) kernel: {mode}", flush=True)
        for sh in range(nsh):
            plan = Q.Plan(s, device=0, shard=sh, nshards=nsh)
            out = torch.empty(max(plan.out_elems, 1), dtype=torch.float64, device="cuda")
            launches = plan.launches()
            for _ in range(2): plan.execute(out.data_ptr(), stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(3):
                e0.record(); plan.execute(out.data_ptr(), stream); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            acc = np.zeros(len(launches))
            for _ in range(3): acc += np.array(plan.execute_timed(out.data_ptr(), stream))
            acc /= 3
            per = {}
            for (cls, tri, rows), t in zip(launches, acc): per[cls] = per.get(cls, 0.0) + float(t)
            exq, _ = plan.executed_quartets()
            print(f"shard {sh}: slice {8e-9 * plan.out_elems:5.2f} GB  execute {min(ts):6.3f} ms  serial {acc.sum():6.3f} | fill {per.get(-1, 0):.3f} | "
                  + " ".join(f"{Q.CLASS_NAMES[c]} {per.get(c, 0):.3f}" for c in range(6))
                  + " | M prim. quartets " + " ".join(f"{exq[c] / 1e6:.2f}" for c in range(6))
                  + " | ps each " + " ".join(f"{1e9 * per.get(c, 0) / max(exq[c], 1):.1f}" for c in range(6)), flush=True)
            del plan, out
        if nsh == 1: break
