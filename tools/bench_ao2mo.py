"""Time the device AO->MO transformation (MP2/RHF shape: occ, vrt, occ, vrt blocks) on a workload's
packed array resident in HBM: ms per transformation and FP64 TFLOP/s of the flops the library
executes (myqc_ao2mo_flops) against the measured DFMA peak.  One JSON line per (workload, tile kind).
usage: bench_ao2mo.py [workload ...]   (default: h2o_16 h2o_32 h2o_64)"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import myqc_b200 as Q
from myqc_b200 import molecules

INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
names = sys.argv[1:] or ["h2o_16", "h2o_32", "h2o_64"]
peak = Q.fp64_peak(0)
dmma = Q.dmma_peak(0)
print(json.dumps({"fp64_dfma_peak_tflops": peak, "fp64_dmma_peak_tflops": dmma}), flush=True)
st = torch.cuda.current_stream().cuda_stream
for name in names:
    with tempfile.TemporaryDirectory() as d:
        s = Q.make_job(d, molecules.zmat(name), INP)
    plan = Q.Plan(s, device=0)
    packed = torch.empty(plan.out_elems, dtype=torch.float64, device="cuda")
    plan.execute(packed.data_ptr(), st)
    torch.cuda.synchronize()
    plan.close()
    n = s.norb
    nocc = int(np.sum(s.atoms)) // 2
    nvrt = n - nocc
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))          # orthonormal columns as stand-in orbitals
    c = torch.from_numpy(np.asfortranarray(q).ravel(order="F").copy()).cuda()
    co, cv = c.data_ptr(), c.data_ptr() + 8 * n * nocc          # Cm(:,0:nocc-1), Cm(:,nocc:ntot-1)
    out = torch.empty(nocc * nvrt * nocc * nvrt, dtype=torch.float64, device="cuda")
    flops = Q.ao2mo_flops(n, nocc, nvrt, nocc, nvrt)
    ws_bytes = Q.ao2mo_workspace_bytes(n, nocc, nvrt, nocc, nvrt)
    ws = torch.empty(ws_bytes // 8 + 2, dtype=torch.float64, device="cuda")
    kinds = os.environ.get("AO2MO_BENCH_KINDS", "mma,simt" if name != "h2o_64" else "mma").split(",")
    for kind in kinds:
        os.environ["MYQC_AO2MO_GEMM"] = kind
        fn = lambda: Q.ao2mo_transform_ws(packed.data_ptr(), n, co, nocc, cv, nvrt, co, nocc, cv, nvrt, out.data_ptr(),
                                          ws.data_ptr(), ws_bytes, st)
        fn(); torch.cuda.synchronize()
        reps = int(os.environ.get("AO2MO_BENCH_REPS", "3" if name != "h2o_64" else "2"))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        # (ia|jb) = (jb|ia): a size-independent check of the whole pipeline at full size
        o4 = out.view(nvrt, nocc, nvrt, nocc)  # C-order view of the Fortran (p,q,r,s) array: [s][r][q][p]
        sym = float((o4 - o4.permute(2, 3, 0, 1)).abs().max())
        print(json.dumps({"workload": name, "tiles": kind, "norb": n, "nocc": nocc, "nvrt": nvrt, "ms": ms,
                          "flops": flops, "tflops": flops / (ms * 1e-3) / 1e12, "fp64_peak_tflops_measured": peak, "dmma_peak_tflops_measured": dmma,
                          "frac": flops / (ms * 1e-3) / 1e12 / peak, "max_asymmetry": sym,
                          "out_gb": 8e-9 * out.numel(), "workspace_gb": 1e-9 * ws_bytes}), flush=True)
    del packed, out, ws
    torch.cuda.empty_cache()
