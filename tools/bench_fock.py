"""Time the device G(D) build on a workload's packed array (resident in HBM): ms per build and the
HBM read rate it corresponds to.  usage: bench_fock.py [workload] [reps]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import myqc_b200 as Q
from myqc_b200 import molecules

INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
name = sys.argv[1] if len(sys.argv) > 1 else "h2o_64"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
with tempfile.TemporaryDirectory() as d:
    s = Q.make_job(d, molecules.zmat(name), INP)
plan = Q.Plan(s, device=0)
out = torch.empty(plan.out_elems, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
plan.execute(out.data_ptr(), st)
torch.cuda.synchronize()
n = s.norb
rng = np.random.default_rng(1)
c = rng.standard_normal((n, n // 2))
dm = torch.from_numpy(np.asfortranarray(c @ c.T / n).ravel(order="F").copy()).cuda()
g = torch.empty(n * n, dtype=torch.float64, device="cuda")
ga = torch.empty_like(g); gb = torch.empty_like(g)
for label, fn in (("rhf", lambda: Q.fock_rhf_device(out.data_ptr(), 0, plan.out_elems, n, dm.data_ptr(), g.data_ptr(), st)),
                  ("uhf", lambda: Q.fock_uhf_device(out.data_ptr(), 0, plan.out_elems, n, dm.data_ptr(), dm.data_ptr(), ga.data_ptr(), gb.data_ptr(), st))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name} {label}: {ms:.2f} ms per G build, {8e-9 * plan.out_elems / (ms * 1e-3):.0f} GB/s of packed ERIs read, "
          f"|G|max {float((g if label == 'rhf' else ga).abs().max()):.6f}")
# with the sparsity mask (built once per integral evaluation, reused by every SCF iteration)
mask = torch.zeros(Q.fock_mask_words(n), dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
Q.fock_mask_build(out.data_ptr(), 0, plan.out_elems, n, mask.data_ptr(), st); torch.cuda.synchronize()
e0.record(); Q.fock_mask_build(out.data_ptr(), 0, plan.out_elems, n, mask.data_ptr(), st); e1.record(); torch.cuda.synchronize()
print(f"{name} mask build: {e0.elapsed_time(e1):.2f} ms ({8e-9 * plan.out_elems / (e0.elapsed_time(e1) * 1e-3):.0f} GB/s), "
      f"nonzero words {int((mask != 0).sum().item())} of {mask.numel()}")
g_ref = g.clone()
for label, fn in (("rhf masked", lambda: Q.fock_rhf_masked_device(out.data_ptr(), 0, plan.out_elems, n, dm.data_ptr(), mask.data_ptr(), g.data_ptr(), st)),
                  ("uhf masked", lambda: Q.fock_uhf_masked_device(out.data_ptr(), 0, plan.out_elems, n, dm.data_ptr(), dm.data_ptr(), mask.data_ptr(), ga.data_ptr(), gb.data_ptr(), st))):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name} {label}: {ms:.2f} ms per G build ({8e-9 * plan.out_elems / (ms * 1e-3):.0f} GB/s of packed ERIs covered)")
Q.fock_rhf_device(out.data_ptr(), 0, plan.out_elems, n, dm.data_ptr(), g_ref.data_ptr(), st)
Q.fock_rhf_masked_device(out.data_ptr(), 0, plan.out_elems, n, dm.data_ptr(), mask.data_ptr(), g.data_ptr(), st)
torch.cuda.synchronize()
print(f"max |G_masked - G| = {float((g - g_ref).abs().max()):.3e}")
nz = int((out != 0).sum().item())
print(f"nonzero unique ERIs: {nz} of {plan.out_elems} ({100.0 * nz / plan.out_elems:.2f} %)")
