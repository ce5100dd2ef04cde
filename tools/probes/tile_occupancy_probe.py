"""Round-2 design probe (host only, numpy): how full would the lanes of a warp be if the quartet space
of (H2O)_64 were walked output-stationary, i.e. one task = (owner shell pair u = (A,B); partner first shell C;
partner kind), lanes over the surviving partner shells D?  Uses the reference's screen on the largest
primitive prefactors of the shell pairs (contracted level).  usage: tile_occupancy_probe.py [workload]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
from myqc_b200 import molecules

INP = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "inputs")
name = sys.argv[1] if len(sys.argv) > 1 else "h2o_64"
mol = O.parse_zmat(molecules.zmat(name))
b = O.build_basis(open(os.path.join(INP, "mybasis")).read(), mol.atoms)
setl = int(b.setinfo[1])
nset = b.nset
cen = np.array([b.setinfo[1 + s * setl + 3] for s in range(nset)])
lmax = np.array([b.setinfo[1 + s * setl + 2] for s in range(nset)])
first = np.array([b.setinfo[1 + s * setl + 4] for s in range(nset)])  # first orbital id of the set
# shells = sets on one centre with the same first orbital (S, or SP)
keys = sorted({(int(first[s]), int(cen[s]), int(lmax[s])) for s in range(nset)})
sh_first = np.array([k[0] for k in keys]); sh_cen = np.array([k[1] for k in keys]); sh_sp = np.array([k[2] for k in keys])
sh_amin = np.array([min(b.set[s] for s in range(nset) if first[s] == k[0] and cen[s] == k[1]) for k in keys])
nsh = len(keys)
xyz = mol.xyz[sh_cen]
r2 = ((xyz[:, None, :] - xyz[None, :, :]) ** 2).sum(-1)
emax = np.exp(-(sh_amin[:, None] * sh_amin[None, :]) * r2 / (sh_amin[:, None] + sh_amin[None, :]))  # largest primitive prefactor
kind = sh_sp[:, None] + sh_sp[None, :]
print(f"{name}: {nsh} shells, {nsh * (nsh + 1) // 2} shell pairs, {int((np.triu(emax) >= 1e-14).sum())} with emax >= 1e-14")

iu = np.triu_indices(nsh)
pairs_A, pairs_B = iu
pe = emax[iu]
order = np.argsort(-pe)
tot_q = 0; tot_slots = 0; tasks = 0
hist = np.zeros(34, dtype=np.int64)
# per partner first shell C: partner shells D >= C sorted by emax(C,D) descending, per kind
per_C = []
for C in range(nsh):
    D = np.arange(C, nsh)
    e = emax[C, D]; k = kind[C, D]
    per_C.append([np.sort(e[k == t])[::-1] for t in range(3)])
keep = pe >= 1e-14
uA = pairs_A[keep]; thr_u = 1e-14 / pe[keep]
group = int(sys.argv[2]) if len(sys.argv) > 2 else 1  # partner first shells handled together by one task
for C0 in range(0, nsh, group):
    for t in range(3):
        cnt = np.zeros(len(uA), dtype=np.int64)
        for C in range(C0, min(nsh, C0 + group)):
            arr = per_C[C][t]
            if len(arr):
                c = np.searchsorted(-arr, -thr_u, side="right")
                cnt += np.where(uA <= C, c, 0)  # owner rule: the pair with the smaller first shell owns the quartet
        nz = cnt[cnt > 0]
        tot_q += int(nz.sum()); tot_slots += int((32 * ((nz + 31) // 32)).sum()); tasks += len(nz)
        hist += np.bincount(np.minimum(nz, 33), minlength=34)
print(f"owner-side tasks (u; {group} partner first shell(s); kind): {tasks}, shell quartets {tot_q}, lane slots {tot_slots}, lane efficiency {tot_q / tot_slots:.3f}")
print("share of tasks with <= 4 / <= 8 / <= 16 / <= 32 partners:", [round(float(hist[:k + 1].sum() / hist.sum()), 3) for k in (4, 8, 16, 32)])
