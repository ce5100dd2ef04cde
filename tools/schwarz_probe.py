"""What would a Schwarz bound remove beyond the reference's own screen?  

The reference keeps a primitive quartet iff EIJ*EGH >= 1e-14 (int2e.f90:257); the product keeps a CONTRACTED
shell quartet (u|v) iff emax_u*emax_v >= 1e-14 (then at least one primitive quartet passes) -- exactly the
reference's rule, nothing more.  A Schwarz skip would add: drop (u|v) when Q_u*Q_v < tau, with
Q_u = max over the function pairs (ij) of shell pair u of sqrt((ij|ij)).  By |(ij|kl)| <= sqrt((ij|ij)(kl|kl)) every
integral of a dropped quartet is below tau in magnitude, and each integral belongs to one contracted quartet, so the
omission per integral is < tau (it does not accumulate).  The diagonal integrals must be the UNSCREENED ones: the
reference's own (ij|ij) is exactly zero once E_ij < 1e-7.

usage: python tools/schwarz_probe.py [workload ...]      (CPU only; uses the oracle, i.e. test infrastructure)
Prints one markdown row per workload and threshold: contracted quartets and primitive quartets the reference rule
keeps, how many of them a Schwarz skip at tau would drop, and -- where the oracle's full packed array is affordable --
the largest |integral| among the dropped quartets (must be < tau)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import INPUTS, oracle_system  # noqa: E402
from oracle import oracle as O  # noqa: E402


def shell_pairs(mol, b):
    setl = int(b.setinfo[1])
    shells = {}
    for st in range(b.nset):
        info = b.setinfo[2 + setl * st: 2 + setl * (st + 1)]
        key = (int(info[2]), tuple(int(x) for x in info[3:3 + int(info[0])]))
        shells.setdefault(key, []).append(float(b.set[st]))
    keys = sorted(shells, key=lambda k: k[1][0])
    cen = np.array([mol.xyz[k[0]] for k in keys])
    return keys, [np.array(shells[k]) for k in keys], cen


def main(names):
    ft = O.read_ftab(os.path.join(INPUTS, "Ftab"))
    mb = open(os.path.join(INPUTS, "mybasis")).read()
    print("| workload | tau | contracted quartets kept by the reference rule | dropped by Schwarz | share | primitive quartets kept | dropped | share | max abs integral among dropped (oracle) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name in names:
        mol, b, _ = oracle_system(name, (ft, mb))
        n = b.norb
        keys, exps, cen = shell_pairs(mol, b)
        ns = len(keys)
        diag = O.int2e_diag_unscreened(mol, b, ft)
        qf = np.sqrt(np.maximum(diag, 0.0))
        fn_shell = np.zeros(n, dtype=int)
        for s, k in enumerate(keys):
            for o in k[1]:
                fn_shell[o] = s
        ii, jj = np.triu_indices(n)
        # per shell pair (A <= B): Q = max sqrt((ij|ij)), the sorted primitive prefactors E
        Q = np.zeros((ns, ns))
        np.maximum.at(Q, (fn_shell[ii], fn_shell[jj]), qf)
        A_idx, B_idx = np.triu_indices(ns)
        E = []
        for A, B in zip(A_idx, B_idx):
            a = exps[A][:, None]; c = exps[B][None, :]
            r2 = float(((cen[A] - cen[B]) ** 2).sum())
            e = np.sort(np.exp(-a * c / (a + c) * r2).ravel())[::-1]
            E.append(e)
        emax = np.array([e[0] for e in E])
        Qp = Q[A_idx, B_idx]
        live = emax >= 1e-14
        idx = np.nonzero(live)[0]
        # canonical contracted quartets u <= v kept by the reference rule, and their surviving primitive quartets
        order = idx[np.argsort(-emax[idx])]
        em = emax[order]; qq = Qp[order]
        Eflat = [E[k] for k in order]
        packed = None
        if n <= 60:
            packed = O.int2e_packed(mol, b, ft)
        npair = n * (n + 1) // 2
        Pfn = ii * n - ii * (ii - 1) // 2 + (jj - ii)
        fp_of_pair = {}
        if packed is not None:
            for t, (A, B) in enumerate(zip(A_idx, B_idx)):
                fp_of_pair[t] = Pfn[(fn_shell[ii] == A) & (fn_shell[jj] == B)]
        Epad = np.zeros((len(order), 9))
        for t, e in enumerate(Eflat):
            Epad[t, :len(e)] = e
        # large molecules: every `stride`-th owner pair x, counts scaled up (the pairs are sorted by emax, so a
        # stride sample covers all magnitudes evenly)
        stride = 1 if len(order) <= 4000 else 16
        for tau in (1e-10, 1e-11, 1e-12, 1e-13):
            kept_q = drop_q = kept_p = drop_p = 0
            worst = 0.0
            for x in range(0, len(order), stride):
                # partners y >= x in this order with em[x]*em[y] >= 1e-14 (a prefix of the descending list)
                hi = np.searchsorted(-em, -(1e-14 / em[x]), side="right")
                if hi <= x:
                    continue
                ys = np.arange(x, hi)
                ys = ys[em[x] * em[ys] >= 1e-14]
                kept_q += len(ys)
                dmask = qq[x] * qq[ys] * (1.0 + 1e-6) < tau
                drop_q += int(dmask.sum())
                # primitive quartets: pairs of primitive prefactors with product >= 1e-14
                cnt = (Epad[x][None, :, None] * Epad[ys][:, None, :] >= 1e-14).sum(axis=(1, 2))
                kept_p += int(cnt.sum())
                drop_p += int(cnt[dmask].sum())
                if packed is not None:
                    for y in ys[dmask]:
                        P1 = fp_of_pair[order[x]]; P2 = fp_of_pair[order[y]]
                        lo = np.minimum(P1[:, None], P2[None, :]); hi2 = np.maximum(P1[:, None], P2[None, :])
                        worst = max(worst, float(np.abs(packed[lo * npair - lo * (lo - 1) // 2 + (hi2 - lo)]).max()))
            kept_q *= stride; drop_q *= stride; kept_p *= stride; drop_p *= stride
            note = "" if stride == 1 else f" (every {stride}th owner pair, scaled)"
            print(f"| {name} | {tau:.0e} | {kept_q} | {drop_q} | {100.0 * drop_q / max(kept_q, 1):.1f} % | {kept_p} | {drop_p} | "
                  f"{100.0 * drop_p / max(kept_p, 1):.1f} % | " + (f"{worst:.2e}" if packed is not None else "not enumerated") + note + " |")
            sys.stdout.flush()


if __name__ == "__main__":
    main(sys.argv[1:] or ["CO2", "h2o_4", "h2o_8"])
