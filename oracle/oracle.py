"""CPU oracle for the myQC int2e hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (myqc_b200) never does.

What lives here
  * ctypes loader for oracle/myqc_oracle.c (the literal C restatement of
    src/integrals/int2e.f90 + auxilary.f90 + int1e.f90), built by oracle/Makefile.
  * numpy/pure-Python restatements of the *input layer* the integrals need:
      parse_zmat     <- src/parser/parser.f90:435-547,622-680 (ZMAT -> atoms, bohr xyz)
      build_basis    <- src/myQC/basis.f90:23-226            (mybasis -> set/bas/setinfo/basinfo)
      read_ftab      <- src/integrals/int2e.f90:161-163       (Fortran unformatted record)
  * scf_rhf / scf_uhf <- src/scf/scf.f90:720-1117, src/I2G/*.f90, src/dens/dens.f90: the
    fixed-point SCF that turns ERIs into orbital energies, used ONLY to pin the oracle to
    the reference's own MOLDEN outputs (examples/O/singlet, examples/Be, examples/NO).

Parity status: "pinned through SCF orbital energies only" (the reference has no per-integral
golden vectors and cannot be compiled here -- no Fortran compiler in the image).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmyqc_oracle.so")

A2B = 1.8897161646320724  # parser.f90:14
ELEMENTS = ["H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne"]  # basis.f90:68
# parser.f90:449 (single-precision literals, all exactly representable)
MASS = [1.0, 4.0, 7.0, 9.0, 11.0, 12.0, 14.0, 16.0, 19.0, 20.0]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "myqc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "libmyqc_oracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        llp = ctypes.POINTER(ctypes.c_longlong)
        L.oracle_int2e_dense.argtypes = [ctypes.c_int, dp, dp, ip, dp, ip, dp, dp, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, llp]
        L.oracle_int2e_dense.restype = ctypes.c_int
        L.oracle_int2e_packed.argtypes = [ctypes.c_int, dp, dp, ip, dp, ip, dp, dp]
        L.oracle_int2e_packed.restype = ctypes.c_int
        L.oracle_int2e_rows.argtypes = [ctypes.c_int, dp, dp, ip, dp, ip, dp, ctypes.c_int, llp, dp]
        L.oracle_int2e_rows.restype = ctypes.c_int
        L.oracle_int1e.argtypes = [ctypes.c_int, dp, ip, dp, ip, dp, ip, dp, dp, dp]
        L.oracle_int1e.restype = ctypes.c_int
        L.oracle_int2e_sample.argtypes = [ctypes.c_int, dp, dp, ip, dp, ip, dp, ctypes.c_int, ip, ip, ctypes.c_int, llp]
        L.oracle_int2e_sample.restype = ctypes.c_int
        L.oracle_boys.argtypes = [dp, ctypes.c_int, ctypes.c_double, dp]
        L.oracle_boys.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


# --------------------------------------------------------------------------------------
# input layer
# --------------------------------------------------------------------------------------
@dataclass
class Molecule:
    atoms: np.ndarray  # int32 [nnuc] atomic numbers
    xyz: np.ndarray  # float64 [nnuc,3] bohr, COM-shifted exactly as parser.f90 does
    options: dict

    @property
    def nnuc(self):
        return len(self.atoms)

    def xyz_fortran(self) -> np.ndarray:
        """xyz(0:nnuc-1,0:2) column-major flattened: xyz[i + nnuc*c]"""
        return np.ascontiguousarray(self.xyz.T).reshape(-1).copy()


def parse_zmat(text: str) -> Molecule:
    """parser.f90: cartesian (:622-680), read_options (:683-735, keys case-sensitive),
    build (:435-547).  T10: the COM accumulator `temp` is uninitialised in the reference;
    temp=0 reproduces examples/NO/MOLDEN:3-4."""
    lines = [ln for ln in text.splitlines()]
    toks = [ln.split() for ln in lines if ln.strip()]
    assert toks[0][0] == "CARTESIAN", "only CARTESIAN input is supported (parser.f90:77-85)"
    atoms, coords = [], []
    k = 1
    while toks[k][0] != "END":
        atoms.append(ELEMENTS.index(toks[k][0]) + 1)
        coords.append([float(x) for x in toks[k][1:4]])
        k += 1
    opts = {}
    for t in toks[k + 1:]:
        if len(t) >= 2:
            opts[t[0]] = t[1]
    units = 1 if opts.get("UNITS=", "ANGSTROM").upper().startswith("B") else 0
    atoms = np.array(atoms, dtype=np.int32)
    xyz = np.array(coords, dtype=np.float64)
    com = np.zeros(3)
    temp = 0.0
    for i in range(len(atoms)):
        m = MASS[atoms[i] - 1]
        com[0] = com[0] + m * xyz[i, 0]
        com[1] = com[1] + m * xyz[i, 1]
        com[2] = com[2] + m * xyz[i, 2]
        temp = temp + m
    com = com / temp
    xyz = xyz - com[None, :]
    if units == 0:
        xyz = xyz * A2B
    return Molecule(atoms=atoms, xyz=xyz, options=opts)


@dataclass
class Basis:
    set: np.ndarray  # float64 [Anum*almax]
    setinfo: np.ndarray  # int32  [2 + Anum*almax*(3+OpS)]
    bas: np.ndarray  # float64 [Anum*almax*OpS]
    basinfo: np.ndarray  # int32  [2 + 5*Omax*Anum]
    maxN: int
    maxL: int

    @property
    def nset(self):
        return int(self.setinfo[0])

    @property
    def norb(self):
        return int(self.basinfo[1])


def build_basis(mybasis_text: str, atoms, name: str = "STO-3G") -> Basis:
    """basis.f90:23-226, list-directed reads restated token-wise."""
    rows = [ln.split() for ln in mybasis_text.splitlines() if ln.strip()]
    start = next(i for i, r in enumerate(rows) if r[0] == name)  # :95-97
    Smax, Cmax, Omax, almax, OpS = (int(x) for x in rows[start + 1][:5])  # :101
    maxN, maxL = (int(x) for x in rows[start + 2][:2])  # :102
    Anum = len(atoms)
    bas = np.zeros(Anum * almax * OpS)
    basinfo = np.zeros(2 + 5 * Omax * Anum, dtype=np.int32)
    set_ = np.zeros(Anum * almax)
    setinfo = np.zeros(2 + Anum * almax * (3 + OpS), dtype=np.int32)
    setnum = 0
    orbnum = 0
    setl = 3 + OpS
    for i in range(Anum):
        sym = ELEMENTS[int(atoms[i]) - 1]
        r = next(k for k in range(start + 1, len(rows)) if rows[k][0] == sym)  # :116-119
        sec, orb, nset = (int(x) for x in rows[r + 1][:3])  # :125
        r += 2
        basinfo[0] = OpS
        basinfo[1] += orb
        setinfo[0] += nset
        setinfo[1] = setl
        for _ in range(sec):
            func, coef, pri, ang, ori = (int(x) for x in rows[r][:5])  # :136
            r += 1
            for _k in range(func):
                vals = [float(x) for x in rows[r][:coef + 1]]  # :143  val(0:coef-1), temp
                r += 1
                val, temp = vals[:coef], vals[coef]
                setn = int(np.floor(temp + 0.5))  # NINT of a non-negative id
                s = setnum + setn
                set_[s] = val[coef - 1]
                setorbs = setinfo[1 + s * setl + 1]
                setinfo[1 + s * setl + 3] = i
                if ori == -1:  # :154-157
                    setinfo[1 + s * setl + 4 + setorbs] = orbnum
                    bas[setorbs + s * OpS] = val[0]
                    setorbs += 1
                elif ori == 2:  # :160-167
                    for m in range(3):
                        setinfo[1 + s * setl + 4 + setorbs] = orbnum + m
                        bas[setorbs + s * OpS] = val[0]
                        setorbs += 1
                    if setinfo[1 + s * setl + 2] < 1:
                        setinfo[1 + s * setl + 2] = 1
                else:
                    raise ValueError("bad angular quantum number (basis.f90:170-174)")
                setinfo[1 + s * setl + 1] = setorbs
            if ori == -1:  # :182-184
                basinfo[2 + 5 * orbnum:2 + 5 * (orbnum + 1)] = [pri, ang, ori, func, i]
                orbnum += 1
            else:  # :187-191
                for m in range(3):
                    basinfo[2 + 5 * (orbnum + m):2 + 5 * (orbnum + m + 1)] = [pri, ang, m, func, i]
                orbnum += 3
        setnum = int(setinfo[0])
    return Basis(set=set_, setinfo=setinfo, bas=bas, basinfo=basinfo, maxN=maxN, maxL=maxL)


def read_ftab(path: str) -> np.ndarray:
    """Fortran unformatted sequential record holding Ft(0:120,0:22), column-major
    (int2e.f90:118,161-163).  Returns the flat array ft[t + 121*j]."""
    raw = open(path, "rb").read()
    n0 = int(np.frombuffer(raw[:4], dtype="<i4")[0])
    assert n0 == 121 * 23 * 8 and len(raw) == n0 + 8, "unexpected Ftab record"
    assert int(np.frombuffer(raw[-4:], dtype="<i4")[0]) == n0
    return np.frombuffer(raw[4:4 + n0], dtype="<f8").copy()


# --------------------------------------------------------------------------------------
# integral drivers (C)
# --------------------------------------------------------------------------------------
def _args(mol: Molecule, b: Basis, ft: np.ndarray):
    xyz = mol.xyz_fortran()
    keep = (xyz, b.set, b.setinfo, b.bas, b.basinfo, ft)
    return keep, (mol.nnuc, _dp(xyz), _dp(b.set), _ip(b.setinfo), _dp(b.bas), _ip(b.basinfo), _dp(ft))


def int2e_dense(mol, b, ft, fill=True, a_stride=1, a_offset=0, compute_only=False):
    """Literal nset^4 loop (int2e.f90:192-284) -> dense XX[i,j,g,h] (numpy index order i,j,g,h
    on a Fortran-ordered array).  Returns (xx, stats)."""
    keep, a = _args(mol, b, ft)
    n = b.norb
    stats = np.zeros(2, dtype=np.int64)
    if compute_only:
        xx, xp = None, ctypes.POINTER(ctypes.c_double)()
    else:
        xx = np.zeros(n ** 4)
        xp = _dp(xx)
    rc = lib().oracle_int2e_dense(*a, xp, int(fill), a_stride, a_offset,
                                  stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    assert rc == 0
    if xx is not None:
        xx = xx.reshape((n, n, n, n), order="F")
    return xx, stats


def int2e_packed(mol, b, ft):
    keep, a = _args(mol, b, ft)
    n = b.norb
    npair = n * (n + 1) // 2
    out = np.zeros(npair * (npair + 1) // 2)
    assert lib().oracle_int2e_packed(*a, _dp(out)) == 0
    return out


def int2e_rows(mol, b, ft, rows, nthreads=None):
    """Complete rows (P|all P') of the packed array; rows are independent and go to `nthreads` workers
    (default: all host threads)."""
    keep, a = _args(mol, b, ft)
    n = b.norb
    npair = n * (n + 1) // 2
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    out = np.zeros((len(rows), npair))
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    assert lib().oracle_int2e_rows_mt(*a, len(rows), rows.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)),
                                      _dp(out), int(nthreads)) == 0
    return out


def int2e_diag_unscreened(mol, b, ft, nthreads=None):
    """(P|P) for every function pair P without the EIJ*EGH screen (Schwarz factors; analysis only)."""
    keep, a = _args(mol, b, ft)
    n = b.norb
    out = np.zeros(n * (n + 1) // 2)
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    L = lib()
    L.oracle_int2e_diag_unscreened.restype = ctypes.c_int
    assert L.oracle_int2e_diag_unscreened(*a, _dp(out), int(nthreads)) == 0
    return out


def int2e_sample(mol, b, ft, a_list, b_list, nthreads=1):
    """Time-only run of the reference's per-(a,b) work for the listed ordered set pairs.
    Returns (seconds, surviving ordered quartets processed)."""
    import time
    keep, a = _args(mol, b, ft)
    al = np.ascontiguousarray(a_list, dtype=np.int32)
    bl = np.ascontiguousarray(b_list, dtype=np.int32)
    stats = np.zeros(2, dtype=np.int64)
    t0 = time.perf_counter()
    rc = lib().oracle_int2e_sample(*a, len(al), _ip(al), _ip(bl), int(nthreads),
                                   stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    dt = time.perf_counter() - t0
    assert rc == 0
    return dt, int(stats[0])


def int1e(mol, b, ft):
    """Returns (S, H) as [norb,norb] arrays (int1e.f90:132-280)."""
    xyz = mol.xyz_fortran()
    n = b.norb
    S = np.zeros(n * n)
    H = np.zeros(n * n)
    atoms = np.ascontiguousarray(mol.atoms, dtype=np.int32)
    rc = lib().oracle_int1e(mol.nnuc, _dp(xyz), _ip(atoms), _dp(b.set), _ip(b.setinfo), _dp(b.bas),
                            _ip(b.basinfo), _dp(ft), _dp(S), _dp(H))
    assert rc == 0
    return S.reshape((n, n), order="F"), H.reshape((n, n), order="F")


def boys(Q: int, T: float, ft: np.ndarray) -> np.ndarray:
    Fj = np.zeros(Q + 1)
    lib().oracle_boys(_dp(Fj), Q, float(T), _dp(ft))
    return Fj


# --------------------------------------------------------------------------------------
# packed <-> dense helpers (layout of SURVEY 8b / include/myqc_eri.h)
# --------------------------------------------------------------------------------------
def pair_index(i, j, n):
    return i * n - i * (i - 1) // 2 + (j - i)


def packed_from_dense(xx: np.ndarray) -> np.ndarray:
    n = xx.shape[0]
    ii, jj = np.triu_indices(n)  # row-major upper triangle == P order
    npair = len(ii)
    out = np.empty(npair * (npair + 1) // 2)
    pos = 0
    for P in range(npair):
        out[pos:pos + npair - P] = xx[ii[P], jj[P], ii[P:], jj[P:]]
        pos += npair - P
    return out


def dense_from_packed(packed: np.ndarray, n: int) -> np.ndarray:
    ii, jj = np.triu_indices(n)
    npair = len(ii)
    M = np.zeros((npair, npair))
    iu = np.triu_indices(npair)
    M[iu] = packed
    M = M + np.triu(M, 1).T
    pid = np.zeros((n, n), dtype=np.int64)
    pid[ii, jj] = np.arange(npair)
    pid[jj, ii] = np.arange(npair)
    return M[pid[:, :, None, None], pid[None, None, :, :]]


# --------------------------------------------------------------------------------------
# SCF, for pinning only
# --------------------------------------------------------------------------------------
def nuclear_repulsion(mol: Molecule) -> float:
    """scf.f90:1122-1152"""
    e = 0.0
    for a in range(mol.nnuc):
        for b in range(a):
            e += mol.atoms[a] * mol.atoms[b] / np.linalg.norm(mol.xyz[a] - mol.xyz[b])
    return float(e)


def _eigh(F, S):
    from scipy.linalg import eigh
    return eigh(F, S)  # LAPACK DSYGV itype=1, as scf.f90:851


def scf_rhf(S, H, xx, nelec, enr, tol=1e-11, maxit=500, orbitals=False):
    """scf.f90 RHF:61-216, RHFiter:720-904, dens.f90:115-124, RHFI2G.f90:80-90.
    Returns (E_total, eps, iterations) [+ (C,) with orbitals=True: the `Cui` / `eig` files]."""
    nocc = nelec // 2
    eps, C = _eigh(H, S)  # initRHF: core guess
    D = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    e_old = 0.0
    for it in range(maxit):
        J = np.einsum("kl,ijkl->ij", D, xx)
        K = np.einsum("kl,ikjl->ij", D, xx)
        K2 = np.einsum("kl,iljk->ij", D, xx)
        G = J - 0.25 * K - 0.25 * K2
        F = H + G
        e_tot = 0.5 * np.sum(D * (F + H)) + enr
        eps, C = _eigh(F, S)
        Dn = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        mdiff = np.max(np.abs(Dn - D))
        D = Dn
        if it > 0 and mdiff < tol:
            break
        e_old = e_tot
    # energy of the converged density
    J = np.einsum("kl,ijkl->ij", D, xx)
    K = np.einsum("kl,ikjl->ij", D, xx)
    K2 = np.einsum("kl,iljk->ij", D, xx)
    F = H + J - 0.25 * K - 0.25 * K2
    e_tot = 0.5 * np.sum(D * (F + H)) + enr
    eps, Cf = _eigh(F, S)
    if orbitals:
        return float(e_tot), eps, it + 1, Cf
    return float(e_tot), eps, it + 1


def scf_uhf(S, H, xx, nA, nB, enr, tol=1e-9, maxit=2000, orbitals=False):
    """scf.f90 UHF:221-395, UHFiter:909-1117, dens.f90:213-228, UHFI2G.f90:80-93."""
    _, C = _eigh(H, S)
    Ca, Cb = C.copy(), C.copy()
    Da = Ca[:, :nA] @ Ca[:, :nA].T
    Db = Cb[:, :nB] @ Cb[:, :nB].T
    for it in range(maxit):
        Dt = Da + Db
        J = np.einsum("kl,ijkl->ij", Dt, xx)
        Ka = np.einsum("kl,ikjl->ij", Da, xx)
        Kb = np.einsum("kl,ikjl->ij", Db, xx)
        Fa, Fb = H + J - Ka, H + J - Kb
        e_tot = 0.5 * (np.sum(Da * (Fa + H)) + np.sum(Db * (Fb + H))) + enr
        ea, Ca = _eigh(Fa, S)
        eb, Cb = _eigh(Fb, S)
        Dan = Ca[:, :nA] @ Ca[:, :nA].T
        Dbn = Cb[:, :nB] @ Cb[:, :nB].T
        mdiff = max(np.max(np.abs(Dan - Da)), np.max(np.abs(Dbn - Db)))
        Da, Db = Dan, Dbn
        if it > 0 and mdiff < tol:
            break
    Dt = Da + Db
    J = np.einsum("kl,ijkl->ij", Dt, xx)
    Fa = H + J - np.einsum("kl,ikjl->ij", Da, xx)
    Fb = H + J - np.einsum("kl,ikjl->ij", Db, xx)
    e_tot = 0.5 * (np.sum(Da * (Fa + H)) + np.sum(Db * (Fb + H))) + enr
    ea, Caf = _eigh(Fa, S)
    eb, Cbf = _eigh(Fb, S)
    if orbitals:
        return float(e_tot), ea, eb, it + 1, Caf, Cbf
    return float(e_tot), ea, eb, it + 1


def electrons(mol: Molecule):
    """parser.f90:498-506"""
    charge = int(mol.options.get("CHARGE=", "0").replace("+", ""))
    mult = int(mol.options.get("MULTI=", "1"))
    unpr = mult - 1
    nelc = int(np.sum(mol.atoms)) - charge
    nB = (nelc - unpr) // 2
    nA = nB + unpr
    return nA, nB


# ----------------------------------------------------------------------------------------------
# ao2mo (src/ao2mo/ao2mo.f90) and the MP2 energies that consume its files (src/mp2/mp2.f90):
# numpy restatement used to check the device transformation (tests only)
# ----------------------------------------------------------------------------------------------
def ao2mo_idx_trans(xx, c1, c2, c3, c4):
    """idx1_trans .. idx4_trans (ao2mo.f90:1306-1439) in the reference's order:
    B(p,q,r,s) = sum_t A(t,q,r,s) x(t,p); then index 2, 3, 4."""
    L = np.einsum("tqrs,tp->pqrs", xx, c1)   # idx1_trans :1306-1328
    M = np.einsum("ptrs,tq->pqrs", L, c2)    # idx2_trans :1343-1365
    N = np.einsum("pqts,tr->pqrs", M, c3)    # idx3_trans :1380-1402
    return np.einsum("pqrt,ts->pqrs", N, c4)  # idx4_trans :1417-1439


def ao2mo_files(kind, xx, CA, CB, nA, nB):
    """The unformatted records the reference's `ao2mo` writes, as {file name: [records]}.
    kind = "mp2_rhf" (ao2mo.f90:465-602), "mp2_uhf" (:614-904), "cis_uhf" (:919-1227)."""
    n = xx.shape[0]
    oA, vA, oB, vB = CA[:, :nA], CA[:, nA:], CB[:, :nB], CB[:, nB:]
    out = {}

    def ijab(Om, upper):
        n1, _, n3, _ = Om.shape
        if upper:  # DO i=0,nocc-2; DO j=i+1,nocc-1; WRITE Om(i,:,j,:)
            return [Om[i, :, j, :].reshape(-1, order="F") for i in range(n1 - 1) for j in range(i + 1, n3)]
        return [Om[i, :, j, :].reshape(-1, order="F") for i in range(n1) for j in range(n3)]

    if kind == "mp2_rhf":
        Om = ao2mo_idx_trans(xx, oA, vA, oA, vA)
        out["ijab_AA"] = ijab(Om, True)
        out["ijab_AB"] = ijab(Om, False)
    elif kind == "mp2_uhf":
        out["ijab_AA"] = ijab(ao2mo_idx_trans(xx, oA, vA, oA, vA), True)
        out["ijab_BB"] = ijab(ao2mo_idx_trans(xx, oB, vB, oB, vB), True)
        out["ijab_AB"] = ijab(ao2mo_idx_trans(xx, oA, vA, oB, vB), False)
    elif kind == "cis_uhf":
        def ajib(Om):  # DO j; DO b; vec(idx(i,a)) = Om(a,i,j,b), idx runs a fastest
            nv, no, nj, nb = Om.shape
            return [np.array([Om[a, i, j, b] for i in range(no) for a in range(nv)]) for j in range(nj) for b in range(nb)]

        def ajbi(Om):  # vec(idx(i,a)) = Om(a,b,j,i)
            nv, nb, nj, no = Om.shape
            return [np.array([Om[a, b, j, i] for i in range(no) for a in range(nv)]) for j in range(nj) for b in range(nb)]

        out["ajib_AA"] = ajib(ao2mo_idx_trans(xx, vA, oA, oA, vA))
        out["ajbi_AA"] = ajbi(ao2mo_idx_trans(xx, vA, vA, oA, oA))
        out["ajib_AB"] = ajib(ao2mo_idx_trans(xx, vA, oA, oB, vB))
        out["ajib_BB"] = ajib(ao2mo_idx_trans(xx, vB, oB, oB, vB))
        out["ajbi_BB"] = ajbi(ao2mo_idx_trans(xx, vB, vB, oB, oB))
    else:
        raise ValueError(kind)
    return out


def mp2_rhf_energy(records_ab, eig, nocc, nvrt):
    """mp2_rhf (mp2.f90:79-150), literally: consumes the ijab_AB records in file order.
    Returns (E(AA) = sum1, E(AB) = sum2, E(MBPT2) = 2 sum1 + sum2)."""
    ntot = nocc + nvrt
    it = iter(records_ab)
    sum1 = sum2 = 0.0
    for i in range(nocc - 1):
        for j in range(i + 1):
            m = next(it).reshape((nvrt, nvrt), order="F")
            for a in range(nvrt):
                for b in range(nvrt):
                    sum2 += m[a, b] ** 2 / (eig[i] + eig[j] - eig[a + nocc] - eig[b + nocc])
        for j in range(i + 1, nocc):
            m = next(it).reshape((nvrt, nvrt), order="F")
            for a in range(nvrt - 1):
                for b in range(a + 1):
                    sum2 += m[a, b] ** 2 / (eig[i] + eig[j] - eig[a + nocc] - eig[b + nocc])
                for b in range(a + 1, nvrt):
                    d = eig[i] + eig[j] - eig[a + nocc] - eig[b + nocc]
                    sum2 += m[a, b] ** 2 / d
                    sum1 += (m[a, b] - m[b, a]) ** 2 / d
            for b in range(nvrt):
                sum2 += m[nvrt - 1, b] ** 2 / (eig[i] + eig[j] - eig[ntot - 1] - eig[b + nocc])
    for j in range(nocc):
        m = next(it).reshape((nvrt, nvrt), order="F")
        for a in range(nvrt):
            for b in range(nvrt):
                sum2 += m[a, b] ** 2 / (eig[nocc - 1] + eig[j] - eig[a + nocc] - eig[b + nocc])
    return sum1, sum2, 2 * sum1 + sum2


def mp2_uhf_energy(files, eigA, eigB, nA, nB, ntot):
    """mp2_uhf (mp2.f90:154-237): (E(AA), E(BB), E(AB), E(MBPT2))."""
    vA, vB = ntot - nA, ntot - nB
    s1 = s2 = s3 = 0.0
    it = iter(files["ijab_AA"])
    for i in range(nA - 1):
        for j in range(i + 1, nA):
            m = next(it).reshape((vA, vA), order="F")
            for a in range(vA - 1):
                for b in range(a + 1, vA):
                    s1 += (m[a, b] - m[b, a]) ** 2 / (eigA[i] + eigA[j] - eigA[a + nA] - eigA[b + nA])
    it = iter(files["ijab_BB"])
    for i in range(nB - 1):
        for j in range(i + 1, nB):
            m = next(it).reshape((vB, vB), order="F")
            for a in range(vB - 1):
                for b in range(a + 1, vB):
                    s2 += (m[a, b] - m[b, a]) ** 2 / (eigB[i] + eigB[j] - eigB[a + nB] - eigB[b + nB])
    it = iter(files["ijab_AB"])
    for i in range(nA):
        for j in range(nB):
            m = next(it).reshape((vA, vB), order="F")
            for a in range(vA):
                for b in range(vB):
                    s3 += m[a, b] ** 2 / (eigA[i] + eigB[j] - eigA[a + nA] - eigB[b + nB])
    return s1, s2, s3, s1 + s2 + s3
