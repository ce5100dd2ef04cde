"""myqc_b200 -- B200-native replacement for myQC's `int2e` two-electron-integral engine.

Host-side mirror of the reference boundary (src/integrals/int2e.f90):

    PROGRAM int2e            -> int2e_main(workdir)            (file in, `XX` file out)
    CALL proc2e(...)         -> proc2e(...) / eri_dense(...)   (arrays in, dense XX out)
                                eri_packed(...)                (8-fold-unique packed array)
    getenv / buildBasis      -> read_env(dir), build_basis(path, atoms)
    READ(1) Ft               -> read_ftab(path)
    WRITE(42) XX             -> write_xx(path, xx), read_xx(path, norb)

Everything computes through the C-ABI shared library `csrc/libmyqc_eri.so`
(include/myqc_eri.h) whose kernels are hand-written CUDA for sm_100a.  There is no CPU
fallback: importing works without a GPU (so the symbols can be checked), computing does not.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MYQC_LIB: experiments only (tools/build_variants.sh builds kernel variants next to the library)
LIB_PATH = os.environ.get("MYQC_LIB") or os.path.join(_HERE, "csrc", "libmyqc_eri.so")

MYQC_OK = 0
ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_BAD_ARG, ERR_IO, ERR_NOMEM = -1, -2, -3, -4, -5, -6

# every symbol include/myqc_eri.h declares
EXPORTS = [
    "myqc_last_error", "myqc_device_count", "myqc_eri_dense", "myqc_eri_packed",
    "myqc_eri_plan_create", "myqc_eri_plan_out_offset", "myqc_eri_plan_out_elems",
    "myqc_eri_plan_execute", "myqc_eri_plan_stats", "myqc_eri_plan_destroy",
    "myqc_eri_expand_dense", "myqc_read_env", "myqc_build_basis", "myqc_read_ftab",
    "myqc_write_xx", "myqc_read_xx", "myqc_int2e_main", "myqc_eri_shard_layout", "myqc_eri_shard_model", "myqc_host_zero",
    "myqc_eri_canonical_stats", "myqc_eri_plan_launch_count", "myqc_eri_plan_launch_info",
    "myqc_eri_plan_execute_timed", "myqc_fp64_peak", "myqc_eri_packed_shard", "myqc_write_xx_ex", "myqc_eri_last_d2h_bytes",
    "myqc_eri_release_cache", "myqc_eri_plan_executed_quartets",
    # include/myqc_fock.h
    "myqc_fock_rhf", "myqc_fock_uhf", "myqc_fock_rhf_host", "myqc_fock_uhf_host",
    "myqc_fock_mask_words", "myqc_fock_mask_build", "myqc_fock_rhf_masked", "myqc_fock_uhf_masked",
    # include/myqc_int1e.h
    "myqc_int1e", "myqc_int1e_main",
    # include/myqc_parse.h
    "myqc_parse_zmat", "myqc_parse_main",
    # include/myqc_ao2mo.h
    "myqc_ao2mo_transform", "myqc_ao2mo_transform_host", "myqc_pack_dense", "myqc_ao2mo_main", "myqc_ao2mo_flops", "myqc_dmma_peak",
    "myqc_ao2mo_workspace_bytes", "myqc_ao2mo_transform_ws",
]


class MyQCError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"myqc_eri error {code}: {msg}")
        self.code = code


_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def lib() -> ctypes.CDLL:
    """Load libmyqc_eri.so; fail loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C myqc_b200/csrc` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_void_p, c_char_p = ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p
    common = [c_int, _dp, c_int, c_int, _dp, _ip, c_int, _dp, _ip, _dp]
    L.myqc_last_error.restype = c_char_p
    L.myqc_device_count.restype = c_int
    L.myqc_eri_dense.argtypes = common + [_dp, c_int]
    L.myqc_eri_packed.argtypes = common + [_dp, c_int]
    L.myqc_eri_plan_create.argtypes = common + [c_int, c_int, c_int, ctypes.POINTER(c_void_p)]
    L.myqc_eri_plan_out_offset.argtypes = [c_void_p]
    L.myqc_eri_plan_out_offset.restype = ctypes.c_int64
    L.myqc_eri_plan_out_elems.argtypes = [c_void_p]
    L.myqc_eri_plan_out_elems.restype = ctypes.c_int64
    L.myqc_eri_plan_execute.argtypes = [c_void_p, c_void_p, c_void_p]
    L.myqc_eri_plan_stats.argtypes = [c_void_p, _i64p, _dp, ctypes.POINTER(c_int)]
    L.myqc_eri_plan_destroy.argtypes = [c_void_p]
    L.myqc_eri_plan_destroy.restype = None
    L.myqc_eri_expand_dense.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
    L.myqc_read_env.argtypes = [c_char_p, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                ctypes.POINTER(c_int), _ip, _dp, _dp, ctypes.POINTER(c_int), _ip]
    L.myqc_build_basis.argtypes = [c_char_p, c_int, c_int, _ip, ctypes.POINTER(c_int),
                                   ctypes.POINTER(c_int), _dp, _ip, _dp, _ip, ctypes.POINTER(c_int),
                                   ctypes.POINTER(c_int), c_char_p]
    L.myqc_read_ftab.argtypes = [c_char_p, _dp]
    L.myqc_write_xx.argtypes = [c_char_p, _dp, c_int]
    L.myqc_read_xx.argtypes = [c_char_p, _dp, c_int]
    L.myqc_write_xx_ex.argtypes = [c_char_p, _dp, c_int, ctypes.c_int64]
    L.myqc_int2e_main.argtypes = [c_char_p, c_int]
    L.myqc_eri_shard_layout.argtypes = common[:-1] + [c_int, _i64p]
    L.myqc_eri_shard_model.argtypes = common[:-1] + [c_int, _dp, _dp]
    L.myqc_host_zero.argtypes = [_dp, ctypes.c_int64]
    L.myqc_host_zero.restype = None
    L.myqc_eri_canonical_stats.argtypes = [c_int, _dp, c_int, c_int, _dp, _ip, _i64p, _dp]
    L.myqc_eri_plan_launch_count.argtypes = [c_void_p]
    L.myqc_eri_plan_launch_info.argtypes = [c_void_p, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), _i64p]
    L.myqc_eri_plan_execute_timed.argtypes = [c_void_p, c_void_p, c_void_p, ctypes.POINTER(ctypes.c_float)]
    L.myqc_fp64_peak.argtypes = [c_int, _dp]
    L.myqc_eri_packed_shard.argtypes = common + [_dp, c_int, c_int, c_int, _i64p]
    L.myqc_eri_last_d2h_bytes.argtypes = []
    L.myqc_eri_last_d2h_bytes.restype = ctypes.c_int64
    L.myqc_eri_plan_executed_quartets.argtypes = [c_void_p, _i64p, _dp]
    L.myqc_eri_release_cache.argtypes = []
    L.myqc_eri_release_cache.restype = None
    c_i64 = ctypes.c_int64
    L.myqc_fock_rhf.argtypes = [c_void_p, c_i64, c_i64, c_int, c_void_p, c_void_p, c_void_p]
    L.myqc_fock_uhf.argtypes = [c_void_p, c_i64, c_i64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.myqc_fock_mask_words.argtypes = [c_int]
    L.myqc_fock_mask_words.restype = c_i64
    L.myqc_fock_mask_build.argtypes = [c_void_p, c_i64, c_i64, c_int, c_void_p, c_void_p]
    L.myqc_fock_rhf_masked.argtypes = [c_void_p, c_i64, c_i64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.myqc_fock_uhf_masked.argtypes = [c_void_p, c_i64, c_i64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.myqc_fock_rhf_host.argtypes = [_dp, c_int, _dp, _dp]
    L.myqc_fock_uhf_host.argtypes = [_dp, c_int, _dp, _dp, _dp, _dp]
    L.myqc_int1e.argtypes = [c_int, _dp, _ip, c_int, c_int, _dp, _ip, c_int, _dp, _ip, _dp, _dp, _dp]
    L.myqc_int1e_main.argtypes = [c_char_p]
    L.myqc_parse_zmat.argtypes = [c_char_p, c_int, ctypes.POINTER(c_int), _ip, _dp, _ip, ctypes.POINTER(c_int),
                                  ctypes.POINTER(c_int), ctypes.POINTER(c_int)]
    L.myqc_parse_main.argtypes = [c_char_p]
    L.myqc_ao2mo_transform.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                       c_void_p, c_void_p]
    L.myqc_ao2mo_transform_host.argtypes = [_dp, c_int, _dp, c_int, _dp, c_int, _dp, c_int, _dp, c_int, _dp]
    L.myqc_pack_dense.argtypes = [_dp, c_int, _dp]
    L.myqc_ao2mo_main.argtypes = [c_char_p]
    L.myqc_dmma_peak.argtypes = [c_int, _dp]
    L.myqc_ao2mo_workspace_bytes.argtypes = [c_int] * 5
    L.myqc_ao2mo_workspace_bytes.restype = ctypes.c_int64
    L.myqc_ao2mo_transform_ws.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                          c_void_p, c_void_p, ctypes.c_int64, c_void_p]
    L.myqc_ao2mo_flops.argtypes = [c_int] * 5
    L.myqc_ao2mo_flops.restype = ctypes.c_double
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("myqc_last_error", "myqc_eri_plan_out_offset", "myqc_eri_plan_out_elems",
                        "myqc_eri_plan_destroy", "myqc_eri_last_d2h_bytes", "myqc_ao2mo_flops", "myqc_fock_mask_words", "myqc_ao2mo_workspace_bytes",
                        "myqc_eri_release_cache"):
            fn.restype = c_int
    _lib = L
    return L


def _check(rc: int):
    if rc != MYQC_OK:
        raise MyQCError(rc, lib().myqc_last_error().decode())


def device_count() -> int:
    return lib().myqc_device_count()


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


@dataclass
class System:
    """The arrays the reference passes to proc2e (int2e.f90:66,78-112), 0-based contents."""
    nnuc: int
    atoms: np.ndarray    # int32 [nnuc]
    xyz: np.ndarray      # float64 [3*nnuc], Fortran xyz(0:nnuc-1,0:2): xyz[i + nnuc*c]
    set: np.ndarray      # float64
    setinfo: np.ndarray  # int32
    bas: np.ndarray      # float64
    basinfo: np.ndarray  # int32
    ftab: np.ndarray     # float64 [2783]
    options: np.ndarray | None = None
    maxN: int = 2
    maxL: int = 1

    @property
    def nset(self) -> int:
        return int(self.setinfo[0])

    @property
    def setl(self) -> int:
        return int(self.setinfo[1])

    @property
    def ops(self) -> int:
        return int(self.basinfo[0])

    @property
    def norb(self) -> int:
        return int(self.basinfo[1])

    @property
    def npair(self) -> int:
        return self.norb * (self.norb + 1) // 2

    @property
    def nunique(self) -> int:
        return self.npair * (self.npair + 1) // 2

    def _common(self):
        return (self.nnuc, _d(self.xyz), self.nset, self.setl, _d(self.set), _i(self.setinfo),
                self.ops, _d(self.bas), _i(self.basinfo), _d(self.ftab))


# ---- file layer ---------------------------------------------------------------------------
def read_env(workdir: str):
    """getenv (env.f90:16-73): returns (atoms, xyz_fortran, nelcA, nelcB, fmem, options)."""
    L = lib()
    n, na, nb, no = (ctypes.c_int() for _ in range(4))
    fmem = ctypes.c_double()
    _check(L.myqc_read_env(workdir.encode(), 0, 0, n, na, nb, None, None, fmem, no, None))
    atoms = np.zeros(n.value, dtype=np.int32)
    xyz = np.zeros(3 * n.value)
    opts = np.zeros(max(no.value, 17), dtype=np.int32)
    _check(L.myqc_read_env(workdir.encode(), n.value, len(opts), n, na, nb, _i(atoms), _d(xyz), fmem, no, _i(opts)))
    return atoms, xyz, na.value, nb.value, fmem.value, opts[:no.value]


def build_basis(mybasis_path: str, atoms, bkey: int = 0, out_dir: str | None = None):
    """buildBasis (basis.f90:23-226): returns (set, setinfo, bas, basinfo, maxN, maxL)."""
    L = lib()
    atoms = np.ascontiguousarray(atoms, dtype=np.int32)
    nsc, noc, mN, mL = (ctypes.c_int() for _ in range(4))
    _check(L.myqc_build_basis(mybasis_path.encode(), bkey, len(atoms), _i(atoms), nsc, noc, None, None, None, None, mN, mL, None))
    ops, setl = 4, 7
    set_ = np.zeros(nsc.value)
    bas = np.zeros(nsc.value * ops)
    setinfo = np.zeros(2 + nsc.value * setl, dtype=np.int32)
    basinfo = np.zeros(2 + 5 * noc.value, dtype=np.int32)
    _check(L.myqc_build_basis(mybasis_path.encode(), bkey, len(atoms), _i(atoms), nsc, noc, _d(set_), _i(setinfo),
                              _d(bas), _i(basinfo), mN, mL, out_dir.encode() if out_dir else None))
    return set_, setinfo, bas, basinfo, mN.value, mL.value


def read_ftab(path: str) -> np.ndarray:
    ft = np.zeros(121 * 23)
    _check(lib().myqc_read_ftab(path.encode(), _d(ft)))
    return ft


def write_xx(path: str, xx: np.ndarray, norb: int, max_subrecord: int | None = None):
    flat = np.ascontiguousarray(np.asarray(xx).reshape(-1, order="F"))
    assert flat.size == norb ** 4
    if max_subrecord is None:
        _check(lib().myqc_write_xx(path.encode(), _d(flat), norb))
    else:
        _check(lib().myqc_write_xx_ex(path.encode(), _d(flat), norb, max_subrecord))


def read_xx(path: str, norb: int) -> np.ndarray:
    flat = np.zeros(norb ** 4)
    _check(lib().myqc_read_xx(path.encode(), _d(flat), norb))
    return flat.reshape((norb,) * 4, order="F")


def load_system(workdir: str, mybasis: str | None = None, ftab: str | None = None) -> System:
    """What PROGRAM int2e reads before calling proc2e (int2e.f90:50-55,161-163)."""
    atoms, xyz, _, _, _, opts = read_env(workdir)
    set_, setinfo, bas, basinfo, mN, mL = build_basis(mybasis or os.path.join(workdir, "mybasis"), atoms,
                                                      int(opts[2]) if len(opts) > 2 else 0)
    ft = read_ftab(ftab or os.path.join(workdir, "Ftab"))
    return System(nnuc=len(atoms), atoms=atoms, xyz=xyz, set=set_, setinfo=setinfo, bas=bas, basinfo=basinfo,
                  ftab=ft, options=opts, maxN=mN, maxL=mL)


def int2e_main(workdir: str, ngpu: int = 1) -> int:
    """PROGRAM int2e (int2e.f90:14-69) in `workdir`; returns the library status (0 = ok)."""
    return lib().myqc_int2e_main(workdir.encode(), ngpu)


# ---- compute ------------------------------------------------------------------------------
def eri_packed(s: System, ngpu: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    """Packed 8-fold-unique ERIs into a host array (layout: include/myqc_eri.h)."""
    if out is None:
        out = np.empty(s.nunique)
    assert out.size == s.nunique and out.dtype == np.float64 and out.flags.c_contiguous
    _check(lib().myqc_eri_packed(*s._common(), _d(out), ngpu))
    return out


_last_h2d_bytes = 0


def eri_packed_shard(s: System, out: np.ndarray, device: int = 0, shard: int = 0, nshards: int = 1) -> np.ndarray:
    """One shard of the packed array into the host array `out` (length from shard_layout)."""
    global _last_h2d_bytes
    assert out.dtype == np.float64 and out.flags.c_contiguous
    h2d = ctypes.c_int64()
    _check(lib().myqc_eri_packed_shard(*s._common(), _d(out), device, shard, nshards, ctypes.byref(h2d)))
    _last_h2d_bytes = h2d.value
    return out


def last_d2h_bytes() -> int:
    """Bytes that crossed the device -> host link in the last eri_packed_shard call of this thread."""
    return int(lib().myqc_eri_last_d2h_bytes())


def release_cache():
    """Drop the plan / device slice the one-shot calls keep per device."""
    lib().myqc_eri_release_cache()


def plan_h2d_bytes(s: System) -> int:
    """Bytes of pair/Boys tables the last eri_packed_shard call uploaded."""
    return _last_h2d_bytes


def eri_dense(s: System, ngpu: int = 1) -> np.ndarray:
    """Dense XX(i,j,g,h) with all 8 images filled: the array proc2e writes (int2e.f90:290-307)."""
    n = s.norb
    flat = np.empty(n ** 4)
    _check(lib().myqc_eri_dense(*s._common(), _d(flat), ngpu))
    return flat.reshape((n, n, n, n), order="F")


def proc2e(bas, basinfo, atoms, options, fmem, nnuc, xyz, set, setinfo, maxL, ftab, ngpu: int = 1):
    """Argument-for-argument mirror of the reference's
    `proc2e(bas,basinfo,atoms,options,fmem,nnuc,xyz,set,setinfo,maxL)` (int2e.f90:78); `ftab` is the
    table the reference reads from the `Ftab` file inside the routine.  Returns XX."""
    s = System(nnuc=nnuc, atoms=np.asarray(atoms, dtype=np.int32), xyz=np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1),
               set=np.ascontiguousarray(set, dtype=np.float64), setinfo=np.ascontiguousarray(setinfo, dtype=np.int32),
               bas=np.ascontiguousarray(bas, dtype=np.float64), basinfo=np.ascontiguousarray(basinfo, dtype=np.int32),
               ftab=np.ascontiguousarray(ftab, dtype=np.float64), options=options, maxL=maxL)
    return eri_dense(s, ngpu)


class Plan:
    """Device-resident execution plan for one shard on one GPU (myqc_eri_plan_*)."""

    def __init__(self, s: System, device: int = 0, shard: int = 0, nshards: int = 1):
        self._h = ctypes.c_void_p()
        self._s = s
        _check(lib().myqc_eri_plan_create(*s._common(), device, shard, nshards, ctypes.byref(self._h)))
        self.device = device
        self.out_offset = lib().myqc_eri_plan_out_offset(self._h)
        self.out_elems = lib().myqc_eri_plan_out_elems(self._h)

    def execute(self, d_out_ptr: int, stream: int = 0):
        """d_out_ptr: device pointer (int) to out_elems doubles; stream: cudaStream_t as int."""
        _check(lib().myqc_eri_plan_execute(self._h, ctypes.c_void_p(d_out_ptr), ctypes.c_void_p(stream)))

    def launches(self):
        """[(class id or -1 for the zero fill, tri flag, rows)] in launch order."""
        out = []
        for k in range(lib().myqc_eri_plan_launch_count(self._h)):
            c, t, r = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
            _check(lib().myqc_eri_plan_launch_info(self._h, k, ctypes.byref(c), ctypes.byref(t), ctypes.byref(r)))
            out.append((c.value, t.value, r.value))
        return out

    def execute_timed(self, d_out_ptr: int, stream: int = 0):
        """Like execute(), but synchronises and returns per-launch milliseconds (CUDA events)."""
        n = lib().myqc_eri_plan_launch_count(self._h)
        ms = (ctypes.c_float * n)()
        _check(lib().myqc_eri_plan_execute_timed(self._h, ctypes.c_void_p(d_out_ptr), ctypes.c_void_p(stream), ms))
        return list(ms)

    def executed_quartets(self):
        """After an execute: (primitive quartets evaluated per class, Schwarz threshold in use)."""
        nq = (ctypes.c_int64 * 6)()
        tau = ctypes.c_double()
        _check(lib().myqc_eri_plan_executed_quartets(self._h, nq, ctypes.byref(tau)))
        return list(nq), tau.value

    def stats(self):
        nq = (ctypes.c_int64 * 6)()
        fl = ctypes.c_double()
        nl = ctypes.c_int()
        _check(lib().myqc_eri_plan_stats(self._h, nq, ctypes.byref(fl), ctypes.byref(nl)))
        return {"nquartets": list(nq), "model_flops": fl.value, "nlaunch": nl.value}

    def close(self):
        if self._h:
            lib().myqc_eri_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


CLASS_NAMES = ["{0,0}", "{0,1}", "{0,2}", "{1,1}", "{1,2}", "{2,2}"]
CLASS_W = [60.0, 99.0, 228.0, 228.0, 693.0, 2691.0]  # model flop per canonical primitive quartet
# nominal dense FP64 (non-tensor DFMA) peak of a B200: 148 SMs x 64 DFMA/clk x 2 flop x 1.965 GHz
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12


def shard_layout(s: System, nshards: int) -> np.ndarray:
    """Host-only: packed-array offsets [nshards+1] of the shard slices."""
    off = np.zeros(nshards + 1, dtype=np.int64)
    _check(lib().myqc_eri_shard_layout(*s._common()[:-1], nshards, off.ctypes.data_as(_i64p)))
    return off


def host_zero(a: np.ndarray) -> None:
    """Host only: zero a contiguous float64 array (or slice) with the streaming-store routine of the sparse route."""
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    lib().myqc_host_zero(a.ctypes.data_as(_dp), a.size)


def shard_model(s: System, nshards: int):
    """Host-only: (class_seconds[nshards], fill_seconds[nshards]) the shard-cut cost model expects."""
    cs, fs = np.zeros(nshards), np.zeros(nshards)
    _check(lib().myqc_eri_shard_model(*s._common()[:-1], nshards, _d(cs), _d(fs)))
    return cs, fs


def canonical_stats(s: System):
    """Host-only: (nquartets[6], model_flops) of the whole molecule."""
    nq = np.zeros(6, dtype=np.int64)
    fl = ctypes.c_double()
    _check(lib().myqc_eri_canonical_stats(s.nnuc, _d(s.xyz), s.nset, s.setl, _d(s.set), _i(s.setinfo),
                                          nq.ctypes.data_as(_i64p), ctypes.byref(fl)))
    return nq, fl.value


def fp64_peak(device: int = 0) -> float:
    """Measured DFMA peak of `device` in TFLOP/s (register-resident microbenchmark)."""
    v = ctypes.c_double()
    _check(lib().myqc_fp64_peak(device, ctypes.byref(v)))
    return v.value


def pair_index(i: int, j: int, norb: int) -> int:
    return i * norb - i * (i - 1) // 2 + (j - i)


# ---- job helpers ----------------------------------------------------------------------------
def make_job(workdir: str, zmat_text: str, inputs_dir: str) -> System:
    """Create a myQC job directory (ZMAT + mybasis + Ftab), run the `parse` stage
    (myqc_parse_main, include/myqc_parse.h) and load what int2e would read."""
    import shutil

    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "ZMAT"), "w") as f:
        f.write(zmat_text)
    for name in ("mybasis", "Ftab"):
        shutil.copyfile(os.path.join(inputs_dir, name), os.path.join(workdir, name))
    rc = parse_main(workdir)
    if rc != MYQC_OK:
        raise MyQCError(rc, lib().myqc_last_error().decode())
    return load_system(workdir)


# ----------------------------------------------------------------------------------------------
# G(D) from the packed array (include/myqc_fock.h; RHFI2G.f90:72-95, UHFI2G.f90:71-99)
# ----------------------------------------------------------------------------------------------
def fock_rhf(packed: np.ndarray, norb: int, da: np.ndarray) -> np.ndarray:
    """Guv of RHFI2G.f90:80-90 from the packed unique ERIs (host arrays in, host array out)."""
    packed = np.ascontiguousarray(packed, dtype=np.float64)
    da = np.asfortranarray(da, dtype=np.float64)
    g = np.zeros((norb, norb), dtype=np.float64, order="F")
    _check(lib().myqc_fock_rhf_host(_d(packed), norb, _d(da), _d(g)))
    return g


def fock_uhf(packed: np.ndarray, norb: int, da: np.ndarray, db: np.ndarray):
    """(GuvA, GuvB) of UHFI2G.f90:80-93 from the packed unique ERIs."""
    packed = np.ascontiguousarray(packed, dtype=np.float64)
    da = np.asfortranarray(da, dtype=np.float64)
    db = np.asfortranarray(db, dtype=np.float64)
    ga = np.zeros((norb, norb), dtype=np.float64, order="F")
    gb = np.zeros((norb, norb), dtype=np.float64, order="F")
    _check(lib().myqc_fock_uhf_host(_d(packed), norb, _d(da), _d(db), _d(ga), _d(gb)))
    return ga, gb


def fock_rhf_device(d_packed: int, out_offset: int, out_elems: int, norb: int, d_da: int, d_g: int, stream: int = 0):
    """Device-pointer form (stream ordered): partial G of a slice of whole packed rows."""
    _check(lib().myqc_fock_rhf(d_packed, out_offset, out_elems, norb, d_da, d_g, stream))


def fock_uhf_device(d_packed: int, out_offset: int, out_elems: int, norb: int, d_da: int, d_db: int,
                    d_ga: int, d_gb: int, stream: int = 0):
    _check(lib().myqc_fock_uhf(d_packed, out_offset, out_elems, norb, d_da, d_db, d_ga, d_gb, stream))


def fock_mask_words(norb: int) -> int:
    """uint32 words of the sparsity mask of a norb-function packed array (myqc_fock.h)."""
    return int(lib().myqc_fock_mask_words(norb))


def fock_mask_build(d_packed: int, out_offset: int, out_elems: int, norb: int, d_mask: int, stream: int = 0):
    """One streaming pass over the slice: bit k of row P set iff (P | k, .) holds a nonzero integral."""
    _check(lib().myqc_fock_mask_build(d_packed, out_offset, out_elems, norb, d_mask, stream))


def fock_rhf_masked_device(d_packed: int, out_offset: int, out_elems: int, norb: int, d_da: int, d_mask: int, d_g: int,
                           stream: int = 0):
    _check(lib().myqc_fock_rhf_masked(d_packed, out_offset, out_elems, norb, d_da, d_mask, d_g, stream))


def fock_uhf_masked_device(d_packed: int, out_offset: int, out_elems: int, norb: int, d_da: int, d_db: int, d_mask: int,
                           d_ga: int, d_gb: int, stream: int = 0):
    _check(lib().myqc_fock_uhf_masked(d_packed, out_offset, out_elems, norb, d_da, d_db, d_mask, d_ga, d_gb, stream))


def _read_records(path: str):
    """Fortran unformatted sequential records (4-byte little-endian markers) as float64 arrays."""
    raw = open(path, "rb").read()
    out, pos = [], 0
    while pos < len(raw):
        n = int(np.frombuffer(raw, dtype="<i4", count=1, offset=pos)[0])
        out.append(np.frombuffer(raw, dtype="<f8", count=n // 8, offset=pos + 4).copy())
        pos += 8 + n
    return out


def _write_records(path: str, arrays):
    with open(path, "wb") as f:
        for a in arrays:
            b = np.asfortranarray(a, dtype="<f8").tobytes(order="F")
            f.write(np.int32(len(b)).tobytes()); f.write(b); f.write(np.int32(len(b)).tobytes())


def pack_dense(xx: np.ndarray) -> np.ndarray:
    """Packed 8-fold-unique array from a dense XX(n,n,n,n) (the inverse of eri_dense's expansion)."""
    n = xx.shape[0]
    ii, jj = np.triu_indices(n)
    m = xx[ii, jj][:, ii, jj]                 # (npair, npair), row P, column P'
    r, c = np.triu_indices(len(ii))
    return np.ascontiguousarray(m[r, c])


def rhf_i2g(workdir: str) -> np.ndarray:
    """PROGRAM RHFI2G (src/I2G/RHFI2G.f90): reads `XX` and `Da`, writes `Guv` in the job directory."""
    norb = int(open(os.path.join(workdir, "basinfo")).read().split()[1])
    xx = read_xx(os.path.join(workdir, "XX"), norb)
    da = _read_records(os.path.join(workdir, "Da"))[0].reshape((norb, norb), order="F")
    g = fock_rhf(pack_dense(xx), norb, da)
    _write_records(os.path.join(workdir, "Guv"), [g])
    return g


def uhf_i2g(workdir: str):
    """PROGRAM UHFI2G (src/I2G/UHFI2G.f90): reads `XX`, `Da` (two records), writes `Guv` (two records)."""
    norb = int(open(os.path.join(workdir, "basinfo")).read().split()[1])
    xx = read_xx(os.path.join(workdir, "XX"), norb)
    recs = _read_records(os.path.join(workdir, "Da"))
    da = recs[0].reshape((norb, norb), order="F")
    db = recs[1].reshape((norb, norb), order="F")
    ga, gb = fock_uhf(pack_dense(xx), norb, da, db)
    _write_records(os.path.join(workdir, "Guv"), [ga, gb])
    return ga, gb


# ----------------------------------------------------------------------------------------------
# one-electron integrals (include/myqc_int1e.h; int1e.f90)
# ----------------------------------------------------------------------------------------------
def int1e(s: System):
    """(Suv, Huv) of proc1e (int1e.f90:132-280): overlap and core Hamiltonian, norb x norb."""
    n = s.norb
    S = np.zeros((n, n), order="F")
    H = np.zeros((n, n), order="F")
    atoms = np.ascontiguousarray(s.atoms, dtype=np.int32)
    _check(lib().myqc_int1e(s.nnuc, _d(s.xyz), _i(atoms), s.nset, s.setl, _d(s.set), _i(s.setinfo), s.ops,
                            _d(s.bas), _i(s.basinfo), _d(s.ftab), _d(S), _d(H)))
    return S, H


def parse_zmat(zmat_text: str):
    """myqc_parse_zmat (PROGRAM parser's arithmetic, parser.f90): returns
    (atoms int32[n], xyz float64[n,3] bohr, options int32[17], nelcA, nelcB, problems bit mask)."""
    L = lib()
    n, na, nb, pr = (ctypes.c_int() for _ in range(4))
    opts = np.zeros(17, dtype=np.int32)
    _check(L.myqc_parse_zmat(zmat_text.encode(), 0, n, None, None, _i(opts), na, nb, pr))
    atoms = np.zeros(n.value, dtype=np.int32)
    xyz = np.zeros(3 * n.value)
    _check(L.myqc_parse_zmat(zmat_text.encode(), n.value, n, _i(atoms), _d(xyz), _i(opts), na, nb, pr))
    return atoms, xyz.reshape(-1, 3), opts, na.value, nb.value, pr.value


def parse_main(workdir: str) -> int:
    """PROGRAM parser (parser.f90:22-108) in `workdir`: ZMAT -> nucpos, envdat, fmem; returns the status."""
    import sys
    sys.stdout.flush()
    return lib().myqc_parse_main(workdir.encode())


def int1e_main(workdir: str) -> int:
    """PROGRAM int1e (int1e.f90:14-131) in `workdir`; returns the library status (0 = ok)."""
    return lib().myqc_int1e_main(workdir.encode())


def read_matrix_text(path: str, norb: int) -> np.ndarray:
    """READ(u,*) M(:,:) of a list-directed text file (Suv / Huv, scf.f90:140-144)."""
    toks = open(path).read().replace(",", " ").replace("D", "E").split()
    return np.array([float(t) for t in toks[:norb * norb]]).reshape((norb, norb), order="F")


# ----------------------------------------------------------------------------------------------
# AO -> MO transformation (include/myqc_ao2mo.h; ao2mo.f90)
# ----------------------------------------------------------------------------------------------
def _col_block(c: np.ndarray, norb: int) -> np.ndarray:
    c = np.asfortranarray(c, dtype=np.float64)
    assert c.ndim == 2 and c.shape[0] == norb, "coefficient blocks are norb x n_k (the reference's Cm(0:ntot-1, cols))"
    return c


def ao2mo_transform(packed: np.ndarray, norb: int, c1, c2, c3, c4) -> np.ndarray:
    """O(p,q,r,s) = sum_uvld C1(u,p) C2(v,q) C3(l,r) C4(d,s) (uv|ld): idx1_trans..idx4_trans of
    ao2mo.f90:1306-1439 composed, from the packed unique ERIs.  Returns O[p,q,r,s] (Fortran order)."""
    packed = np.ascontiguousarray(packed, dtype=np.float64)
    cs = [_col_block(c, norb) for c in (c1, c2, c3, c4)]
    dims = tuple(c.shape[1] for c in cs)
    out = np.zeros(int(np.prod(dims)))
    _check(lib().myqc_ao2mo_transform_host(_d(packed), norb, _d(cs[0]), dims[0], _d(cs[1]), dims[1],
                                           _d(cs[2]), dims[2], _d(cs[3]), dims[3], _d(out)))
    return out.reshape(dims, order="F")


def ao2mo_transform_device(d_packed: int, norb: int, d_c1: int, n1: int, d_c2: int, n2: int, d_c3: int, n3: int,
                           d_c4: int, n4: int, d_out: int, stream: int = 0):
    """Device-pointer form (stream ordered, no synchronisation)."""
    _check(lib().myqc_ao2mo_transform(d_packed, norb, d_c1, n1, d_c2, n2, d_c3, n3, d_c4, n4, d_out, stream))


def dmma_peak(device: int = 0) -> float:
    """Measured FP64 tensor-pipe (DMMA m8n8k4) peak of `device` in TFLOP/s."""
    v = ctypes.c_double()
    _check(lib().myqc_dmma_peak(device, ctypes.byref(v)))
    return v.value


def ao2mo_workspace_bytes(norb: int, n1: int, n2: int, n3: int, n4: int) -> int:
    return int(lib().myqc_ao2mo_workspace_bytes(norb, n1, n2, n3, n4))


def ao2mo_transform_ws(d_packed: int, norb: int, d_c1: int, n1: int, d_c2: int, n2: int, d_c3: int, n3: int,
                       d_c4: int, n4: int, d_out: int, d_ws: int, ws_bytes: int, stream: int = 0):
    """Device-pointer form with caller-provided scratch: no allocation, no synchronisation."""
    _check(lib().myqc_ao2mo_transform_ws(d_packed, norb, d_c1, n1, d_c2, n2, d_c3, n3, d_c4, n4, d_out, d_ws, ws_bytes, stream))


def ao2mo_flops(norb: int, n1: int, n2: int, n3: int, n4: int) -> float:
    return lib().myqc_ao2mo_flops(norb, n1, n2, n3, n4)


def pack_dense_c(xx: np.ndarray) -> np.ndarray:
    """myqc_pack_dense: the host helper the `ao2mo` program uses on the XX record it reads."""
    n = xx.shape[0]
    flat = np.ascontiguousarray(np.asarray(xx).reshape(-1, order="F"))
    npair = n * (n + 1) // 2
    out = np.zeros(npair * (npair + 1) // 2)
    _check(lib().myqc_pack_dense(_d(flat), n, _d(out)))
    return out


def ao2mo_main(workdir: str) -> int:
    """PROGRAM ao2mo (ao2mo.f90:22-98) in `workdir`; returns the library status (0 = ok)."""
    return lib().myqc_ao2mo_main(workdir.encode())


def write_matrix_text(path: str, mats):
    """WRITE(u,*) M(:,:) for each matrix (the `Cui` / `eig` files scf.f90 leaves for ao2mo / mp2)."""
    with open(path, "w") as f:
        for m in mats:
            v = np.asarray(m, dtype=np.float64).reshape(-1, order="F")
            for k in range(0, len(v), 3):
                f.write("".join("  %24.16E" % x for x in v[k:k + 3]) + "\n")
