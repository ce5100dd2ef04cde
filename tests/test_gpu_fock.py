"""Parity of the device G(D) build (include/myqc_fock.h) with the reference's
RHFI2G / UHFI2G loops, restated literally with numpy on the oracle's dense XX.  Needs a B200."""
import os

import numpy as np
import pytest

import myqc_b200 as Q
from conftest import INPUTS, oracle_system, product_system
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1.0e-10  # absolute, Hartree (same bar as the integrals: G is a sum of O(n^2) of them times D)


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if Q.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device; the Fock build has no CPU fallback")


def ref_rhf(xx, d):
    """RHFI2G.f90:80-90, literally."""
    return (np.einsum("kl,ijkl->ij", d, xx) - 0.25 * np.einsum("kl,ikjl->ij", d, xx)
            - 0.25 * np.einsum("kl,iljk->ij", d, xx))


def ref_uhf(xx, da, db):
    """UHFI2G.f90:80-93, literally."""
    j = np.einsum("kl,ijkl->ij", da + db, xx)
    return j - np.einsum("kl,ikjl->ij", da, xx), j - np.einsum("kl,ikjl->ij", db, xx)


def sym_density(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((n, max(1, n // 2)))
    return scale * (c @ c.T) / n


@pytest.mark.parametrize("name", ["H2", "HF", "CO2", "NO", "h2o_4", "h2o_8"])
def test_rhf_g_matches_reference_loops(name, tmp_path, oracle_inputs):
    s = product_system(name, tmp_path)
    packed = Q.eri_packed(s)
    mol, b, ft = oracle_system(name, oracle_inputs)
    xx, _ = O.int2e_dense(mol, b, ft)
    xx = np.array(xx)
    d = sym_density(s.norb, 11)
    g = Q.fock_rhf(packed, s.norb, d)
    ref = ref_rhf(xx, d)
    assert np.abs(g - ref).max() < TOL * max(1.0, np.abs(ref).max())
    assert np.abs(g - g.T).max() < 1e-12


@pytest.mark.parametrize("name", ["OH", "NO", "h2o_4"])
def test_uhf_g_matches_reference_loops(name, tmp_path, oracle_inputs):
    s = product_system(name, tmp_path)
    packed = Q.eri_packed(s)
    mol, b, ft = oracle_system(name, oracle_inputs)
    xx = np.array(O.int2e_dense(mol, b, ft)[0])
    da, db = sym_density(s.norb, 3), sym_density(s.norb, 4, 0.7)
    ga, gb = Q.fock_uhf(packed, s.norb, da, db)
    ra, rb = ref_uhf(xx, da, db)
    scale = max(1.0, np.abs(ra).max(), np.abs(rb).max())
    assert np.abs(ga - ra).max() < TOL * scale and np.abs(gb - rb).max() < TOL * scale


def test_rhf_nonsymmetric_density_uses_the_symmetric_part(tmp_path, oracle_inputs):
    """The reference's RHF formula only depends on (D + D^T)/2; so does the device build."""
    s = product_system("CO", tmp_path)
    packed = Q.eri_packed(s)
    mol, b, ft = oracle_system("CO", oracle_inputs)
    xx = np.array(O.int2e_dense(mol, b, ft)[0])
    d = np.random.default_rng(5).standard_normal((s.norb, s.norb))
    assert np.abs(Q.fock_rhf(packed, s.norb, d) - ref_rhf(xx, d)).max() < 1e-9


def test_scf_energy_through_the_device_g_build(tmp_path, oracle_inputs):
    """RHF fixed-point iteration of scf.f90:720-904 with G(D) from the device build every iteration:
    CO2 energy of BASELINE.md section 2 to 1e-9 Eh."""
    from scipy.linalg import eigh
    s = product_system("CO2", tmp_path)
    packed = Q.eri_packed(s)
    mol, b, ft = oracle_system("CO2", oracle_inputs)
    S, H = O.int1e(mol, b, ft)
    nA, nB = O.electrons(mol)
    nocc = (nA + nB) // 2
    enr = O.nuclear_repulsion(mol)
    _, C = eigh(H, S)
    D = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    for it in range(500):
        F = H + Q.fock_rhf(packed, s.norb, D)
        e_tot = 0.5 * np.sum(D * (F + H)) + enr
        _, C = eigh(F, S)
        Dn = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        done = it > 0 and np.max(np.abs(Dn - D)) < 1e-11
        D = Dn
        if done:
            break
    F = H + Q.fock_rhf(packed, s.norb, D)
    e_tot = 0.5 * np.sum(D * (F + H)) + enr
    assert abs(e_tot - (-183.32315970625)) < 1e-9


def test_partial_g_of_row_slices_add_up(tmp_path):
    """Multi-GPU form on one device: the partial G of every shard's slice sums to G (the only
    collective a sharded SCF needs is one all-reduce of norb^2 doubles)."""
    import torch
    s = product_system("h2o_8", tmp_path)
    n = s.norb
    packed = torch.from_numpy(Q.eri_packed(s)).cuda()
    d = torch.from_numpy(np.asfortranarray(sym_density(n, 2)).ravel(order="F").copy()).cuda()
    full = torch.empty(n * n, dtype=torch.float64, device="cuda")
    Q.fock_rhf_device(packed.data_ptr(), 0, packed.numel(), n, d.data_ptr(), full.data_ptr())
    for nsh in (2, 5):
        off = Q.shard_layout(s, nsh)
        acc = torch.zeros_like(full)
        part = torch.empty_like(full)
        for k in range(nsh):
            lo, hi = int(off[k]), int(off[k + 1])
            Q.fock_rhf_device(packed.data_ptr() + 8 * lo, lo, hi - lo, n, d.data_ptr(), part.data_ptr())
            acc += part
        torch.cuda.synchronize()
        assert float((acc - full).abs().max()) < 1e-12
    # a slice that is not made of whole rows is refused
    with pytest.raises(Q.MyQCError):
        Q.fock_rhf_device(packed.data_ptr() + 8, 1, 10, n, d.data_ptr(), full.data_ptr())


def test_linearity_and_symmetry_at_112_functions(tmp_path):
    """(H2O)_16: G is linear in D and symmetric (size-independent properties)."""
    s = product_system("h2o_16", tmp_path)
    packed = Q.eri_packed(s)
    n = s.norb
    d1, d2 = sym_density(n, 8), sym_density(n, 9)
    g1, g2 = Q.fock_rhf(packed, n, d1), Q.fock_rhf(packed, n, d2)
    g12 = Q.fock_rhf(packed, n, 0.3 * d1 - 1.7 * d2)
    assert np.abs(g12 - (0.3 * g1 - 1.7 * g2)).max() < 1e-10
    assert np.abs(g1 - g1.T).max() < 1e-12
    ga, gb = Q.fock_uhf(packed, n, 0.5 * d1, 0.5 * d1)
    # closed shell as UHF: Ga = Gb = G_RHF(D) with Da = Db = D/2
    assert np.abs(ga - g1).max() < 1e-10 and np.abs(gb - g1).max() < 1e-10


def test_rhfi2g_file_mirror(tmp_path, oracle_inputs):
    """PROGRAM RHFI2G at the process boundary: XX + Da in, Guv out (same record framing)."""
    s = product_system("HF", tmp_path)
    assert Q.int2e_main(str(tmp_path), 1) == 0
    n = s.norb
    d = sym_density(n, 21)
    Q._write_records(os.path.join(tmp_path, "Da"), [d])
    g = Q.rhf_i2g(str(tmp_path))
    raw = open(os.path.join(tmp_path, "Guv"), "rb").read()
    assert len(raw) == 8 * n * n + 8
    guv = np.frombuffer(raw[4:-4], dtype="<f8").reshape((n, n), order="F")
    mol, b, ft = oracle_system("HF", oracle_inputs)
    xx = np.array(O.int2e_dense(mol, b, ft)[0])
    assert np.abs(guv - ref_rhf(xx, d)).max() < TOL and np.array_equal(guv, g)


@pytest.mark.parametrize("name", ["CO2", "h2o_8", "c4h10"])
def test_masked_builds_are_identical_to_unmasked(name, tmp_path):
    """The sparsity mask (myqc_fock_mask_build) only skips all-zero (row, k) blocks: RHF and UHF results
    equal the unmasked builds; every set bit has a nonzero behind it and every clear bit has none."""
    import torch
    s = product_system(name, tmp_path)
    n = s.norb
    packed_h = Q.eri_packed(s)
    packed = torch.from_numpy(packed_h).cuda()
    d1 = torch.from_numpy(np.asfortranarray(sym_density(n, 3)).ravel(order="F").copy()).cuda()
    d2 = torch.from_numpy(np.asfortranarray(sym_density(n, 4)).ravel(order="F").copy()).cuda()
    mask = torch.full((Q.fock_mask_words(n),), -1, dtype=torch.int32, device="cuda")
    Q.fock_mask_build(packed.data_ptr(), 0, packed.numel(), n, mask.data_ptr())
    g0, g1 = torch.empty(n * n, dtype=torch.float64, device="cuda"), torch.empty(n * n, dtype=torch.float64, device="cuda")
    Q.fock_rhf_device(packed.data_ptr(), 0, packed.numel(), n, d1.data_ptr(), g0.data_ptr())
    Q.fock_rhf_masked_device(packed.data_ptr(), 0, packed.numel(), n, d1.data_ptr(), mask.data_ptr(), g1.data_ptr())
    torch.cuda.synchronize()
    assert float((g0 - g1).abs().max()) < 1e-12
    ga0, gb0, ga1, gb1 = (torch.empty(n * n, dtype=torch.float64, device="cuda") for _ in range(4))
    Q.fock_uhf_device(packed.data_ptr(), 0, packed.numel(), n, d1.data_ptr(), d2.data_ptr(), ga0.data_ptr(), gb0.data_ptr())
    Q.fock_uhf_masked_device(packed.data_ptr(), 0, packed.numel(), n, d1.data_ptr(), d2.data_ptr(), mask.data_ptr(),
                             ga1.data_ptr(), gb1.data_ptr())
    torch.cuda.synchronize()
    assert float((ga0 - ga1).abs().max()) < 1e-12 and float((gb0 - gb1).abs().max()) < 1e-12
    # the mask itself against numpy
    mw = (n + 31) // 32
    m = mask.cpu().numpy().view(np.uint32).reshape(-1, mw)
    npair = n * (n + 1) // 2
    kk = np.repeat(np.arange(n), np.arange(n, 0, -1))
    pos = 0
    for P in range(npair):
        row = packed_h[pos:pos + npair - P] != 0
        pos += npair - P
        cnt = np.bincount(kk[P:], weights=row, minlength=n) > 0
        bits = np.array([(m[P, k >> 5] >> (k & 31)) & 1 for k in range(n)], dtype=bool)
        assert np.array_equal(bits, cnt), P
