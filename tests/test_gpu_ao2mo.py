"""GPU parity tests of the AO->MO transformation (include/myqc_ao2mo.h) against
the numpy restatement of ao2mo.f90's idx1_trans..idx4_trans loop nests (oracle.ao2mo_idx_trans)
on the oracle's dense XX.  Tolerance: 1e-10 absolute per transformed integral."""
import os

import numpy as np
import pytest

import myqc_b200 as Q
from oracle import oracle as O

from conftest import oracle_system, product_system

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if Q.device_count() < 1:
        pytest.fail("no CUDA device: ao2mo has no CPU fallback")


def blocks(n, dims, seed):
    rng = np.random.default_rng(seed)
    return [np.asfortranarray(rng.standard_normal((n, k))) for k in dims]


def oracle_xx(name, oracle_inputs):
    mol, b, ft = oracle_system(name, oracle_inputs)
    return mol, b, ft, np.array(O.int2e_dense(mol, b, ft)[0])


@pytest.mark.parametrize("name,dims", [("H2", (1, 1, 1, 1)), ("HF", (5, 1, 5, 1)), ("CO2", (11, 4, 11, 4)),
                                       ("CO2", (4, 11, 3, 15)), ("NO", (8, 2, 7, 3)), ("h2o_4", (20, 8, 20, 8)),
                                       ("h2o_8", (40, 16, 13, 56))])
def test_transform_matches_reference_loops(name, dims, tmp_path, oracle_inputs):
    """Random coefficient blocks of every shape the reference uses (occ/vrt, vrt/occ, ragged)."""
    s = product_system(name, tmp_path)
    packed = Q.eri_packed(s)
    _, _, _, xx = oracle_xx(name, oracle_inputs)
    c = blocks(s.norb, dims, 5)
    got = Q.ao2mo_transform(packed, s.norb, *c)
    ref = O.ao2mo_idx_trans(xx, *c)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < TOL * max(1.0, np.abs(ref).max())


def test_tensor_pipe_and_plain_dfma_tiles_agree(tmp_path, monkeypatch):
    """The DMMA fragment layout against the plain-DFMA instantiation of the same tiling."""
    s = product_system("h2o_4", tmp_path)
    packed = Q.eri_packed(s)
    c = blocks(s.norb, (20, 8, 9, 28), 11)
    a = Q.ao2mo_transform(packed, s.norb, *c)
    monkeypatch.setenv("MYQC_AO2MO_GEMM", "simt")
    b = Q.ao2mo_transform(packed, s.norb, *c)
    assert np.abs(a - b).max() < 1e-12 * max(1.0, np.abs(a).max())


def test_many_panels_and_column_chunks(tmp_path, oracle_inputs, monkeypatch):
    """A 1 MB scratch budget forces the multi-panel / multi-chunk path on a small molecule."""
    s = product_system("h2o_4", tmp_path)
    packed = Q.eri_packed(s)
    _, _, _, xx = oracle_xx("h2o_4", oracle_inputs)
    c = blocks(s.norb, (20, 8, 20, 8), 3)
    monkeypatch.setenv("MYQC_AO2MO_SCRATCH_MB", "1")
    got = Q.ao2mo_transform(packed, s.norb, *c)
    assert np.abs(got - O.ao2mo_idx_trans(xx, *c)).max() < TOL * 10


def test_caller_workspace_form_matches(tmp_path):
    """myqc_ao2mo_transform_ws (no allocation inside) == the allocating call; too small a workspace is refused."""
    import torch
    s = product_system("h2o_4", tmp_path)
    n = s.norb
    packed_h = Q.eri_packed(s)
    c = blocks(n, (20, 8, 20, 8), 9)
    ref = Q.ao2mo_transform(packed_h, n, *c)
    packed = torch.from_numpy(packed_h).cuda()
    dc = [torch.from_numpy(x.ravel(order="F").copy()).cuda() for x in c]
    out = torch.empty(ref.size, dtype=torch.float64, device="cuda")
    nbytes = Q.ao2mo_workspace_bytes(n, 20, 8, 20, 8)
    ws = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device="cuda")
    for _ in range(2):  # the buffer is reusable
        Q.ao2mo_transform_ws(packed.data_ptr(), n, dc[0].data_ptr(), 20, dc[1].data_ptr(), 8, dc[2].data_ptr(), 20,
                             dc[3].data_ptr(), 8, out.data_ptr(), ws.data_ptr(), nbytes)
        torch.cuda.synchronize()
        assert np.abs(out.cpu().numpy().reshape(ref.shape, order="F") - ref).max() < 1e-12
    with pytest.raises(Q.MyQCError):
        Q.ao2mo_transform_ws(packed.data_ptr(), n, dc[0].data_ptr(), 20, dc[1].data_ptr(), 8, dc[2].data_ptr(), 20,
                             dc[3].data_ptr(), 8, out.data_ptr(), ws.data_ptr(), nbytes - 8)


def _scf_files(name, tmp_path, oracle_inputs, uhf):
    """int2e through the product, SCF through the oracle: leaves XX, basinfo, Cui in the job directory."""
    s = product_system(name, tmp_path)
    assert Q.int2e_main(str(tmp_path), 1) == 0
    mol, b, ft, xx = oracle_xx(name, oracle_inputs)
    S, H = O.int1e(mol, b, ft)
    nA, nB = O.electrons(mol)
    enr = O.nuclear_repulsion(mol)
    if uhf:
        _, ea, eb, _, CA, CB = O.scf_uhf(S, H, xx, nA, nB, enr, orbitals=True)
        Q.write_matrix_text(os.path.join(tmp_path, "Cui"), [CA, CB])
    else:
        _, ea, _, CA = O.scf_rhf(S, H, xx, nA + nB, enr, orbitals=True)
        CB, eb = CA, ea
        Q.write_matrix_text(os.path.join(tmp_path, "Cui"), [CA])
    return s, xx, CA, CB, ea, eb, nA, nB


def _compare_files(tmp_path, ref):
    for name, recs in ref.items():
        got = Q._read_records(os.path.join(tmp_path, name))
        assert len(got) == len(recs), name
        for g, r in zip(got, recs):
            assert g.shape == r.shape and np.abs(g - r).max() < TOL, name


def test_ao2mo_program_mp2_rhf_files_and_energy(tmp_path, oracle_inputs):
    """PROGRAM ao2mo, CALC=MP2 REF=RHF (ao2mo.f90:465-602): ijab_AA / ijab_AB records, then mp2.f90's
    energy from those files against the CFOUR value shipped with the reference (examples/CO2/cfour/out:
    E2(TOT) = -0.089867321680; the reference's float32 pi moves it by 1e-6)."""
    s, xx, CA, CB, ea, eb, nA, nB = _scf_files("CO2", tmp_path, oracle_inputs, False)
    assert Q.ao2mo_main(str(tmp_path)) == 0 and not (tmp_path / "error").exists()
    _compare_files(tmp_path, O.ao2mo_files("mp2_rhf", xx, CA, CB, nA, nB))
    recs = Q._read_records(os.path.join(tmp_path, "ijab_AB"))
    e_aa, e_ab, e2 = O.mp2_rhf_energy(recs, ea, nA, s.norb - nA)
    assert abs(e2 - (-0.089867321680)) < 3e-6 and abs(e_aa - (-0.011001822459)) < 1e-6


def test_ao2mo_program_mp2_uhf_files(tmp_path, oracle_inputs):
    """CALC=MP2 REF=UHF (ao2mo.f90:614-904) on the reference's NO example: ijab_AA, ijab_BB, ijab_AB."""
    s, xx, CA, CB, ea, eb, nA, nB = _scf_files("NO", tmp_path, oracle_inputs, True)
    assert Q.ao2mo_main(str(tmp_path)) == 0 and not (tmp_path / "error").exists()
    _compare_files(tmp_path, O.ao2mo_files("mp2_uhf", xx, CA, CB, nA, nB))


def test_ao2mo_program_cis_uhf_files(tmp_path, oracle_inputs):
    """EXCITE=CIS REF=UHF (ao2mo.f90:919-1227) on the reference's OH example: the five ajib / ajbi files."""
    s, xx, CA, CB, ea, eb, nA, nB = _scf_files("OH", tmp_path, oracle_inputs, True)
    assert Q.ao2mo_main(str(tmp_path)) == 0 and not (tmp_path / "error").exists()
    _compare_files(tmp_path, O.ao2mo_files("cis_uhf", xx, CA, CB, nA, nB))


def test_ao2mo_program_rejects_what_the_reference_rejects(tmp_path, oracle_inputs):
    """CALC=SCF without EXCITE: 'that transform type has not been coded yet' + touch error (ao2mo.f90:88-92)."""
    product_system("HF", tmp_path)
    assert Q.int2e_main(str(tmp_path), 1) == 0
    assert Q.ao2mo_main(str(tmp_path)) != 0 and (tmp_path / "error").exists()


def test_symmetry_and_linearity_at_112_functions(tmp_path):
    """(H2O)_16: (pq|rs) = (rs|pq) with swapped blocks, linear in each block, and equal to a numpy
    transformation of the dense array rebuilt from the same packed integrals."""
    s = product_system("h2o_16", tmp_path)
    packed = Q.eri_packed(s)
    n = s.norb
    c1, c2, c3, c4 = blocks(n, (9, 5, 7, 6), 17)
    o = Q.ao2mo_transform(packed, n, c1, c2, c3, c4)
    o_sw = Q.ao2mo_transform(packed, n, c3, c4, c1, c2)
    assert np.abs(o - o_sw.transpose(2, 3, 0, 1)).max() < 1e-10
    o_q = Q.ao2mo_transform(packed, n, c2, c1, c3, c4)
    assert np.abs(o - o_q.transpose(1, 0, 2, 3)).max() < 1e-10
    d1 = blocks(n, (9,), 18)[0]
    o_lin = Q.ao2mo_transform(packed, n, 0.5 * c1 - 2.0 * d1, c2, c3, c4)
    o_d = Q.ao2mo_transform(packed, n, d1, c2, c3, c4)
    assert np.abs(o_lin - (0.5 * o - 2.0 * o_d)).max() < 1e-9
    xx = np.array(O.dense_from_packed(packed, n))
    assert np.abs(o - O.ao2mo_idx_trans(xx, c1, c2, c3, c4)).max() < 1e-10 * max(1.0, np.abs(o).max())
