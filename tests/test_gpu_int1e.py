"""Parity of the device one-electron integrals (include/myqc_int1e.h) with the oracle's
restatement of int1e.f90 (overlap :321-386, kinetic :391-474, coulomb :479-582).  Needs a B200."""
import os

import numpy as np
import pytest

import myqc_b200 as Q
from conftest import EXAMPLES, oracle_system, product_system
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1.0e-10  # Hartree, absolute, per matrix element


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if Q.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device; the integral engine has no CPU fallback")


@pytest.mark.parametrize("name", EXAMPLES + ["h2o_4", "c4h10", "h2o_16"])
def test_s_and_h_match_the_oracle(name, tmp_path, oracle_inputs):
    s = product_system(name, tmp_path)
    S, H = Q.int1e(s)
    mol, b, ft = oracle_system(name, oracle_inputs)
    Sr, Hr = O.int1e(mol, b, ft)
    assert np.abs(S - Sr).max() < TOL and np.abs(H - Hr).max() < TOL
    # both triangles come from the same ordered loop as in the reference
    assert np.abs(S - S.T).max() < 1e-12 and np.abs(H - H.T).max() < 1e-10
    assert np.abs(np.diag(S) - 1.0).max() < 1e-6  # normalised contracted functions


def test_int1e_program_mirror_and_scf_energy(tmp_path, oracle_inputs):
    """PROGRAM int1e + PROGRAM int2e in one job directory, then the SCF of scf.f90:720-904 on their
    files with G(D) from the device: CO2 energy of BASELINE.md to 1e-9 Eh -- the whole integral stage
    and the per-iteration consumer without a Fortran binary."""
    from scipy.linalg import eigh
    s = product_system("CO2", tmp_path)
    assert Q.int1e_main(str(tmp_path)) == 0 and not (tmp_path / "error").exists()
    assert Q.int2e_main(str(tmp_path), 1) == 0
    n = s.norb
    S = Q.read_matrix_text(os.path.join(tmp_path, "Suv"), n)
    H = Q.read_matrix_text(os.path.join(tmp_path, "Huv"), n)
    mol, b, ft = oracle_system("CO2", oracle_inputs)
    Sr, Hr = O.int1e(mol, b, ft)
    assert np.abs(S - Sr).max() < TOL and np.abs(H - Hr).max() < TOL
    assert abs(float(open(tmp_path / "fmem").read().split()[0]) - 1000.0) < 1e-6
    # second run: both files exist -> nothing recomputed, Sold / Hold touched (int1e.f90:98-111)
    before = open(tmp_path / "Suv").read()
    assert Q.int1e_main(str(tmp_path)) == 0
    assert (tmp_path / "Sold").exists() and (tmp_path / "Hold").exists() and open(tmp_path / "Suv").read() == before
    packed = Q.pack_dense(Q.read_xx(os.path.join(tmp_path, "XX"), n))
    nA, nB = O.electrons(mol)
    nocc = (nA + nB) // 2
    enr = O.nuclear_repulsion(mol)
    _, C = eigh(H, S)
    D = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    for it in range(500):
        F = H + Q.fock_rhf(packed, n, D)
        _, C = eigh(F, S)
        Dn = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        done = it > 0 and np.max(np.abs(Dn - D)) < 1e-11
        D = Dn
        if done:
            break
    F = H + Q.fock_rhf(packed, n, D)
    e_tot = 0.5 * np.sum(D * (F + H)) + enr
    assert abs(e_tot - (-183.32315970625)) < 1e-9
