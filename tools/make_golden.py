"""Regenerates tests/golden/ from the reference tree (run in the build container only; the GPU
box has no /root/reference).

  tests/golden/inputs/        Ftab, mybasis and the example ZMATs, copied byte for byte from
                              /root/reference (they are the *inputs* of the boundary)
  tests/golden/molden.json    orbital energies printed in the reference's own MOLDEN outputs
                              (examples/O/singlet, examples/Be, examples/NO): the only numbers the
                              reference ships that pin this path
  tests/golden/oracle_anchors.json  SCF energies / (00|00) values of the oracle, to be compared with
                              BASELINE.md section 2 (recorded there from the survey session)
  tests/golden/packed_<mol>.npy    oracle packed ERIs of the small examples (GPU parity fixtures)
  tests/golden/cfour_mp2.json      MP2 energies of the CFOUR output the reference ships (examples/CO2/cfour/out)
  tests/golden/cfour_scf.json      CFOUR SCF energies (HeH, OH, CO2) the reference ships
  tests/golden/NO_Cui.npy          myQC's own UHF orbital coefficients of NO (examples/NO/Cui)
  tests/golden/ao2mo_CO2.npz       Cui / eig of the oracle's CO2 SCF and its (ia|jb) block: the ao2mo fixture
                                   (`python tools/make_golden.py ao2mo` regenerates only these two)
"""
import hashlib, json, os, re, shutil, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
G = os.path.join(ROOT, "tests", "golden")
INP = os.path.join(G, "inputs")
EX = {"CO2": "CO2", "HF": "HF", "CO": "CO", "NO": "NO", "OH": "OH", "HeH": "HeH", "H2": "H2", "H": "H",
      "Be": "Be", "O_singlet": "O/singlet", "O_triplet": "O/triplet"}


def molden_energies(path):
    txt = open(path).read().split("[MO]")[1]
    a, b = [], []
    for e, s in re.findall(r"Ene=\s*(\S+)\s*\n\s*Spin=\s*(\w+)", txt):
        (a if s == "Alpha" else b).append(float(e))
    return a, b


def main():
    os.makedirs(INP, exist_ok=True)
    for f in ("Ftab", "mybasis"):
        shutil.copyfile(os.path.join(REF, "src", f), os.path.join(INP, f))
    for k, v in EX.items():
        os.makedirs(os.path.join(INP, k), exist_ok=True)
        shutil.copyfile(os.path.join(REF, "examples", v, "ZMAT"), os.path.join(INP, k, "ZMAT"))
    md5 = hashlib.md5(open(os.path.join(INP, "Ftab"), "rb").read()).hexdigest()
    assert md5 == "593fd03e156236312577a4390288756a", md5
    molden = {}
    for k, v in (("O_singlet", "O/singlet"), ("Be", "Be"), ("NO", "NO")):
        a, b = molden_energies(os.path.join(REF, "examples", v, "MOLDEN"))
        molden[k] = {"alpha": a, "beta": b, "source": f"examples/{v}/MOLDEN"}
    json.dump(molden, open(os.path.join(G, "molden.json"), "w"), indent=1)
    ft = O.read_ftab(os.path.join(INP, "Ftab"))
    mb = open(os.path.join(INP, "mybasis")).read()
    anchors = {}
    for k in EX:
        if k == "O_triplet":
            continue
        mol = O.parse_zmat(open(os.path.join(INP, k, "ZMAT")).read())
        b = O.build_basis(mb, mol.atoms)
        xx, _ = O.int2e_dense(mol, b, ft)
        pk = O.packed_from_dense(xx)
        np.save(os.path.join(G, f"packed_{k}.npy"), pk)
        S, H = O.int1e(mol, b, ft)
        nA, nB = O.electrons(mol)
        enr = O.nuclear_repulsion(mol)
        if nA == nB:
            E, eps, _ = O.scf_rhf(S, H, xx, nA + nB, enr)
        else:
            E, ea, eb, _ = O.scf_uhf(S, H, xx, nA, nB, enr)
        anchors[k] = {"E_scf": E, "xx0000": float(xx[0, 0, 0, 0]), "norb": b.norb, "nset": b.nset,
                      "nonzero_canonical_gt_1e-13": int((np.abs(pk) > 1e-13).sum())}
        print(k, anchors[k])
    json.dump(anchors, open(os.path.join(G, "oracle_anchors.json"), "w"), indent=1)


def ao2mo_golden():
    out = open(os.path.join(REF, "examples", "CO2", "cfour", "out")).read()
    vals = {k: float(re.search(re.escape(k) + r"\s*=\s*(-?\d+\.\d+)", out).group(1))
            for k in ("E(SCF)", "E2(AA)", "E2(AB)", "E2(TOT)")}
    vals["source"] = "examples/CO2/cfour/out (CFOUR, true pi; myQC's float32 pi moves E2 by 1e-6)"
    json.dump(vals, open(os.path.join(G, "cfour_mp2.json"), "w"), indent=1)
    ft = O.read_ftab(os.path.join(INP, "Ftab"))
    mb = open(os.path.join(INP, "mybasis")).read()
    mol = O.parse_zmat(open(os.path.join(INP, "CO2", "ZMAT")).read())
    b = O.build_basis(mb, mol.atoms)
    xx, _ = O.int2e_dense(mol, b, ft)
    S, H = O.int1e(mol, b, ft)
    nA, nB = O.electrons(mol)
    E, eps, _, C = O.scf_rhf(S, H, xx, nA + nB, O.nuclear_repulsion(mol), orbitals=True)
    om = O.ao2mo_idx_trans(xx, C[:, :nA], C[:, nA:], C[:, :nA], C[:, nA:])
    e_aa, e_ab, e2 = O.mp2_rhf_energy(O.ao2mo_files("mp2_rhf", xx, C, C, nA, nB)["ijab_AB"], eps, nA, b.norb - nA)
    np.savez(os.path.join(G, "ao2mo_CO2.npz"), C=C, eig=eps, iajb=om, nocc=nA, e_scf=E, e2_aa=e_aa, e2_ab=e_ab, e2=e2)
    print(vals, E, e_aa, e_ab, e2)


def cfour_and_cui_golden():
    """Independent SCF numbers the reference ships: CFOUR total energies (examples/{HeH,OH,CO2}/cfour/out; examples/H2/cfour
    is another geometry, R = 1.0 A, and is not used) and myQC's own UHF orbital coefficients of NO (examples/NO/Cui,
    written by scf.f90:1082-1083 as CuiA(:,:) then CuiB(:,:), list-directed, column-major)."""
    vals = {}
    for name in ("HeH", "OH", "CO2"):
        out = open(os.path.join(REF, "examples", name, "cfour", "out")).read()
        vals[name] = {"E(SCF)": float(re.search(r"E\(SCF\)\s*=\s*(-?\d+\.\d+)", out).group(1)),
                      "source": f"examples/{name}/cfour/out"}
    json.dump(vals, open(os.path.join(G, "cfour_scf.json"), "w"), indent=1)
    cui = np.array([float(x.replace("D", "E")) for x in open(os.path.join(REF, "examples", "NO", "Cui")).read().split()])
    assert cui.size == 200
    np.save(os.path.join(G, "NO_Cui.npy"), cui.reshape(2, 10, 10))  # [spin][MO i][AO u]: Cui(u,i), u fastest
    print(vals)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ao2mo":
        ao2mo_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "cfour":
        cfour_and_cui_golden()
    else:
        main()
        ao2mo_golden()
        cfour_and_cui_golden()
