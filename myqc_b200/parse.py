"""ZMAT -> envdat / nucpos / fmem: the `parse` stage upstream of int2e.

Restates src/parser/parser.f90 for the part int2e depends on: `cartesian` (:622-680),
`read_options` (:683-735; keys are case-sensitive, unknown keys are ignored) and `build`
(:435-547: centre-of-mass shift, Angstrom->bohr, electron counts, file writers).
T10: the reference's COM mass accumulator is uninitialised; zero reproduces the
geometry in examples/NO/MOLDEN.
"""
from __future__ import annotations

import os

import numpy as np

A2B = 1.8897161646320724  # parser.f90:14
ELEMENTS = ["H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne"]
MASS = [1.0, 4.0, 7.0, 9.0, 11.0, 12.0, 14.0, 16.0, 19.0, 20.0]  # parser.f90:449

# options(0:16) defaults, parser.f90:60
DEFAULTS = [0, 0, 0, 0, 0, 1, 1000, 1, 7, 0, 1, 0, 1, 0, 0, 0, 0]


def _opt_value(key: str, val: str):
    """(index, value) for the keys read_options understands (parser.f90:683-735) with the value tables of
    getcalc .. get_prop (:147-431); keys and values are case-sensitive, as in the reference."""
    if key == "CALC=":
        if val in ("SCF", "HF"):
            return 1, 0
        if val in ("MP2", "CIS"):
            return 1, {"MP2": 1, "CIS": 2}[val]
        raise ValueError("Sorry, that method has not been implimented. Exiting...")
    if key == "BASIS=":
        if val not in ("STO-3G", "tester1", "tester2", "tester3"):
            raise ValueError("Sorry, that basis has not been implimented. Exiting...")
        return 2, {"STO-3G": 0, "tester1": 1, "tester2": 2, "tester3": 3}[val]
    if key == "REF=":
        if val not in ("RHF", "UHF", "ROHF"):
            raise ValueError("Sorry, that reference has not been implimented. Exiting...")
        return 3, {"RHF": 0, "UHF": 1, "ROHF": 2}[val]
    if key == "PAR=":
        return 4, {"OMP": 2, "MPI": 3}.get(val, 1)
    if key == "NODES=":
        return 5, 1
    if key == "MEMORY=":
        return 6, 1000 if int(val) < 0 else int(val)
    if key == "VERB=":
        return 7, {"1": 1, "2": 2, "3": 3}.get(val, 0)
    if key == "SCF_Conv=":
        return 8, int(val)
    if key == "CHARGE=":
        return 9, int(val.replace("+", ""))
    if key == "MULTI=":
        if int(val) <= 0:
            raise ValueError("bad value for multiplicity, exiting.")
        return 10, int(val)
    if key == "UNITS=":
        return 11, 1 if val == "Bohr" else 0
    if key == "AO2MO=":  # getao2mo, parser.f90:359-371: every value selects the slow transform
        return 12, 1
    if key == "EXCITE=":  # getexcite, parser.f90:375-386
        return 13, 1 if val in ("CIS", "1") else 0
    if key == "ROOT_ALG=":  # getroot_alg, parser.f90:390-399
        return 14, 0
    if key == "E_NUM=":  # gete_num, parser.f90:403-414
        return 15, 1 if int(val) < 0 else int(val)
    if key == "PROP=":  # get_prop, parser.f90:418-431
        return 16, {"FIRST": 1, "1": 1, "SECOND": 2, "2": 2}.get(val, 0)
    return None


def parse_zmat(text: str):
    """Returns (atoms int32[n], xyz float64[n,3] in bohr after the COM shift, options int32[17])."""
    lines = [ln.replace(",", " ").split() for ln in text.splitlines()]
    k = 0
    while k < len(lines) and not lines[k]:
        k += 1
    if k == len(lines) or lines[k][0] not in ("CARTESIAN", "INTERNAL"):
        raise ValueError("Bad system type input. Exiting...")  # getsys, parser.f90:147-162
    if lines[k][0] != "CARTESIAN":
        raise ValueError("Sorry, that input style not supported yet")  # parser.f90:81-84
    atoms, coords = [], []
    k += 1
    while k < len(lines) and (not lines[k] or lines[k][0] != "END"):
        if lines[k]:  # list-directed reads skip blank records
            atoms.append(ELEMENTS.index(lines[k][0]) + 1)
            coords.append([float(x.replace("D", "E").replace("d", "E")) for x in lines[k][1:4]])
        k += 1
    if k == len(lines):
        raise ValueError("You need to put 'END' marker in ZMAT")
    if not atoms:
        raise ValueError("No atoms in system")
    options = list(DEFAULTS)
    # read_options (parser.f90:683-735): `READ(1,*)` skips one record after END, then KEY= VALUE records
    for r in lines[k + 2:]:
        if len(r) >= 2:
            kv = _opt_value(r[0], r[1])
            if kv is not None:
                options[kv[0]] = kv[1]
    atoms = np.array(atoms, dtype=np.int32)
    xyz = np.array(coords, dtype=np.float64)
    com = np.zeros(3)
    temp = 0.0
    for i in range(len(atoms)):
        m = MASS[atoms[i] - 1]
        for c in range(3):
            com[c] = com[c] + m * xyz[i, c]
        temp = temp + m
    com = com / temp
    xyz = xyz - com[None, :]
    if options[11] == 0:
        xyz = xyz * A2B
    return atoms, xyz, np.array(options, dtype=np.int32)


def electron_counts(atoms, options):
    """parser.f90:498-506"""
    charge, unpr = int(options[9]), int(options[10]) - 1
    nelc = int(np.sum(atoms)) - charge
    nB = (nelc - unpr) // 2
    return nB + unpr, nB


def write_job_files(workdir: str, atoms, xyz, options):
    """nucpos / envdat / fmem as `build` writes them (parser.f90:490-538), list-directed readable."""
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "nucpos"), "w") as f:
        for a, r in zip(atoms, xyz):
            f.write(f" {int(a):11d}  {r[0]:.17E}  {r[1]:.17E}  {r[2]:.17E}\n")
    nA, nB = electron_counts(atoms, options)
    with open(os.path.join(workdir, "envdat"), "w") as f:
        f.write(f" {len(atoms):11d}\n {nA:11d} {nB:11d}\n {len(options):11d}\n")
        f.write(" " + " ".join(f"{int(o):20d}" for o in options) + "\n\n")
        f.write(" #number of nuclei\n #number of electrons\n #length of options array\n options array\n")
    with open(os.path.join(workdir, "fmem"), "w") as f:
        f.write(f" {int(options[6]):20d}\n")


def parse(workdir: str):
    """The `parse` executable: ZMAT in `workdir` -> nucpos, envdat, fmem."""
    atoms, xyz, options = parse_zmat(open(os.path.join(workdir, "ZMAT")).read())
    write_job_files(workdir, atoms, xyz, options)
    return atoms, xyz, options
