"""Deterministic synthetic inputs of BASELINE.json / SURVEY.md section 8d, as ZMAT text.

No RNG; coordinates in Angstrom with 8 decimals, exactly as specified there:
  water monomer  O (0,0,0), H (+-0.7569503, 0.5858823, 0)
  (H2O)_n        simple-cubic lattice, spacing 3.0 A, identical orientation, atom order O,H,H,
                 molecules ordered i (slowest), j, k
  C20H42         all-trans alkane, C_k = (k*1.54*sin(theta), 0, +-0.77*cos(theta)),
                 theta = 54.7356103 deg, + for even k
"""
from __future__ import annotations

import math

_FOOT = "END\n\nCALC= SCF\nBASIS= STO-3G\nCHARGE= 0\nMULTI= 1\nREF= RHF\nMEMORY= 1000\nVERB= 1\n"


def _zmat(atoms):
    body = "".join(f"{s} {x:.8f} {y:.8f} {z:.8f}\n" for s, x, y, z in atoms)
    return "CARTESIAN\n" + body + _FOOT


def water_cluster(nx: int, ny: int, nz: int, spacing: float = 3.0) -> str:
    mono = [("O", 0.0, 0.0, 0.0), ("H", 0.7569503, 0.5858823, 0.0), ("H", -0.7569503, 0.5858823, 0.0)]
    atoms = []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                for s, x, y, z in mono:
                    atoms.append((s, x + spacing * i, y + spacing * j, z + spacing * k))
    return _zmat(atoms)


def alkane(nc: int = 20) -> str:
    th = math.radians(54.7356103)
    st, ct = math.sin(th), math.cos(th)
    carbons, sign = [], []
    for k in range(nc):
        s = 1.0 if k % 2 == 0 else -1.0
        carbons.append((k * 1.54 * st, 0.0, s * 0.77 * ct))
        sign.append(s)
    atoms = [("C",) + c for c in carbons]
    for (x, y, z), s in zip(carbons, sign):
        atoms.append(("H", x, y + 1.09 * 0.8164966, z + 1.09 * s * 0.5773503))
        atoms.append(("H", x, y - 1.09 * 0.8164966, z + 1.09 * s * 0.5773503))
    x, y, z = carbons[0]
    atoms.append(("H", x - 1.09 * st, y, z - 1.09 * sign[0] * ct))
    x, y, z = carbons[-1]
    atoms.append(("H", x + 1.09 * st, y, z - 1.09 * sign[-1] * ct))
    return _zmat(atoms)


WORKLOADS = {
    "h2o": lambda: water_cluster(1, 1, 1),
    "h2o_2": lambda: water_cluster(2, 1, 1),
    "h2o_4": lambda: water_cluster(2, 2, 1),
    "h2o_8": lambda: water_cluster(2, 2, 2),
    "h2o_16": lambda: water_cluster(4, 2, 2),
    "h2o_32": lambda: water_cluster(4, 4, 2),
    "h2o_64": lambda: water_cluster(4, 4, 4),
    "c20h42": lambda: alkane(20),
    "c4h10": lambda: alkane(4),
}


def zmat(name: str) -> str:
    return WORKLOADS[name]()
