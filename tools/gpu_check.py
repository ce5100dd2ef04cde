"""Quick parity sweep on a GPU box: product (CUDA) vs oracle (CPU) on small inputs."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import myqc_b200 as Q
from myqc_b200 import molecules
from oracle import oracle as O

INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
names = sys.argv[1:] or ["H", "H2", "HeH", "Be", "O_singlet", "HF", "OH", "CO", "NO", "CO2", "h2o", "h2o_2", "h2o_4"]
print("devices:", Q.device_count())
worst = 0.0
for name in names:
    zm = open(os.path.join(INP, name, "ZMAT")).read() if os.path.isdir(os.path.join(INP, name)) else molecules.zmat(name)
    with tempfile.TemporaryDirectory() as d:
        s = Q.make_job(d, zm, INP)
    mol = O.parse_zmat(zm)
    b = O.build_basis(open(os.path.join(INP, "mybasis")).read(), mol.atoms)
    assert np.array_equal(b.setinfo[:len(s.setinfo)], s.setinfo[:len(b.setinfo)]) and np.array_equal(b.set, s.set)
    t0 = time.time(); ref = O.int2e_packed(mol, b, s.ftab); t1 = time.time()
    got = Q.eri_packed(s); t2 = time.time()
    err = np.abs(got - ref).max()
    nz_ref = int((ref != 0).sum()); nz_got = int((got != 0).sum())
    worst = max(worst, err)
    print(f"{name:10s} norb={s.norb:4d} unique={s.nunique:10d} max|gpu-oracle|={err:.3e} nonzero ref/gpu={nz_ref}/{nz_got} oracle {t1-t0:.2f}s gpu(e2e) {t2-t1:.3f}s", flush=True)
print("WORST", worst)
sys.exit(0 if worst < 1e-10 else 1)
