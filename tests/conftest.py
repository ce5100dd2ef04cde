import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUTS = os.path.join(GOLDEN, "inputs")

EXAMPLES = ["H", "H2", "HeH", "Be", "O_singlet", "HF", "OH", "CO", "NO", "CO2"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


def example_zmat(name: str) -> str:
    d = os.path.join(INPUTS, name)
    if os.path.isdir(d):
        return open(os.path.join(d, "ZMAT")).read()
    from myqc_b200 import molecules
    return molecules.zmat(name)


@pytest.fixture(scope="session")
def oracle_inputs():
    from oracle import oracle as O
    ft = O.read_ftab(os.path.join(INPUTS, "Ftab"))
    mb = open(os.path.join(INPUTS, "mybasis")).read()
    return ft, mb


def oracle_system(name, oracle_inputs):
    from oracle import oracle as O
    ft, mb = oracle_inputs
    mol = O.parse_zmat(example_zmat(name))
    return mol, O.build_basis(mb, mol.atoms), ft


def product_system(name, tmpdir):
    import myqc_b200 as Q
    return Q.make_job(str(tmpdir), example_zmat(name), INPUTS)
