import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pincell_model():
    import raytracing_jl_b200 as rt

    d = np.load(os.path.join(ROOT, "tests", "golden", "pincell.npz"))
    return rt.UnstructuredDiscreteModel(d["node_coordinates"], d["cell_ptrs"], d["cell_data"])


@pytest.fixture(scope="session")
def pincell_mesh(pincell_model):
    import raytracing_jl_b200 as rt

    return rt.Mesh(pincell_model)


@pytest.fixture(scope="session")
def pincell_oracle_mesh(pincell_mesh):
    from oracle.oracle import OracleMesh

    return OracleMesh.from_mesh(pincell_mesh)
