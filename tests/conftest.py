"""pytest configuration: the `gpu` marker, library build, oracle fixtures."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build whatever is missing (cheap no-op when up to date)
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def vv():
    import voxvision_b200
    return voxvision_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    return o


@pytest.fixture(scope="session")
def port(orc):
    """The CPU restatement (oracle/vox_oracle.c)."""
    return orc.Oracle("port")


@pytest.fixture(scope="session")
def ref(orc):
    """The reference's own object code (oracle/_ref); absent if never built."""
    if not orc.available("ref"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return orc.Oracle("ref")


@pytest.fixture(scope="session")
def checker(orc):
    """Oracle used to check the CUDA path: the reference build when present, else the port."""
    return orc.Oracle("ref" if orc.available("ref") else "port")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        meta = json.load(f)
    arrays = dict(np.load(os.path.join(GOLDEN_DIR, "golden.npz")))
    return meta, arrays


@pytest.fixture(scope="session")
def ds():
    from tools import datasets
    return datasets


@pytest.fixture(scope="session")
def skull(ds):
    return ds.volume_voxels("skull")


@pytest.fixture(scope="session")
def bonsai(ds):
    return ds.volume_voxels("bonsai")


@pytest.fixture(scope="session")
def gpu(vv):
    if not vv.gpu_available():
        pytest.skip("no CUDA device")
    vv.set_device(0)
    return vv
