import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def tiny_cfgs():
    from gst_visdial_b200 import weights as W
    return W.load_json_config(W.TINY_ENC_CONFIG), W.load_json_config(W.TINY_DEC_CONFIG)


@pytest.fixture(scope="session")
def tiny_sd(tiny_cfgs):
    from gst_visdial_b200 import weights as W
    return W.synthetic_state_dict(*tiny_cfgs, seed=0)


@pytest.fixture(scope="session")
def full_cfgs():
    from gst_visdial_b200 import weights as W
    return W.load_json_config(W.DEFAULT_ENC_CONFIG), W.load_json_config(W.DEFAULT_DEC_CONFIG)


@pytest.fixture(scope="session")
def full_sd(full_cfgs):
    from gst_visdial_b200 import weights as W
    return W.synthetic_state_dict(*full_cfgs, seed=0)
