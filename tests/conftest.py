import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("iridium-sniffer_b200.synth")


@pytest.fixture(scope="session")
def port():
    from oracle import bindings as ob
    ob.build(port=True, ref=False)
    return ob.Port()


def _ref(path_attr):
    from oracle import bindings as ob
    path = getattr(ob, path_attr)
    if not os.path.exists(path):
        if os.path.exists("/root/reference/burst_detect.c"):
            ob.build(port=False, ref=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return ob.Ref(path)


@pytest.fixture(scope="session")
def ref():
    return _ref("REF_SO")


@pytest.fixture(scope="session")
def ref_dif():
    return _ref("REF_DIF_SO")


@pytest.fixture(scope="session")
def rec_small(synth):
    """BASELINE config 1: seed 1234, 10 MHz cf32, 1.5 s, 12 bursts."""
    return synth.make_recording(1234, duration_s=1.5, n_bursts=12)
