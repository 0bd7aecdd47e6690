import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a device kernel that never returns must not hold the GPU box for ever: with pytest-timeout present every GPU test gets a generous
    # limit (the longest takes ~20 s); the "thread" method ends the process even while the main thread sits in a CUDA call
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(600, method="thread"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def npb():
    # the CUDA extension must exist: no fallback, fail loudly
    import __graft_entry__ as g
    import noparama_b200
    if not os.path.exists(noparama_b200.LIB_PATH):
        g.build()
    noparama_b200.load_library()
    return noparama_b200


@pytest.fixture(scope="session")
def ctx(npb):
    c = npb.Context(0)
    yield c
    c.close()
