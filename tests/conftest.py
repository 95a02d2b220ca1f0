import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


def load_eagen():
    """import the product package (its directory name has hyphens, so go through importlib)"""
    import importlib.util
    name = "eagen_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg = os.path.join(os.path.dirname(HERE), "halo2-liam-eagen-msm_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg, "__init__.py"), submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def eagen():
    mod = load_eagen()
    if not os.path.exists(mod.LIB_PATH):
        import subprocess
        subprocess.check_call([sys.executable, os.path.join(os.path.dirname(mod.LIB_PATH), "build.py")])
    mod.lib()
    return mod


@pytest.fixture(scope="session")
def gpu_ctx(eagen):
    """contexts per curve on cuda:0; GPU tests fail (not skip) if the device or the library is missing"""
    ctxs = {}

    def get(curve):
        if curve not in ctxs:
            ctxs[curve] = eagen.Context(curve, 0)
        return ctxs[curve]
    yield get
    for c in ctxs.values():
        c.close()
