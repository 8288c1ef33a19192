import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def engine_module():
    return importlib.import_module("experimental-tfhe_b200")


@pytest.fixture(scope="session")
def gate_oracle():
    import oracle_lib
    return oracle_lib.GateOracle(seed=42)


@pytest.fixture(scope="session")
def cb_oracle_nopriv():
    import oracle_lib
    return oracle_lib.CBOracle(seed=42, with_privks=False)


@pytest.fixture(scope="session")
def cb_oracle():
    import oracle_lib
    return oracle_lib.CBOracle(seed=42, with_privks=True)


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine on cuda:0.  Fails loudly (no fallback) when the library or the GPU is missing."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    mod = engine_module()
    return mod.Engine(0)


@pytest.fixture(scope="session")
def gate_engine(engine, gate_oracle):
    engine.load_gate_keys(gate_oracle.engine_params(), gate_oracle.bk, gate_oracle.ks)
    return engine
