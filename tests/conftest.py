import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tools"), os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BENCH_SENTENCE = "The quick brown fox jumped over the sleeping dog."   # reference demos/pocket-tts.cpp:231


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def model_dir():
    from make_assets import default_model_dir
    return default_model_dir(eos_mode="never")


@pytest.fixture(scope="session")
def model_dir_eos():
    from make_assets import default_model_dir
    return default_model_dir(eos_mode="mid")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def orc(oracle_mod, model_dir):
    return oracle_mod.Oracle(model_dir, threads=os.cpu_count() or 1)


@pytest.fixture(scope="session")
def orc_eos(oracle_mod, model_dir_eos):
    return oracle_mod.Oracle(model_dir_eos, threads=os.cpu_count() or 1)


@pytest.fixture(scope="session")
def P():
    import ptts_b200
    ptts_b200.build()
    return ptts_b200


def snr_db(ref, x):
    import numpy as np
    err = (ref.astype(np.float64) - x.astype(np.float64))
    return 10 * np.log10((ref.astype(np.float64) ** 2).sum() / max((err ** 2).sum(), 1e-30))
