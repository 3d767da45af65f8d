import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dit():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "dit_small.npz"))


@pytest.fixture(scope="session")
def golden_fns():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "diffusion_fns.npz"))


@pytest.fixture(scope="session")
def golden_interleaved():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "interleaved.npz"))


@pytest.fixture(scope="session")
def golden_timecond():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "timecond.npz"))


@pytest.fixture(scope="session")
def golden_cfg1():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "cfg1.npz"))


@pytest.fixture(scope="session")
def golden_sampler_logits():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "sampler_logits.npz"))


@pytest.fixture(scope="session")
def golden_attn_cache():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "attn_cache.npz"))
