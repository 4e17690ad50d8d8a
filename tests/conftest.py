import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_available() -> bool:
    try:
        from unmicst_b200 import _lib
        return _lib.lib().umx_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only host: the gpu-marked tests are skipped (they only run with a CUDA device)."""
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def sample_raw():
    from PIL import Image
    return np.array(Image.open(os.path.join(GOLDEN, "sample", "105.tif"))).astype(np.uint16)


@pytest.fixture(scope="session")
def sample_goldens():
    from PIL import Image
    g = Image.open(os.path.join(GOLDEN, "sample", "105_ContoursPM_1.tif"))
    g.seek(0)
    contours = np.array(g)
    g.seek(1)
    raw_page = np.array(g)
    nuclei = np.array(Image.open(os.path.join(GOLDEN, "sample", "105_NucleiPM_1.tif")))
    return dict(contours=contours, raw=raw_page, nuclei=nuclei)


@pytest.fixture(scope="session")
def nuclei_model():
    from unmicst_b200 import modelzoo
    return modelzoo.load_model(os.path.join(GOLDEN, "models", "nucleiDAPI"))


@pytest.fixture(scope="session")
def cyto_model():
    from unmicst_b200 import modelzoo
    return modelzoo.load_model(os.path.join(GOLDEN, "models", "CytoplasmIncell"))
