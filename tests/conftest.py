import os
import sys
import warnings

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

# np.in1d is deprecated under numpy 2.x; neither oracle nor product use it, but
# helper code that imports the reference may (never under -W error).
warnings.filterwarnings('ignore', category=DeprecationWarning)


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def tmp_h5(tmp_path):
    return str(tmp_path / 'out.h5')
