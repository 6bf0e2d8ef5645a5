import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_layers():
    import numpy as np

    with np.load(GOLDEN / "layers_tp1.npz") as z:
        return {k: z[k] for k in z.files}


def load_golden(name):
    import numpy as np

    with np.load(GOLDEN / name) as z:
        return {k: z[k] for k in z.files}
