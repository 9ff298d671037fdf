import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def seeded_models():
    """Product modules initialised exactly like the reference constructors under torch.manual_seed(0)
    (bitwise equality of the init is asserted against golden checksums in test_host_logic.py)."""
    import torch
    from moldiff_b200 import BondPredictor, MolDiff
    from moldiff_b200.config import builtin_config
    torch.manual_seed(0)
    md = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).eval()
    torch.manual_seed(0)
    bp = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).eval()
    return md, bp
