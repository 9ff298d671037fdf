import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the in-tree library: without them they are skipped with the reason (they would
    otherwise fail one by one on 'Found no NVIDIA driver'); the driver's GPU run records which ran."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    have_lib = os.path.exists(os.path.join(ROOT, "moldiff_b200", "libmoldiff_b200.so"))
    if have_gpu and have_lib:
        return
    why = "no CUDA device" if not have_gpu else "libmoldiff_b200.so not built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_loss():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_loss.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_ref64():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "bondpred_ref64.pt"), weights_only=False)


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
