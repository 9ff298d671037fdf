"""GPU (-m gpu): the tcgen05 / TMEM / bulk-copy GEMM pipeline against torch fp64 -- pins the shared-memory
descriptors, the canonical operand layouts, the split-bf16 scheme and the mbarrier protocol."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,n", [(64, 256), (256, 256), (256, 64), (128, 128), (64, 32), (80, 64), (64, 128), (32, 64), (128, 64)])
@pytest.mark.parametrize("twice", [False, True])
def test_tc_gemm_matches_fp64(k, n, twice):
    from moldiff_b200 import engine
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(k * 1000 + n)
    x = torch.randn(128, k, generator=g)
    w = torch.randn(k, n, generator=g) / k ** 0.5
    y = engine.tc_selftest(x.to(dev), w, twice=twice).cpu().double()
    ref = (x.double() @ w.double()) * (2.0 if twice else 1.0)
    err = float((y - ref).abs().max() / ref.abs().max())
    assert err < 4e-6, err        # bf16 + fp16 split (3 MMAs): ~2^-19 per product; a bf16 lo plane gives ~1e-5, single-pass bf16 ~4e-3


def test_tc_gemm_structured_input_detects_layout_errors():
    """Identity-like operands: any row / column / k permutation in the operand layouts shows up exactly."""
    from moldiff_b200 import engine
    dev = torch.device("cuda:0")
    k, n = 256, 256
    x = torch.zeros(128, k)
    x[torch.arange(128), torch.arange(128) * 2 % k] = 1.0          # row r selects k = 2r mod 256
    w = (torch.arange(k).float()[:, None] * 1000 + torch.arange(n).float()[None, :]) / 1024.0
    y = engine.tc_selftest(x.to(dev), w).cpu()
    assert torch.allclose(y, x @ w, rtol=0, atol=1e-3)
