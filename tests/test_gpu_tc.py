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


@pytest.mark.parametrize("k,n", [(64, 256), (256, 256), (256, 64), (128, 128), (64, 32), (80, 64)])
@pytest.mark.parametrize("twice", [False, True])
def test_tc_gemm_cross_first_order(k, n, twice):
    """Cross-first accumulation order (csrc/tc_pipe.cuh; all lo*hi / hi*lo MMAs before the hi*hi ones, hi planes of two K
    steps per ring slot in the second pass, odd stage counts for K = 80): same product, ~3x closer to fp64 because the
    hardware's truncating accumulate no longer cuts every cross term against a large accumulator (measured 4.8e-7 at
    K = 256 against 1.4e-6 interleaved and 5.3e-7 for cuBLAS fp32)."""
    from moldiff_b200 import engine
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(k * 1000 + n)
    x = torch.randn(128, k, generator=g)
    w = torch.randn(k, n, generator=g) / k ** 0.5
    ref = (x.double() @ w.double()) * (2.0 if twice else 1.0)
    y = engine.tc_selftest(x.to(dev), w, twice=twice, cross_first=True).cpu().double()
    err = float((y - ref).abs().max() / ref.abs().max())
    # `twice` chains a second GEMM onto the full accumulator: its cross terms meet a large value again (interleaved-order error)
    assert err < (4e-6 if twice else 1.5e-6), err


def test_accumulator_truncates_with_two_guard_bits():
    """Known-answer vectors of the tcgen05.mma kind::f16 fp32 accumulate on sm_100a (tools/tc_numerics.py): addends of one
    instruction are aligned to the largest exponent with two guard bits and the sum is TRUNCATED -- the reason for the
    cross-first order.  All operands are exactly representable in fp16 (x scaled by 2^10, small slots weighted 2^-12)."""
    from moldiff_b200 import engine
    dev = torch.device("cuda:0")
    K, N = 64, 32
    w = torch.ones(K, N)
    w[1:9] = 2.0 ** -12
    x = torch.zeros(128, K)
    s = 2.0 ** 10
    x[0, 0], x[0, 1] = 1.0, 1.5 * 2.0 ** -24 * 2.0 ** 12             # 1 + 0.75 ulp -> 1 (round-to-nearest would give 1 + ulp)
    x[1, 0] = 1.0
    x[1, 1:5] = 2.0 ** -25 * 2.0 ** 12                                # 4 x 2^-25: inside the two guard bits -> 1 + ulp
    x[2, 0] = 1.0
    x[2, 1:9] = 2.0 ** -26 * 2.0 ** 12                                # 8 x 2^-26: below them -> lost
    y = engine.tc_selftest((x * s).to(dev), w).cpu()[:, 0] / s
    assert float(y[0]) == 1.0
    assert float(y[1]) == 1.0 + 2.0 ** -23
    assert float(y[2]) == 1.0
