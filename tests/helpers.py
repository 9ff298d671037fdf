"""Seeded synthetic inputs shared by the CPU and GPU tests (identical to tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import restatement as R


def batch_inputs(B, seed_graph=2023, seed_inputs=1, pos_scale=1.0, t_values=(500,), kn=8, ke=6, max_size=None):
    np.random.seed(seed_graph)
    ph = R.make_data_placeholder(B, max_size=max_size)
    bn, hei, bh = ph["batch_node"], ph["halfedge_index"], ph["batch_halfedge"]
    g = torch.Generator().manual_seed(seed_inputs)
    N, Eh = len(bn), len(bh)
    h_node = torch.nn.functional.one_hot(torch.randint(0, kn, (N,), generator=g), kn).float()
    h_half = torch.nn.functional.one_hot(torch.randint(0, ke, (Eh,), generator=g), ke).float()
    pos = torch.randn(N, 3, generator=g) * pos_scale
    t = torch.tensor([t_values[i % len(t_values)] for i in range(B)], dtype=torch.long)
    return dict(batch_node=bn, halfedge_index=hei, batch_halfedge=bh, h_node=h_node, h_half=h_half, pos=pos, t=t)


def doubled(inp):
    ei = torch.cat([inp["halfedge_index"], inp["halfedge_index"].flip(0)], dim=1)
    be = torch.cat([inp["batch_halfedge"], inp["batch_halfedge"]], dim=0)
    he = torch.cat([inp["h_half"], inp["h_half"]], dim=0)
    return ei, be, he


def oracle_moldiff(sd, inp, **kw):
    ei, be, he = doubled(inp)
    with torch.no_grad():
        return R.moldiff_forward(sd, inp["h_node"], inp["pos"], inp["batch_node"], he, ei, be, inp["t"], **kw)


def to_dev(inp, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}


def per_molecule_rel_err(a, b, batch_node):
    """max|a-b| / max|b| per molecule (molecules are independent graphs, so errors do not mix across them)."""
    out = []
    for m in range(int(batch_node.max()) + 1):
        sel = batch_node == m
        out.append(float((a[sel].double() - b[sel].double()).abs().max() / b[sel].double().abs().max().clamp_min(1e-30)))
    return torch.tensor(out)


def typical(e, q=0.4):
    """Robust location of per-molecule errors: the 40th percentile.  A ReLU-mask flip moves ONE molecule by 1e-3 .. 1e-2
    (see assert_gradient_parity); in a 16-molecule batch 3 .. 8 molecules are affected, depending on the order of the float
    atomics of that particular forward, so the median proper sits on the edge of the outlier group."""
    return float(torch.quantile(e.double().flatten(), q))


def assert_gradient_parity(got, ref32, ref64, batch_node, what="", median_bar=5e-5, majority=0.5):
    """Parity bar for d objective / d pos (guidance).  The function is piecewise smooth (80+ ReLU layers), so a
    pre-activation within ~1e-7 of zero flips its mask between ANY two fp32 evaluation orders and moves that one
    molecule's gradient by 1e-3..1e-2 -- the reference's own fp32 autograd does this against its fp64 self
    (measured: 4 of 24 molecules beyond 1e-4, max 2e-2; DESIGN.md "guidance gradient parity").  Hence:
      * the typical molecule (40th percentile, see `typical`) must match to `median_bar` = 5e-5 (2x tighter than the 1e-4 bar; measured
        2e-6 .. 2e-5 for the guidance objectives; callers with a harder upstream gradient pass the 1e-4 bar itself),
      * a majority of molecules must individually meet 1e-4, at most one may be off by more than 5e-2 (observed once in ~300
        molecule evaluations, 7e-2, on the fp32 FFMA path) and none by more than 0.5,
      * against the fp64 truth we may not be worse than the fp32 reference itself is (small-sample slack)."""
    e32 = per_molecule_rel_err(got, ref32, batch_node)
    assert typical(e32) < median_bar, (what, "typical (40th percentile)", typical(e32), float(e32.median()))
    assert float((e32 < 1e-4).float().mean()) >= majority, (what, e32)
    assert int((e32 > 5e-2).sum()) <= 1 and float(e32.max()) < 0.5, (what, e32)
    if ref64 is not None:
        mine = int((per_molecule_rel_err(got, ref64, batch_node) > 1e-4).sum())
        theirs = int((per_molecule_rel_err(ref32, ref64, batch_node) > 1e-4).sum())
        assert mine <= theirs + max(2, len(e32) // 4), (what, mine, theirs)
