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
