"""CPU: the packed-blob dataflow emulator (oracle/blob_emulator.py == what the kernels compute, in PyTorch)
against the as-written oracle.  Proves packing.py + the hoisted/CSR-ordered algorithm without a GPU."""
import torch

from moldiff_b200 import packing
from oracle import blob_emulator as BE
from oracle import restatement as R
from tests.helpers import batch_inputs, doubled, oracle_moldiff

TOL = 2e-5


def _blob(sd, **kw):
    return BE.Blob(*packing.pack_network(sd, **kw))


def test_moldiff_dataflow(seeded_models):
    sd = seeded_models[0].state_dict()
    W = _blob(sd, kind=1, net_prefix="denoiser", num_blocks=6, update_pos=True, time_dim=10)
    inp = batch_inputs(B=3, t_values=(999, 400, 0), pos_scale=1.5)
    ei, be, he = doubled(inp)
    ref = oracle_moldiff(sd, inp)
    with torch.no_grad():
        pn, pp, ph = BE.forward(W, kind=1, num_blocks=6, update_pos=True, rbf_lo=0.0, rbf_hi=15.0, time_dim=10,
                                T=1000.0, kn=8, ke=6, h_node_in=inp["h_node"], pos=inp["pos"], h_edge_in=he,
                                edge_index=ei, batch_node=inp["batch_node"], batch_edge=be, t=inp["t"])
    assert R.rel_err(pn, ref["pred_node"]) < TOL
    assert R.rel_err(pp, ref["pred_pos"]) < TOL
    assert R.rel_err(ph, ref["pred_halfedge"]) < TOL


def test_bondpred_dataflow(seeded_models):
    sd = seeded_models[1].state_dict()
    W = _blob(sd, kind=2, net_prefix="encoder", num_blocks=8, update_pos=False, time_dim=20)
    inp = batch_inputs(B=3, t_values=(999, 400, 0))
    ei, be, _ = doubled(inp)
    with torch.no_grad():
        ref = R.bondpred_forward(sd, inp["h_node"], inp["pos"], inp["batch_node"], ei, be, inp["t"])
        out = BE.forward(W, kind=2, num_blocks=8, update_pos=False, rbf_lo=0.0, rbf_hi=20.0, time_dim=20,
                         T=1000.0, kn=8, ke=5, h_node_in=inp["h_node"], pos=inp["pos"], h_edge_in=None,
                         edge_index=ei, batch_node=inp["batch_node"], batch_edge=be, t=inp["t"])
    assert R.rel_err(out, ref) < TOL


def test_bare_net_shuffled_edges(seeded_models):
    sd = {"net." + k: v for k, v in seeded_models[0].denoiser.state_dict().items()}
    W = _blob(sd, kind=0, net_prefix="net", num_blocks=6, update_pos=True)
    inp = batch_inputs(B=2, t_values=(100, 900))
    ei, be, _ = doubled(inp)
    g = torch.Generator().manual_seed(9)
    sh = torch.randperm(ei.shape[1], generator=g)
    ei, be = ei[:, sh], be[sh]
    N, E = len(inp["batch_node"]), ei.shape[1]
    h_node, h_edge = torch.randn(N, 256, generator=g), torch.randn(E, 64, generator=g)
    nt = (inp["t"][inp["batch_node"]].float() / 1000).unsqueeze(-1)
    et = torch.rand(E, 1, generator=g)     # edge_time deliberately NOT tied to node_time
    with torch.no_grad():
        ref = R.node_edge_net(sd, "net", h_node, inp["pos"], h_edge, ei, nt, et, num_blocks=6, cutoff=15.0)
        out = BE.forward(W, kind=0, num_blocks=6, update_pos=True, rbf_lo=0.0, rbf_hi=15.0, time_dim=0, T=1.0,
                         kn=0, ke=0, h_node_in=h_node, pos=inp["pos"], h_edge_in=h_edge, edge_index=ei,
                         node_time=nt, edge_time=et)
    for a, b in zip(out, ref):
        assert R.rel_err(a, b) < TOL


def test_bondpred_backward_dataflow(seeded_models):
    """Manual input-gradient chain (what mdb_backward.cuh implements) == autograd through the oracle: exact in
    fp64 (1e-12), and in fp32 as reproducible as the reference's own gradient (see assert_gradient_parity)."""
    from tests.helpers import assert_gradient_parity
    sd = seeded_models[1].state_dict()
    inp = batch_inputs(B=8, t_values=(999, 400, 0, 650))
    ei, be, _ = doubled(inp)
    g = torch.Generator().manual_seed(3)
    d_logits = torch.randn(ei.shape[1] // 2, 5, generator=g)
    res = {}
    for dtype in (torch.float64, torch.float32):
        sdd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
        pos = inp["pos"].to(dtype).clone().requires_grad_(True)
        logits_ref = R.bondpred_forward(sdd, inp["h_node"].to(dtype), pos, inp["batch_node"], ei, be, inp["t"])
        grad_ref = torch.autograd.grad((logits_ref * d_logits.to(dtype)).sum(), pos)[0]
        blob, ho, bo = packing.pack_network(sd, kind=2, net_prefix="encoder", num_blocks=8, update_pos=False, time_dim=20)
        torch.set_default_dtype(dtype)
        try:
            with torch.no_grad():
                logits, d_pos = BE.bondpred_forward_backward(
                    BE.Blob(blob.to(dtype), ho, bo), num_blocks=8, rbf_lo=0.0, rbf_hi=20.0, time_dim=20, T=1000.0,
                    kn=8, ke=5, h_node_in=inp["h_node"].to(dtype), pos=inp["pos"].to(dtype), edge_index=ei,
                    batch_node=inp["batch_node"], batch_edge=be, t=inp["t"], d_logits=d_logits.to(dtype))
        finally:
            torch.set_default_dtype(torch.float32)
        res[dtype] = (logits, d_pos, logits_ref.detach(), grad_ref)
    l64, d64, lr64, g64 = res[torch.float64]
    assert R.rel_err(l64, lr64) < 1e-12 and R.rel_err(d64, g64) < 1e-10      # the formulas are exact
    l32, d32, lr32, g32 = res[torch.float32]
    assert R.rel_err(l32, lr32) < TOL
    assert_gradient_parity(d32, g32, g64, inp["batch_node"], "emulator fp32")
