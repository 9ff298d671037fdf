"""GPU parity of the fused reverse-transition step (mdb_transition_step) against the unfused
PyTorch operators that restate models/transition.py:44-63,285-315 and models/diffusion.py:79-85, on the same random
variates (torch's generator is consumed in the same order: positions, node types, half-edge types)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup(B, t_values, seed=3):
    from moldiff_b200 import MolDiff
    from moldiff_b200.config import builtin_config
    from moldiff_b200.placeholder import make_data_placeholder
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).to(dev).eval()
    np.random.seed(2023)
    ph = make_data_placeholder(B, device=dev)
    N, Eh = len(ph["batch_node"]), len(ph["batch_halfedge"])
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = torch.tensor([t_values[i % len(t_values)] for i in range(B)], dtype=torch.long, device=dev)
    x = dict(pos=torch.randn(N, 3, generator=g).to(dev), pred_pos=torch.randn(N, 3, generator=g).to(dev),
             pred_node=3 * torch.randn(N, 8, generator=g).to(dev), pred_half=3 * torch.randn(Eh, 6, generator=g).to(dev))
    x["log_node"] = torch.log(F.one_hot(torch.randint(0, 8, (N,), generator=g), 8).float().clamp(min=1e-30)).to(dev)
    x["log_half"] = torch.log(F.one_hot(torch.randint(0, 6, (Eh,), generator=g), 6).float().clamp(min=1e-30)).to(dev)
    return model, ph, t, x


def _unfused(model, ph, t, x):
    from moldiff_b200.transitions import gumbel_argmax
    pos_prev = model.pos_transition.get_prev_from_recon(x_t=x["pos"], x_recon=x["pred_pos"], t=t, batch=ph["batch_node"])
    log_node = model.node_transition.q_v_posterior(F.log_softmax(x["pred_node"], dim=-1), x["log_node"], t,
                                                   ph["batch_node"], v0_prob=True)
    cls_node = gumbel_argmax(log_node)
    log_half = model.edge_transition.q_v_posterior(F.log_softmax(x["pred_half"], dim=-1), x["log_half"], t,
                                                   ph["batch_halfedge"], v0_prob=True)
    cls_half = gumbel_argmax(log_half)
    return pos_prev, log_node, cls_node, log_half, cls_half


@pytest.mark.parametrize("t_values", [(999,), (500, 499, 1), (0,), (0, 1, 999, 250)])
def test_fused_transition_matches_pytorch_ops(t_values):
    from moldiff_b200 import engine
    model, ph, t, x = _setup(16, t_values)
    torch.manual_seed(11)
    ref = _unfused(model, ph, t, x)
    torch.manual_seed(11)
    pos_prev, log_node, h_node, log_half, h_edge2, half_type = engine.transition_step(
        model.pos_transition, model.node_transition, model.edge_transition, t, ph["batch_node"], ph["batch_halfedge"],
        x["pos"], x["pred_pos"], x["pred_node"], x["log_node"], x["pred_half"], x["log_half"])
    torch.cuda.synchronize()
    Eh = len(ph["batch_halfedge"])
    assert torch.allclose(pos_prev, ref[0], rtol=1e-6, atol=1e-6)
    # log-posteriors: entries floored at -32 (+ -32) are exact; the rest agree to fp32 rounding of log / exp
    assert torch.allclose(log_node, ref[1], rtol=1e-5, atol=2e-5)
    assert torch.allclose(log_half, ref[3], rtol=1e-5, atol=2e-5)
    # sampled classes: identical except where two Gumbel-perturbed scores tie to within rounding (none expected)
    assert (h_node.argmax(-1) == ref[2]).float().mean().item() > 0.999
    assert (half_type == ref[4]).float().mean().item() > 0.999
    assert torch.equal(h_node, F.one_hot(h_node.argmax(-1), 8).float())
    assert torch.equal(h_edge2[:Eh], F.one_hot(half_type, 6).float()) and torch.equal(h_edge2[Eh:], h_edge2[:Eh])


@pytest.mark.parametrize("t_values", [(999,), (500, 499, 1), (0,), (0, 1, 999, 250)])
def test_fused_transition_matches_oracle(t_values):
    """mdb_transition_step against the ORACLE's restatement of the reference operators (oracle/restatement.py:
    pos_prev_from_recon <- transition.py:44-63, q_v_posterior <- :285-315, log_sample_categorical <- diffusion.py:79-85) on CPU,
    with the Gaussian / uniform variates drawn once and handed to both sides."""
    from moldiff_b200 import engine
    from oracle import restatement as R
    model, ph, t, x = _setup(16, t_values)
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(17)
    N, Eh = x["pos"].shape[0], x["pred_half"].shape[0]
    noise = (torch.randn(N, 3, generator=g), torch.rand(N, 8, generator=g), torch.rand(Eh, 6, generator=g))
    c = {k: v.cpu() for k, v in x.items()}
    tc_, bn, bh = t.cpu(), ph["batch_node"].cpu(), ph["batch_halfedge"].cpu()
    ref_pos = R.pos_prev_from_recon(sd, "pos_transition", c["pos"], c["pred_pos"], tc_, bn, noise[0])
    ref_log_node = R.q_v_posterior(sd, "node_transition", F.log_softmax(c["pred_node"], -1), c["log_node"], tc_, bn)
    ref_node = R.log_sample_categorical(ref_log_node, noise[1])
    ref_log_half = R.q_v_posterior(sd, "edge_transition", F.log_softmax(c["pred_half"], -1), c["log_half"], tc_, bh)
    ref_half = R.log_sample_categorical(ref_log_half, noise[2])
    pos_prev, log_node, h_node, log_half, h_edge2, half_type = engine.transition_step(
        model.pos_transition, model.node_transition, model.edge_transition, t, ph["batch_node"], ph["batch_halfedge"],
        x["pos"], x["pred_pos"], x["pred_node"], x["log_node"], x["pred_half"], x["log_half"],
        noise=tuple(n.to(x["pos"].device) for n in noise))
    torch.cuda.synchronize()
    assert torch.allclose(pos_prev.cpu(), ref_pos, rtol=1e-6, atol=1e-6)
    assert torch.allclose(log_node.cpu(), ref_log_node, rtol=1e-5, atol=2e-5)
    assert torch.allclose(log_half.cpu(), ref_log_half, rtol=1e-5, atol=2e-5)
    # integer outputs: bit-exact except where two Gumbel-perturbed scores tie to within fp32 rounding of log / exp
    node_cls = h_node.argmax(-1).cpu()
    assert (node_cls == ref_node).float().mean().item() > 0.999 and (half_type.cpu() == ref_half).float().mean().item() > 0.999
    for got, ref, logp, u in ((node_cls, ref_node, ref_log_node, noise[1]), (half_type.cpu(), ref_half, ref_log_half, noise[2])):
        bad = (got != ref).nonzero().squeeze(-1)
        if len(bad):                      # every disagreement must be a near-tie of the two perturbed scores
            score = logp[bad] - torch.log(-torch.log(u[bad] + 1e-30) + 1e-30)
            gap = score.gather(1, ref[bad, None]) - score.gather(1, got[bad, None])
            assert float(gap.abs().max()) < 1e-4, gap
    assert torch.equal(h_node.cpu(), F.one_hot(node_cls, 8).float())
    assert torch.equal(h_edge2[:Eh].cpu(), F.one_hot(half_type.cpu(), 6).float()) and torch.equal(h_edge2[Eh:], h_edge2[:Eh])


def test_sample_step_fused_equals_unfused_distributionally():
    """Same seed, fused vs unfused sampler step: positions agree to the denoiser's run-to-run (atomic order) noise and
    the sampled types agree almost everywhere."""
    model, ph, _, _ = _setup(8, (999,))
    outs = []
    for fused in (False, True):
        model.fused_transition = fused
        torch.manual_seed(5)
        st = model.sample_begin(8, ph["batch_node"], ph["halfedge_index"], ph["batch_halfedge"])
        for step in (999, 998):
            model.sample_step(st, step)
        outs.append((st["pos"].clone(), st["h_node"].clone(), st["h_half"].clone()))
    model.fused_transition = True
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-4, atol=1e-4)
    assert (outs[0][1].argmax(-1) == outs[1][1].argmax(-1)).float().mean().item() > 0.99
    assert (outs[0][2].argmax(-1) == outs[1][2].argmax(-1)).float().mean().item() > 0.99


@pytest.mark.parametrize("guided", [False, True])
def test_graphed_step_matches_eager_step(guided):
    """CUDA-graphed loop body (MolDiff.graphed_step) against the eager `sample_step` from the same state: the denoiser
    predictions (no randomness involved) agree to the run-to-run atomic-order noise, replays advance the chain (time step
    read from the static tensor), states stay one-hot / finite, and the captured graph holds the whole step."""
    from moldiff_b200 import BondPredictor
    from moldiff_b200.config import builtin_config
    model, ph, _, _ = _setup(8, (999,))
    dev = ph["batch_node"].device
    bond, guidance = None, None
    if guided:
        torch.manual_seed(0)
        bond = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).to(dev).eval()
        guidance = ("uncertainty", 1e-4)
    torch.manual_seed(5)
    st = model.sample_begin(8, ph["batch_node"], ph["halfedge_index"], ph["batch_halfedge"])
    eager = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
    g = model.graphed_step(st, bond_predictor=bond, guidance=guidance)
    assert g.launches_per_replay >= (25 if not guided else 90)
    for k in ("h_node", "pos", "log_node", "log_half", "h_half"):          # construction rewinds the warm-up steps
        assert torch.equal(g.st[k], eager[k]), k
    for step in (999, 998, 997):
        state_before = {k: g.st[k].clone() for k in ("h_node", "pos", "h_half", "log_node", "log_half")}
        preds = {k: v.clone() for k, v in g.run(step).items()}
        work = dict(eager)
        work.update(state_before)
        ref = model.sample_step(work, step, bond_predictor=bond, guidance=guidance)
        torch.cuda.synchronize()
        for k in ref:
            scale = float(ref[k].abs().max())
            assert float((preds[k] - ref[k]).abs().max()) <= 2e-5 * scale, (step, k)
        assert torch.isfinite(g.st["pos"]).all()
        assert torch.all(g.st["h_node"].sum(-1) == 1) and torch.all(g.st["h_half"].sum(-1) == 1)
        assert not torch.equal(g.st["pos"], state_before["pos"])
    # the fused step's own noise: positions move by the posterior std of the step, not by something else
    assert float((g.st["pos"] - state_before["pos"]).abs().max()) < 10.0
