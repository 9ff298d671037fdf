"""CPU: the oracle (oracle/restatement.py) against the golden vectors produced by the unmodified reference."""
import os

import pytest
import torch

from oracle import restatement as R
from oracle.ref_shim import reference_available
from tests.helpers import batch_inputs, doubled, oracle_moldiff

TOL = 2e-5   # oracle and reference are the same fp32 op sequence: observed 0 (bitwise) in the build container


def test_moldiff_forward_matches_reference_goldens(golden, seeded_models):
    sd = seeded_models[0].state_dict()
    for name, case in golden["moldiff_forward"].items():
        if case["args"]["B"] > 8:
            continue   # the B=32 case runs in test_moldiff_forward_b32 (slower)
        out = oracle_moldiff(sd, batch_inputs(**case["args"]))
        for k, ref in case["out"].items():
            assert R.rel_err(out[k], ref) < TOL, (name, k)


def test_moldiff_forward_b32(golden, seeded_models):
    sd = seeded_models[0].state_dict()
    case = golden["moldiff_forward"]["B32_t500"]
    inp = batch_inputs(**case["args"])
    assert len(inp["batch_node"]) == case["n_nodes"] == 747      # config 1 sizes
    assert len(inp["batch_halfedge"]) == case["n_half"] == 8907
    out = oracle_moldiff(sd, inp)
    for k, ref in case["out"].items():
        assert R.rel_err(out[k], ref) < TOL, k


def test_block_trace(golden, seeded_models):
    sd = seeded_models[0].state_dict()
    trace = []
    oracle_moldiff(sd, batch_inputs(**golden["block_trace"]["args"]), trace=trace)
    b0 = golden["block_trace"]["blocks"][0]
    for nm, x in zip(("h_node", "pos", "h_edge"), trace[0]):
        assert R.rel_err(x, b0[nm]) < TOL
    for i, blk in enumerate(golden["block_trace"]["blocks"][1:], start=1):
        assert R.rel_err(trace[i][1], blk["pos"]) < TOL
        assert R.rel_err(trace[i][0][:8], blk["h_node_rows"]) < TOL
        assert R.rel_err(trace[i][2][:16], blk["h_edge_rows"]) < TOL


def test_bondpred_and_guidance(golden, seeded_models):
    sd = seeded_models[1].state_dict()
    for name, case in golden["bondpred"].items():
        inp = batch_inputs(**case["args"])
        ei, be, _ = doubled(inp)
        for gui in ("uncertainty", "entropy"):
            delta, logits = R.guidance_delta(sd, inp["h_node"], inp["pos"], inp["batch_node"], ei, be, inp["t"],
                                             gui_type=gui, gui_scale=1e-4)
            assert R.rel_err(logits, case["out"]["logits"]) < TOL
            assert R.rel_err(delta, case["out"][gui]) < 1e-4   # the bar north_star states for fp32 outputs


def test_transition_functions(golden, seeded_models):
    sd = seeded_models[0].state_dict()
    tr = golden["transitions"]
    i = tr["inputs"]
    inp = batch_inputs(B=6, t_values=(999, 600, 599, 1, 0, 300))
    bn, bh = inp["batch_node"], inp["batch_halfedge"]
    assert torch.equal(inp["t"], i["t"])
    assert R.rel_err(R.q_v_posterior(sd, "edge_transition", i["log_v0"], i["log_vt"], i["t"], bh), tr["edge_post"]) < 1e-6
    assert R.rel_err(R.q_v_posterior(sd, "node_transition", i["log_n0"], i["log_nt"], i["t"], bn), tr["node_post"]) < 1e-6
    assert R.rel_err(R.q_vt_pred(sd, "edge_transition", i["log_vt"], i["t"], bh), tr["edge_qvt"]) < 1e-6
    torch.manual_seed(11)
    nz = torch.randn_like(i["x_t"])
    assert R.rel_err(R.pos_prev_from_recon(sd, "pos_transition", i["x_t"], i["x0"], i["t"], bn, nz), tr["pos_prev_seed11"]) < 1e-6
    assert torch.equal(R.log_sample_categorical(i["log_v0"], i["uniform"]), tr["gumbel_argmax"])


def test_sample50_teacher_forced(golden):
    """Per-step predictions of the reference's own 50-step run, teacher-forced through the oracle."""
    from moldiff_b200 import MolDiff
    from moldiff_b200.config import builtin_config
    import numpy as np
    cfg = builtin_config("train/train_MolDiff_simple.yml").model
    cfg.diff.num_timesteps = 50
    torch.manual_seed(0)
    sd = MolDiff(cfg, 8, 6).state_dict()
    for k, (s, a) in golden["sample50"]["checksum"].items():
        assert float(sd[k].double().sum()) == s and float(sd[k].double().abs().sum()) == a, k
    np.random.seed(2023)
    ph = R.make_data_placeholder(3)
    ei = torch.cat([ph["halfedge_index"], ph["halfedge_index"].flip(0)], dim=1)
    be = torch.cat([ph["batch_halfedge"], ph["batch_halfedge"]], dim=0)
    for i, st in golden["sample50"]["steps"].items():
        t = torch.full((3,), st["step"], dtype=torch.long)
        with torch.no_grad():
            pr = R.moldiff_forward(sd, st["h_node"], st["pos"], ph["batch_node"], torch.cat([st["h_half"]] * 2, 0),
                                   ei, be, t, num_timesteps=50)
        for k, ref in st["preds"].items():
            assert R.rel_err(pr[k], ref) < TOL, (i, k)


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_oracle_vs_live_reference_fresh_inputs(seeded_models):
    """Re-run the unmodified reference on inputs that are NOT in the fixtures."""
    from oracle.ref_shim import load_reference, load_yaml_config
    ref = load_reference()
    torch.manual_seed(0)
    m = ref.model.MolDiff(load_yaml_config("configs/train/train_MolDiff.yml").model, 8, 6).eval()
    inp = batch_inputs(B=3, seed_graph=99, seed_inputs=5, t_values=(10, 990, 456), pos_scale=2.0)
    ei, be, he = doubled(inp)
    with torch.no_grad():
        out = m(inp["h_node"], inp["pos"], inp["batch_node"], he, ei, be, inp["t"])
    mine = oracle_moldiff(seeded_models[0].state_dict(), inp)
    for k in out:
        assert R.rel_err(mine[k], out[k]) < TOL
