"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MolDiff denoising hot path.

A plain, *as-written* (unfused, un-hoisted) restatement of the reference algorithm in functional
PyTorch on CPU tensors, driven by a reference-schema ``state_dict``.  Every function cites the
reference file:line it follows (paths relative to pengxingang/MolDiff @ db62fa1b).  Nothing in the
product path (``moldiff_b200/``, ``models/``) may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do, and only
as the checker or the CPU baseline.

Parity pinning: the reference ships no tests / golden vectors, so this oracle is
pinned against outputs of the *unmodified reference itself*, run in the build container through
``oracle/ref_shim.py`` and committed as fixtures by ``tests/golden/make_golden.py``
(``tests/test_oracle_golden.py`` re-checks them on every run, and ``tests/test_oracle_vs_reference.py``
re-runs the live reference when /root/reference is present).

dtype: works in whatever dtype the state_dict / inputs carry (float32 for parity, float64 to
measure the noise floor).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------------
def linear(sd, name, x):
    """nn.Linear with optional bias (torch.nn.Linear semantics)."""
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    return F.linear(x, w, b)


def layer_norm(sd, name, x):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[name + ".bias"], 1e-5)


def mlp(sd, name, x, num_layer=2):
    """models/common.py:181-201 -- Linear -> LayerNorm -> ReLU -> ... -> Linear (norm=True, act relu,
    act_last=False).  Sequential indices: layer j's Linear sits at 3*j, its LayerNorm at 3*j+1."""
    for j in range(num_layer):
        x = linear(sd, f"{name}.net.{3 * j}", x)
        if j < num_layer - 1:
            x = layer_norm(sd, f"{name}.net.{3 * j + 1}", x)
            x = torch.relu(x)
    return x


def gaussian_smearing(sd, name, dist, start, stop):
    """models/common.py:233-237 -- clamp to [start, stop], exp(coeff * (d - offset)^2)."""
    d = dist.clamp_min(start).clamp_max(stop)
    d = d.reshape(-1, 1) - sd[name + ".offset"].reshape(1, -1)
    # NB: coeff * (d^2), in this order -- (coeff * d) * d rounds differently and, for the time embedding
    # (|exponent| up to ~40), perturbs exp() by ~2e-6, which the guidance gradient amplifies to ~2e-4.
    return torch.exp(sd[name + ".coeff"] * torch.pow(d, 2))


def scatter_sum(src, index, n):
    """torch_scatter.scatter_sum(src, index, dim=0, dim_size=n) as used at graph.py:50,279,283,394."""
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


# --------------------------------------------------------------------------------------------
# NodeEdgeNet and its blocks
# --------------------------------------------------------------------------------------------
def node_block(sd, p, x, edge_index, edge_attr, node_time):
    """models/graph.py:29-55 (NodeBlock.forward, use_gate=True)."""
    n = x.shape[0]
    row, col = edge_index[0], edge_index[1]
    h_node = mlp(sd, p + ".node_net", x)
    h_edge = mlp(sd, p + ".edge_net", edge_attr)
    msg = linear(sd, p + ".msg_net", h_edge * h_node[col])
    gate = mlp(sd, p + ".gate", torch.cat([edge_attr, x[col], node_time[col]], dim=-1))
    msg = msg * torch.sigmoid(gate)
    aggr = scatter_sum(msg, row, n)
    out = linear(sd, p + ".centroid_lin", x) + aggr
    out = layer_norm(sd, p + ".layer_norm", out)
    return linear(sd, p + ".out_transform", torch.relu(out))


def bond_ffn(sd, p, bond_in, node_in, time):
    """models/graph.py:133-141 (BondFFN.forward, use_gate=True)."""
    inter = linear(sd, p + ".bond_linear", bond_in) * linear(sd, p + ".node_linear", node_in)
    inter = mlp(sd, p + ".inter_module", inter)
    gate = mlp(sd, p + ".gate", torch.cat([bond_in, node_in, time], dim=-1))
    return inter * torch.sigmoid(gate)


def edge_block(sd, p, h_bond, bond_index, h_node, bond_time):
    """models/graph.py:268-295 (EdgeBlock.forward)."""
    n = h_node.shape[0]
    left, right = bond_index[0], bond_index[1]
    m_left = bond_ffn(sd, p + ".bond_ffn_left", h_bond, h_node[left], bond_time)
    m_left = scatter_sum(m_left, right, n)[left]
    m_right = bond_ffn(sd, p + ".bond_ffn_right", h_bond, h_node[right], bond_time)
    m_right = scatter_sum(m_right, left, n)[right]
    h = (m_left + m_right
         + linear(sd, p + ".node_ffn_left", h_node[left])
         + linear(sd, p + ".node_ffn_right", h_node[right])
         + linear(sd, p + ".self_ffn", h_bond))
    h = layer_norm(sd, p + ".layer_norm", h)
    return linear(sd, p + ".out_transform", torch.relu(h))


def pos_update(sd, p, h_node, h_edge, edge_index, rel, dist, edge_time):
    """models/graph.py:384-396 (PosUpdate.forward).  No epsilon on dist, like the reference."""
    left, right = edge_index[0], edge_index[1]
    lf = mlp(sd, p + ".left_lin_edge", h_node[left])
    rf = mlp(sd, p + ".right_lin_edge", h_node[right])
    w = bond_ffn(sd, p + ".edge_lin", h_edge, lf * rf, edge_time)
    force = w * rel / dist.unsqueeze(-1) / (dist.unsqueeze(-1) + 1.0)
    return scatter_sum(force, left, h_node.shape[0])


def node_edge_net(sd, p, h_node, pos, h_edge, edge_index, node_time, edge_time, *,
                  num_blocks, cutoff, update_pos=True, start=0.0, trace=None):
    """models/graph.py:348-374 (NodeEdgeNet.forward, update_edge=True).  `trace`, if a list, receives
    (h_node, pos, h_edge) after every block (used for block-level parity fixtures)."""
    rel = dist = g = None
    for i in range(num_blocks):
        if update_pos or i == 0:
            rel = pos[edge_index[0]] - pos[edge_index[1]]
            dist = torch.linalg.vector_norm(rel, dim=-1)
            g = gaussian_smearing(sd, p + ".distance_expansion", dist, start, cutoff)
        h_edge = linear(sd, f"{p}.edge_embs.{i}", torch.cat([h_edge, g], dim=-1))
        dn = node_block(sd, f"{p}.node_blocks_with_edge.{i}", h_node, edge_index, h_edge, node_time)
        h_edge = h_edge + edge_block(sd, f"{p}.edge_blocks.{i}", h_edge, edge_index, h_node, edge_time)
        h_node = h_node + dn
        if update_pos:
            pos = pos + pos_update(sd, f"{p}.pos_blocks.{i}", h_node, h_edge, edge_index, rel, dist, edge_time)
        if trace is not None:
            trace.append((h_node, pos, h_edge))
    return h_node, pos, h_edge


# --------------------------------------------------------------------------------------------
# MolDiff.forward / BondPredictor.forward
# --------------------------------------------------------------------------------------------
def moldiff_forward(sd, h_node_pert, pos_pert, batch_node, h_edge_pert, edge_index, batch_edge, t, *,
                    num_timesteps=1000, num_blocks=6, cutoff=15.0, trace=None):
    """models/model.py:204-234."""
    tn = t.index_select(0, batch_node)
    te = t.index_select(0, batch_edge)
    emb_n = gaussian_smearing(sd, "time_emb.0", tn.to(pos_pert.dtype), 0.0, float(num_timesteps))
    emb_e = gaussian_smearing(sd, "time_emb.0", te.to(pos_pert.dtype), 0.0, float(num_timesteps))
    h_node = torch.cat([linear(sd, "node_embedder", h_node_pert), emb_n], dim=-1)
    h_edge = torch.cat([linear(sd, "edge_embedder", h_edge_pert), emb_e], dim=-1)
    h_node, pos, h_edge = node_edge_net(
        sd, "denoiser", h_node, pos_pert, h_edge, edge_index,
        (tn.unsqueeze(-1) / num_timesteps).to(pos_pert.dtype),
        (te.unsqueeze(-1) / num_timesteps).to(pos_pert.dtype),
        num_blocks=num_blocks, cutoff=cutoff, update_pos=True, trace=trace)
    nh = h_edge.shape[0] // 2
    return {
        "pred_node": mlp(sd, "node_decoder", h_node),
        "pred_pos": pos,
        "pred_halfedge": mlp(sd, "edge_decoder", h_edge[:nh] + h_edge[nh:]),
    }


def bondpred_forward(sd, h_node, pos, batch_node, edge_index, batch_edge, t, *,
                     num_timesteps=1000, num_blocks=8, cutoff=20.0):
    """models/bond_predictor.py:128-162 (num_timesteps != 0 branch)."""
    h_edge = torch.cat([h_node[edge_index[0]], h_node[edge_index[1]]], dim=-1)
    tn = t.index_select(0, batch_node)
    te = t.index_select(0, batch_edge)
    emb_n = gaussian_smearing(sd, "time_emb", tn.to(pos.dtype), 0.0, float(num_timesteps))
    emb_e = gaussian_smearing(sd, "time_emb", te.to(pos.dtype), 0.0, float(num_timesteps))
    h_node = torch.cat([linear(sd, "node_embedder", h_node), emb_n], dim=-1)
    h_edge = torch.cat([linear(sd, "edge_embedder", h_edge), emb_e], dim=-1)
    h_node, _, h_edge = node_edge_net(
        sd, "encoder", h_node, pos, h_edge, edge_index,
        (tn.unsqueeze(-1) / max(num_timesteps, 1)).to(pos.dtype),
        (te.unsqueeze(-1) / max(num_timesteps, 1)).to(pos.dtype),
        num_blocks=num_blocks, cutoff=cutoff, update_pos=False)
    nh = h_edge.shape[0] // 2
    ext = torch.cat([h_edge[:nh] + h_edge[nh:],
                     h_node[edge_index[0, :nh]] + h_node[edge_index[1, :nh]]], dim=-1)
    return mlp(sd, "edge_decoder", ext, num_layer=3)


def guidance_delta(sd_bond, h_node_pert, pos_pert, batch_node, edge_index, batch_edge, t, *,
                   gui_type="uncertainty", gui_scale=1e-4, **kw):
    """models/model.py:309-325 -- `uncertainty` (the only variant shipped in a config) and `entropy`."""
    with torch.enable_grad():
        pos_in = pos_pert.detach().clone().requires_grad_(True)
        logits = bondpred_forward(sd_bond, h_node_pert.detach(), pos_in, batch_node, edge_index, batch_edge, t, **kw)
        if gui_type == "uncertainty":
            obj = torch.sigmoid(-torch.logsumexp(logits, dim=-1)).log().sum()
        elif gui_type == "entropy":
            prob = torch.softmax(logits, dim=-1)
            obj = (-torch.sum(prob * torch.log(prob + 1e-12), dim=-1)).log().sum()
        else:
            raise NotImplementedError(gui_type)
        grad = torch.autograd.grad(obj, pos_in)[0]
    return -grad * gui_scale, logits.detach()


# --------------------------------------------------------------------------------------------
# transitions (per-step posterior sampling)
# --------------------------------------------------------------------------------------------
def extract(coef, t, batch):
    """models/diffusion.py:60-72 with ndim=1 (caller unsqueezes)."""
    return coef[t][batch]


def pos_prev_from_recon(sd, p, x_t, x_recon, t, batch, noise):
    """models/transition.py:44-63; `noise` is the N(0,1) draw the reference takes with randn_like."""
    c0 = extract(sd[p + ".coef_x0"], t, batch).unsqueeze(-1)
    ct = extract(sd[p + ".coef_xt"], t, batch).unsqueeze(-1)
    mu = c0 * x_recon + ct * x_t
    sigma = extract(sd[p + ".std"], t, batch).unsqueeze(-1)
    x_prev = mu + sigma * noise
    return torch.where((t[batch] == 0).unsqueeze(-1), mu, x_prev)


def q_v_posterior(sd, p, log_v0, log_vt, t, batch, eps=1e-30):
    """models/transition.py:285-315 with v0_prob=True."""
    tm1 = torch.where(t - 1 < 0, torch.zeros_like(t), t - 1)
    fact1 = extract(sd[p + ".transpopse_q_onestep_mats"], t, batch)
    fact1 = torch.einsum("bj,bjk->bk", torch.exp(log_vt), fact1)
    fact2 = extract(sd[p + ".q_mats"], tm1, batch)
    fact2 = torch.einsum("bj,bjk->bk", torch.exp(log_v0), fact2)
    out = torch.log(fact1 + eps).clamp_min(-32.0) + torch.log(fact2 + eps).clamp_min(-32.0)
    out = out - torch.logsumexp(out, dim=-1, keepdim=True)
    return torch.where(t[batch].unsqueeze(-1) == 0, log_v0, out)


def log_sample_categorical(logits, uniform):
    """models/diffusion.py:79-85; `uniform` is the U[0,1) draw the reference takes with rand_like."""
    gumbel = -torch.log(-torch.log(uniform + 1e-30) + 1e-30)
    return (gumbel + logits).argmax(dim=-1)


def q_vt_pred(sd, p, log_v0, t, batch, eps=1e-30):
    """models/transition.py:262-268."""
    q = torch.einsum("...i,...ij->...j", log_v0.exp(), extract(sd[p + ".q_mats"], t, batch))
    return torch.log(q + eps).clamp_min(-32.0)


def compute_v_Lt(log_post_true, log_post_pred, log_v0, t, batch):
    """models/transition.py:317-329 with models/diffusion.py:85-91 (categorical_kl, log_categorical) inlined:
    KL(q(v_{t-1} | v_t, v_0) || p(v_{t-1} | v_t)) per item, the decoder NLL instead for items whose graph sits at t = 0."""
    kl = (log_post_true.exp() * (log_post_true - log_post_pred)).sum(dim=-1)
    nll = -(log_v0.exp() * log_post_pred).sum(dim=-1)
    mask = (t == 0).float()[batch]
    return mask * nll + (1 - mask) * kl


def loss_terms(sd, node_pos, time_step, batch_node, batch_halfedge, pred_node, pred_pos, pred_half,
               log_node_t, log_node_0, log_half_t, log_half_0):
    """models/model.py:166-201, discrete categorical space, bond_len_loss off (configs/train/train_MolDiff.yml):
    MSE on the positions + 100 x mean variational-bound term for atom and bond types."""
    loss_pos = F.mse_loss(pred_pos, node_pos)
    terms = {}
    for key, p, logits, log_t, log_0, batch in (("node", "node_transition", pred_node, log_node_t, log_node_0, batch_node),
                                                ("edge", "edge_transition", pred_half, log_half_t, log_half_0, batch_halfedge)):
        log_recon = F.log_softmax(logits, dim=-1)
        post_true = q_v_posterior(sd, p, log_0, log_t, time_step, batch)
        post_pred = q_v_posterior(sd, p, log_recon, log_t, time_step, batch)
        terms[key] = torch.mean(compute_v_Lt(post_true, post_pred, log_0, time_step, batch)) * 100
    total = loss_pos + terms["node"] + terms["edge"]
    return {"loss": total, "loss_pos": loss_pos, "loss_node": terms["node"], "loss_edge": terms["edge"]}


def index_to_log_onehot(x, k):
    """models/diffusion.py:53-57."""
    return torch.log(F.one_hot(x, k).float().clamp(min=1e-30))


def sample_step(sd, state, batch_node, halfedge_index, batch_halfedge, step, n_graphs, noise, *,
                sd_bond=None, guidance=None, num_node_types=8, num_edge_types=6,
                fwd_kw=None, bond_kw=None):
    """One iteration of the loop body at models/model.py:271-372 (discrete categorical space),
    teacher-forced: `state` = (h_node_pert, pos_pert, h_halfedge_pert, log_node_type, log_halfedge_type)
    and `noise` = dict(pos=N(0,1)[N,3], node=U[N,Kn], edge=U[Eh,Ke]) replace the reference's RNG draws.
    Returns (preds, new_state)."""
    h_node_pert, pos_pert, h_half_pert, log_node, log_half = state
    edge_index = torch.cat([halfedge_index, halfedge_index.flip(0)], dim=1)
    batch_edge = torch.cat([batch_halfedge, batch_halfedge], dim=0)
    t = torch.full((n_graphs,), step, dtype=torch.long)
    h_edge_pert = torch.cat([h_half_pert, h_half_pert], dim=0)
    preds = moldiff_forward(sd, h_node_pert, pos_pert, batch_node, h_edge_pert, edge_index, batch_edge, t,
                            **(fwd_kw or {}))
    pos_prev = pos_prev_from_recon(sd, "pos_transition", pos_pert, preds["pred_pos"], t, batch_node, noise["pos"])
    log_node = q_v_posterior(sd, "node_transition", F.log_softmax(preds["pred_node"], -1), log_node, t, batch_node)
    node_prev = log_sample_categorical(log_node, noise["node"])
    log_half = q_v_posterior(sd, "edge_transition", F.log_softmax(preds["pred_halfedge"], -1), log_half, t,
                             batch_halfedge)
    half_prev = log_sample_categorical(log_half, noise["edge"])
    if guidance is not None and guidance[1] > 0:
        delta, _ = guidance_delta(sd_bond, h_node_pert, pos_pert, batch_node, edge_index, batch_edge, t,
                                  gui_type=guidance[0], gui_scale=guidance[1], **(bond_kw or {}))
        pos_prev = pos_prev + delta
    new_state = (F.one_hot(node_prev, num_node_types).float(), pos_prev,
                 F.one_hot(half_prev, num_edge_types).float(), log_node, log_half)
    return preds, new_state


# --------------------------------------------------------------------------------------------
# synthetic batch (the reference's placeholder recipe)
# --------------------------------------------------------------------------------------------
def make_data_placeholder(n_graphs, max_size=None, rng=None):
    """utils/transforms.py:125-156 restated with numpy only (the module itself needs rdkit/lmdb/PyG at
    import time).  Uses the *global* numpy RNG when rng is None, exactly like the reference."""
    import numpy as np
    rs = np.random if rng is None else rng
    if max_size is None:
        n_list = rs.normal(24.923464980477522, 5.516291901819105, size=n_graphs)
    else:
        n_list = np.array([max_size] * n_graphs)
    n_list = n_list.astype("int64")
    batch_node, he, bh = [], [], []
    start = 0
    for i, n in enumerate(n_list):
        n = int(n)
        batch_node.append(np.full(n, i))
        iu = np.triu_indices(n, 1)
        he.append(np.stack(iu) + start)
        bh.append(np.full(len(iu[0]), i))
        start += n
    return {
        "batch_node": torch.from_numpy(np.concatenate(batch_node)).long(),
        "halfedge_index": torch.from_numpy(np.concatenate(he, axis=1)).long(),
        "batch_halfedge": torch.from_numpy(np.concatenate(bh)).long(),
    }


def rel_err(a, b):
    """Norm-relative error used by every parity test: max|a-b| / max|b|."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
