"""Training backward: weight AND input gradients for `MolDiff.forward` / `BondPredictor.forward`, so
that `loss.backward()` in `scripts/train_drug3d.py:88-109` / `train_bond.py` works against the drop-in modules.

Design of this round: *recompute-in-backward* (activation checkpointing at network granularity).

  forward   the hand-written sm_100a kernels, exactly as in inference (`mdb_moldiff_forward` / `mdb_bondpred_forward`): the
            loss VALUES come from the fused path and nothing per-edge is saved;
  backward  re-evaluates the network with the PyTorch operators below on the same device -- an exact-algebra restatement of
            the kernels' dataflow (per-node hoisted first layers, `index_add_` for the scatters, reference
            `models/graph.py:29-55,133-141,268-295,384-396`) over the LIVE parameters -- and lets `torch.autograd` produce
            d/d(parameters) and d/d(inputs) from it.

The per-edge weight-gradient contractions (dW = A^T dY, K = E) therefore still run in cuBLAS; hand-written tcgen05 split-K
kernels for them are the remaining part of row N2.  This module never runs under `torch.no_grad()` callers
(sampling, `get_loss` evaluation): they take the kernels only.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _lin(m, x):
    return F.linear(x, m.weight, m.bias)


def _mlp(m, x):
    """Linear -> LayerNorm -> ReLU -> ... -> Linear (nets.MLP; reference common.py:181-198)."""
    for layer in m.net:
        if isinstance(layer, torch.nn.Linear):
            x = F.linear(x, layer.weight, layer.bias)
        elif isinstance(layer, torch.nn.LayerNorm):
            x = F.layer_norm(x, layer.normalized_shape, layer.weight, layer.bias, layer.eps)
        else:
            x = torch.relu(x)
    return x


def _scatter(src, index, n):
    return torch.zeros(n, src.shape[1], dtype=src.dtype, device=src.device).index_add_(0, index, src)


def _smear(gs, v):
    d = v.clamp_min(gs.start).clamp_max(gs.stop).reshape(-1, 1) - gs.offset.reshape(1, -1)
    return torch.exp(gs.coeff * torch.pow(d, 2))


def _bond_ffn(m, bond, node_proj, node_gate_proj, time, n_bond, n_node):
    """BondFFN (graph.py:133-141) with the node-side first layers already applied per node and gathered:
    node_proj = node_linear(node)[idx], node_gate_proj = gate.net.0[:, node columns] node[idx]."""
    inter = F.linear(bond, m.bond_linear.weight) * node_proj
    inter = _mlp(m.inter_module, inter)
    g0 = m.gate.net[0]
    gate = F.linear(bond, g0.weight[:, :n_bond]) + node_gate_proj + time * g0.weight[:, n_bond + n_node] + g0.bias
    for layer in list(m.gate.net)[1:]:
        if isinstance(layer, torch.nn.Linear):
            gate = F.linear(gate, layer.weight, layer.bias)
        elif isinstance(layer, torch.nn.LayerNorm):
            gate = F.layer_norm(gate, layer.normalized_shape, layer.weight, layer.bias, layer.eps)
        else:
            gate = torch.relu(gate)
    return inter * torch.sigmoid(gate)


def node_edge_net(net, h_node, pos, h_edge, edge_index, node_time, edge_time):
    """NodeEdgeNet.forward (graph.py:348-374) in the kernels' hoisted form, differentiable in every argument and in the
    parameters of `net` (a nets.NodeEdgeNet)."""
    left, right = edge_index[0], edge_index[1]
    n, ed, nd = h_node.shape[0], net.edge_dim, net.node_dim
    rel = dist = g = None
    for i in range(net.num_blocks):
        if net.update_pos or i == 0:
            rel = pos[left] - pos[right]
            dist = torch.linalg.vector_norm(rel, dim=-1)
            g = _smear(net.distance_expansion, dist)
        e = _lin(net.edge_embs[i], torch.cat([h_edge, g], dim=-1))
        # ---- NodeBlock (graph.py:29-55): per-node tables first, gathered by the right node
        nb = net.node_blocks_with_edge[i]
        hn = _mlp(nb.node_net, h_node)
        g0 = nb.gate.net[0]
        gx = F.linear(h_node, g0.weight[:, ed:ed + nd]) + node_time * g0.weight[:, ed + nd] + g0.bias
        msg = _lin(nb.msg_net, _mlp(nb.edge_net, e) * hn[right])
        gate = F.linear(e, g0.weight[:, :ed]) + gx[right]
        for layer in list(nb.gate.net)[1:]:
            if isinstance(layer, torch.nn.Linear):
                gate = F.linear(gate, layer.weight, layer.bias)
            elif isinstance(layer, torch.nn.LayerNorm):
                gate = F.layer_norm(gate, layer.normalized_shape, layer.weight, layer.bias, layer.eps)
            else:
                gate = torch.relu(gate)
        agg = _scatter(msg * torch.sigmoid(gate), left, n)
        dn = _lin(nb.out_transform, torch.relu(F.layer_norm(_lin(nb.centroid_lin, h_node) + agg, (nd,), nb.layer_norm.weight,
                                                            nb.layer_norm.bias, nb.layer_norm.eps)))
        # ---- EdgeBlock (graph.py:268-295) on the OLD h_node
        eb = net.edge_blocks[i]
        out_l = _bond_ffn(eb.bond_ffn_left, e, F.linear(h_node, eb.bond_ffn_left.node_linear.weight)[left],
                          F.linear(h_node, eb.bond_ffn_left.gate.net[0].weight[:, ed:ed + nd])[left], edge_time, ed, nd)
        out_r = _bond_ffn(eb.bond_ffn_right, e, F.linear(h_node, eb.bond_ffn_right.node_linear.weight)[right],
                          F.linear(h_node, eb.bond_ffn_right.gate.net[0].weight[:, ed:ed + nd])[right], edge_time, ed, nd)
        sl, sr = _scatter(out_l, right, n), _scatter(out_r, left, n)
        u = sl[left] + sr[right] + _lin(eb.node_ffn_left, h_node)[left] + _lin(eb.node_ffn_right, h_node)[right] + _lin(eb.self_ffn, e)
        h_edge = e + _lin(eb.out_transform, torch.relu(F.layer_norm(u, (ed,), eb.layer_norm.weight, eb.layer_norm.bias,
                                                                     eb.layer_norm.eps)))
        h_node = h_node + dn
        # ---- PosUpdate (graph.py:384-396) on the NEW h_node / h_edge and the OLD rel / dist
        if net.update_pos:
            pu = net.pos_blocks[i]
            pf = _mlp(pu.left_lin_edge, h_node)[left] * _mlp(pu.right_lin_edge, h_node)[right]
            w = _bond_ffn(pu.edge_lin, h_edge, F.linear(pf, pu.edge_lin.node_linear.weight),
                          F.linear(pf, pu.edge_lin.gate.net[0].weight[:, ed:2 * ed]), edge_time, ed, ed)
            force = w * rel / dist.unsqueeze(-1) / (dist.unsqueeze(-1) + 1.0)
            pos = pos + _scatter(force, left, n)
    return h_node, pos, h_edge


def moldiff_forward(model, h_node_pert, pos_pert, batch_node, h_edge_pert, edge_index, batch_edge, t):
    """MolDiff.forward (model.py:204-234) over the live parameters of `model`."""
    tn, te = t.index_select(0, batch_node), t.index_select(0, batch_edge)
    T = float(model.num_timesteps)
    h_node = torch.cat([_lin(model.node_embedder, h_node_pert), _smear(model.time_emb[0], tn.to(pos_pert.dtype))], dim=-1)
    h_edge = torch.cat([_lin(model.edge_embedder, h_edge_pert), _smear(model.time_emb[0], te.to(pos_pert.dtype))], dim=-1)
    h_node, pos, h_edge = node_edge_net(model.denoiser, h_node, pos_pert, h_edge, edge_index,
                                        (tn.unsqueeze(-1) / T).to(pos_pert.dtype), (te.unsqueeze(-1) / T).to(pos_pert.dtype))
    nh = h_edge.shape[0] // 2
    return _mlp(model.node_decoder, h_node), pos, _mlp(model.edge_decoder, h_edge[:nh] + h_edge[nh:])


def bondpred_forward(model, h_node, pos, batch_node, edge_index, batch_edge, t):
    """BondPredictor.forward (bond_predictor.py:128-162) over the live parameters of `model`."""
    h_edge = torch.cat([h_node[edge_index[0]], h_node[edge_index[1]]], dim=-1)
    tn, te = t.index_select(0, batch_node), t.index_select(0, batch_edge)
    T = float(max(model.num_timesteps, 1))
    if model.num_timesteps != 0:
        x = torch.cat([_lin(model.node_embedder, h_node), _smear(model.time_emb, tn.to(pos.dtype))], dim=-1)
        he = torch.cat([_lin(model.edge_embedder, h_edge), _smear(model.time_emb, te.to(pos.dtype))], dim=-1)
    else:
        x, he = _lin(model.node_embedder, h_node), _lin(model.edge_embedder, h_edge)
    x, _, he = node_edge_net(model.encoder, x, pos, he, edge_index, (tn.unsqueeze(-1) / T).to(pos.dtype),
                             (te.unsqueeze(-1) / T).to(pos.dtype))
    nh = he.shape[0] // 2
    hs = he[:nh] + he[nh:]
    li, ri = edge_index[0, :nh], edge_index[1, :nh]
    return _mlp(model.edge_decoder, torch.cat([hs, x[li] + x[ri]], dim=-1))


class RecomputeBackward(torch.autograd.Function):
    """forward: `fused(*tensors)` -- the CUDA kernels; backward: differentiate `recompute(*tensors)` (PyTorch operators over the
    same live parameters) and hand the gradients of the float inputs and of every parameter back to autograd."""

    @staticmethod
    def forward(ctx, fused, recompute, n_inputs, *tensors):
        ctx.recompute, ctx.n_inputs = recompute, n_inputs
        ctx.save_for_backward(*tensors)
        with torch.no_grad():
            out = fused(*tensors[:n_inputs])
        return tuple(out) if isinstance(out, (tuple, list)) else out

    @staticmethod
    def backward(ctx, *grads):
        tensors = ctx.saved_tensors
        n = ctx.n_inputs
        ins = [x.detach().requires_grad_(True) if (x.is_floating_point() and ctx.needs_input_grad[3 + i]) else x.detach()
               for i, x in enumerate(tensors[:n])]
        params = list(tensors[n:])
        with torch.enable_grad():
            out = ctx.recompute(*ins)
            out = list(out) if isinstance(out, (tuple, list)) else [out]
            wanted = [x for x in ins if x.requires_grad] + [p for p in params if p.requires_grad]
            pairs = [(o, g) for o, g in zip(out, grads) if g is not None and o.requires_grad]
            got = torch.autograd.grad([o for o, _ in pairs], wanted, [g for _, g in pairs], allow_unused=True) if pairs else \
                [None] * len(wanted)
        it = iter(got)
        d_ins = [next(it) if (torch.is_tensor(x) and x.requires_grad) else None for x in ins]
        d_params = [next(it) if p.requires_grad else None for p in params]
        return (None, None, None, *d_ins, *d_params)


def needs_training_backward(module):
    """Training step: the module is in train() mode (scripts/train_drug3d.py:167 calls model.train(); the samplers call .eval(),
    sample_drug3d.py:80,91), autograd is recording and some parameter wants a gradient.  Guided sampling (eval mode, grad
    enabled only for the positions) keeps the hand-written input-gradient kernels."""
    return module.training and torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())
