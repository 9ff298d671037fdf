"""TEST INFRASTRUCTURE ONLY -- PyTorch emulation of the *kernel dataflow* on the *packed weight blob*.

Mirrors moldiff_b200/csrc/mdb_forward.cu step by step (CSR edge order, per-node hoisted tables, split
first layers, padded decoders) but in plain fp32 PyTorch on CPU.  Purpose: prove on a CPU-only box that
(a) moldiff_b200/packing.py lays the weights out the way the kernels read them and (b) the hoisted /
re-ordered algorithm equals the as-written oracle (oracle/restatement.py), so that a GPU mismatch can
only be a CUDA bug.  Never imported by the product path.
"""
import torch
import torch.nn.functional as F

from moldiff_b200 import packing

D, C, G = 256, 64, 16


class Blob:
    def __init__(self, blob, head_off, block_off):
        self.blob, self.head_off, self.block_off = blob, head_off, block_off

    def h(self, name, *shape):
        o = self.head_off[packing.HEAD_SLOTS.index(name)]
        assert o >= 0, name
        n = 1
        for s in shape:
            n *= s
        return self.blob[o:o + n].reshape(*shape)

    def b(self, i, name, *shape):
        o = self.block_off[i][packing.BLOCK_SLOTS.index(name)]
        assert o >= 0, name
        n = 1
        for s in shape:
            n *= s
        return self.blob[o:o + n].reshape(*shape)


def ln(x, g, b, relu=True):
    y = F.layer_norm(x, (x.shape[-1],), g, b, 1e-5)
    return torch.relu(y) if relu else y


def smear(v, lo, hi, off, coeff):
    d = v.clamp(lo, hi).reshape(-1, 1) - off.reshape(1, -1)
    return torch.exp(coeff * (d * d))


def sort_edges(edge_index, n_nodes):
    key = edge_index[0] * n_nodes + edge_index[1]
    perm = torch.argsort(key, stable=True)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(len(perm))
    return edge_index[0][perm], edge_index[1][perm], perm, inv


def scatter(src, idx, n):
    return torch.zeros(n, src.shape[1], dtype=src.dtype).index_add_(0, idx, src)


def forward(W: Blob, *, kind, num_blocks, update_pos, rbf_lo, rbf_hi, time_dim, T, kn, ke,
            h_node_in, pos, h_edge_in, edge_index, batch_node=None, batch_edge=None, t=None,
            node_time=None, edge_time=None, trace=None):
    N = h_node_in.shape[0]
    left, right, perm, inv = sort_edges(edge_index, N)
    E = len(perm)
    # ---- init kernels ----
    if kind == 0:
        x, tn = h_node_in.clone(), node_time.reshape(-1).clone()
        hedge, te = h_edge_in[perm].clone(), edge_time.reshape(-1)[perm].clone()
    else:
        tt = t[batch_node].float()
        toff, tco = W.h("TIME_OFFSET", time_dim), W.h("TIME_COEFF", time_dim)
        x = torch.cat([h_node_in @ W.h("NODE_EMB_W", kn, D - time_dim), smear(tt, 0.0, T, toff, tco)], -1)
        tn = tt / T
        tte = t[batch_edge[perm]].float()
        if kind == 1:
            emb = h_edge_in[perm] @ W.h("EDGE_EMB_W", ke, C - time_dim)
        else:
            we = W.h("EDGE_EMB_W", 2 * kn, C - time_dim)
            emb = h_node_in[left] @ we[:kn] + h_node_in[right] @ we[kn:]
        hedge = torch.cat([emb, smear(tte, 0.0, T, toff, tco)], -1)
        te = tte / T
    roff, rco = W.h("RBF_OFFSET", G), W.h("RBF_COEFF", G)

    def pre(i):
        tb = {}
        a = ln(x @ W.b(i, "NB_NN1_W", D, D) + W.b(i, "NB_NN1_B", D), W.b(i, "NB_NN1_G", D), W.b(i, "NB_NN1_BE", D))
        tb["hn"] = a @ W.b(i, "NB_NN2_W", D, D) + W.b(i, "NB_NN2_B", D)
        tb["gx"] = x @ W.b(i, "NB_GX_W", D, D) + W.b(i, "NB_G1_B", D) + tn[:, None] * W.b(i, "NB_GT_W", D)
        tb["cen"] = x @ W.b(i, "NB_CEN_W", D, D) + W.b(i, "NB_CEN_B", D)
        for s, S in (("l", "EL"), ("r", "ER")):
            tb["nl" + s] = x @ W.b(i, S + "_NL_W", D, 128)
            tb["gn" + s] = x @ W.b(i, S + "_GN_W", D, 32) + W.b(i, S + "_G1_B", 32)
        tb["fl"] = x @ W.b(i, "EB_NFL_W", D, C) + W.b(i, "EB_NFL_B", C)
        tb["fr"] = x @ W.b(i, "EB_NFR_W", D, C) + W.b(i, "EB_NFR_B", C)
        return tb

    for i in range(num_blocks):
        tb = pre(i)
        # ---- edge kernel B ----
        rel = pos[left] - pos[right]
        d = torch.sqrt((rel * rel).sum(-1))
        g = smear(d, rbf_lo, rbf_hi, roff, rco)
        e = torch.cat([hedge, g], -1) @ W.b(i, "EE_W", C + G, C) + W.b(i, "EE_B", C)
        a = ln(e @ W.b(i, "NB_EN1_W", C, D) + W.b(i, "NB_EN1_B", D), W.b(i, "NB_EN1_G", D), W.b(i, "NB_EN1_BE", D))
        m = (a @ W.b(i, "NB_EN2_W", D, D) + W.b(i, "NB_EN2_B", D)) * tb["hn"][right]
        a = ln(e @ W.b(i, "NB_GE_W", C, D) + tb["gx"][right], W.b(i, "NB_G1_G", D), W.b(i, "NB_G1_BE", D))
        sg = torch.sigmoid(a @ W.b(i, "NB_G2_W", D, D) + W.b(i, "NB_G2_B", D))
        out = (m @ W.b(i, "NB_MSG_W", D, D) + W.b(i, "NB_MSG_B", D)) * sg
        agg = scatter(out, left, N)
        o = {}
        for s, S, node in (("l", "EL", left), ("r", "ER", right)):
            inter = (e @ W.b(i, S + "_BL_W", C, 128)) * tb["nl" + s][node]
            a = ln(inter @ W.b(i, S + "_I1_W", 128, 128) + W.b(i, S + "_I1_B", 128), W.b(i, S + "_I1_G", 128), W.b(i, S + "_I1_BE", 128))
            i2 = a @ W.b(i, S + "_I2_W", 128, C) + W.b(i, S + "_I2_B", C)
            g1 = e @ W.b(i, S + "_GB_W", C, 32) + tb["gn" + s][node] + te[:, None] * W.b(i, S + "_GT_W", 32)
            g1 = ln(g1, W.b(i, S + "_G1_G", 32), W.b(i, S + "_G1_BE", 32))
            o[s] = i2 * torch.sigmoid(g1 @ W.b(i, S + "_G2_W", 32, C) + W.b(i, S + "_G2_B", C))
        SL, SR = scatter(o["l"], right, N), scatter(o["r"], left, N)
        # ---- node kernel (mid) ----
        a = ln(tb["cen"] + agg, W.b(i, "NB_LN_G", D), W.b(i, "NB_LN_BE", D))
        x = x + a @ W.b(i, "NB_OUT_W", D, D) + W.b(i, "NB_OUT_B", D)
        if update_pos:
            lf = ln(x @ W.b(i, "PU_LL1_W", D, C) + W.b(i, "PU_LL1_B", C), W.b(i, "PU_LL1_G", C), W.b(i, "PU_LL1_BE", C)) \
                @ W.b(i, "PU_LL2_W", C, C) + W.b(i, "PU_LL2_B", C)
            rf = ln(x @ W.b(i, "PU_RL1_W", D, C) + W.b(i, "PU_RL1_B", C), W.b(i, "PU_RL1_G", C), W.b(i, "PU_RL1_BE", C)) \
                @ W.b(i, "PU_RL2_W", C, C) + W.b(i, "PU_RL2_B", C)
        # ---- edge kernel D ----
        u = e @ W.b(i, "EB_SELF_W", C, C) + W.b(i, "EB_SELF_B", C) + SL[left] + SR[right] + tb["fl"][left] + tb["fr"][right]
        a = ln(u, W.b(i, "EB_LN_G", C), W.b(i, "EB_LN_BE", C))
        hedge = e + a @ W.b(i, "EB_OUT_W", C, C) + W.b(i, "EB_OUT_B", C)
        if update_pos:
            P = lf[left] * rf[right]
            a = (hedge @ W.b(i, "PU_PB_W", C, D)) * (P @ W.b(i, "PU_PN_W", C, D))
            a = ln(a @ W.b(i, "PU_I1_W", D, D) + W.b(i, "PU_I1_B", D), W.b(i, "PU_I1_G", D), W.b(i, "PU_I1_BE", D))
            wi = a @ W.b(i, "PU_I2_W", D) + W.b(i, "PU_I2_B", 1)
            g1 = hedge @ W.b(i, "PU_GB_W", C, 32) + P @ W.b(i, "PU_GN_W", C, 32) + W.b(i, "PU_G1_B", 32) \
                + te[:, None] * W.b(i, "PU_GT_W", 32)
            g1 = ln(g1, W.b(i, "PU_G1_G", 32), W.b(i, "PU_G1_BE", 32))
            w = wi * torch.sigmoid(g1 @ W.b(i, "PU_G2_W", 32) + W.b(i, "PU_G2_B", 1))
            force = w[:, None] * rel / d[:, None] / (d[:, None] + 1.0)
            pos = pos + scatter(force, left, N)
        if trace is not None:
            trace.append((x.clone(), pos.clone(), hedge[inv].clone()))
    if kind == 0:
        return x, pos, hedge[inv]
    nh = E // 2
    hs = hedge[inv[:nh]] + hedge[inv[nh:]]
    if kind == 1:
        a = ln(x @ W.h("NDEC1_W", D, D) + W.h("NDEC1_B", D), W.h("NDEC1_G", D), W.h("NDEC1_BE", D))
        pred_node = (a @ W.h("NDEC2_W", D, 32) + W.h("NDEC2_B", 32))[:, :kn]
        a = ln(hs @ W.h("EDEC1_W", C, C) + W.h("EDEC1_B", C), W.h("EDEC1_G", C), W.h("EDEC1_BE", C))
        pred_half = (a @ W.h("EDEC2_W", C, 32) + W.h("EDEC2_B", 32))[:, :ke]
        return pred_node, pos, pred_half
    dect = x @ W.h("EDEC1N_W", D, C)
    lh, rh = left[inv[:nh]], right[inv[:nh]]
    a = ln(hs @ W.h("EDEC1_W", C, C) + W.h("EDEC1_B", C) + dect[lh] + dect[rh], W.h("EDEC1_G", C), W.h("EDEC1_BE", C))
    a = ln(a @ W.h("EDEC2_W", C, C) + W.h("EDEC2_B", C), W.h("EDEC3_G", C), W.h("EDEC3_BE", C))
    return (a @ W.h("EDEC3_W", C, 32) + W.h("EDEC3_B", 32))[:, :ke]
