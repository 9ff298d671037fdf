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


# ------------------------------------------------------------------------------------------------------------
# Backward dataflow of the bond predictor (mirrors moldiff_b200/csrc/mdb_backward.cuh, manual gradients)
# ------------------------------------------------------------------------------------------------------------
def ln_stats(x):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    return (x - mean) * rstd, rstd


def ln_relu_bwd(d, xh, rstd, g, b):
    dxh = torch.where(xh * g + b > 0, d * g, torch.zeros_like(d))
    m1 = dxh.mean(-1, keepdim=True)
    m2 = (dxh * xh).mean(-1, keepdim=True)
    return rstd * (dxh - m1 - xh * m2)


_backward_mm = None


def set_backward_mm(fn):
    """Test hook (tools/emulate_tc_backward.py): fn(a, w, grad) replaces `a @ w` for every GEMM that the tensor-core
    backward kernels execute (grad = True for gradient operands, False for their forward recompute)."""
    global _backward_mm
    _backward_mm = fn


_forward_mm = None


def set_forward_mm(fn):
    """Test hook: fn(a, w) replaces `a @ w` for every GEMM of the forward half of bondpred_forward_backward."""
    global _forward_mm
    _forward_mm = fn


def _fmm(a, w):
    return a @ w if _forward_mm is None else _forward_mm(a, w)


def _mm(a, w, grad=True):
    return a @ w if _backward_mm is None else _backward_mm(a, w, grad)


def bondpred_forward_backward(W: Blob, *, num_blocks, rbf_lo, rbf_hi, time_dim, T, kn, ke,
                              h_node_in, pos, edge_index, batch_node, batch_edge, t, d_logits):
    """Returns (logits, d_pos) with d_pos = d sum(logits * d_logits) / d pos, computed the way the kernels do."""
    N = h_node_in.shape[0]
    left, right, perm, inv = sort_edges(edge_index, N)
    E = len(perm)
    nh = E // 2
    tt = t[batch_node].float()
    toff, tco = W.h("TIME_OFFSET", time_dim), W.h("TIME_COEFF", time_dim)
    x = torch.cat([h_node_in @ W.h("NODE_EMB_W", kn, D - time_dim), smear(tt, 0.0, T, toff, tco)], -1)
    tn = tt / T
    tte = t[batch_edge[perm]].float()
    we = W.h("EDGE_EMB_W", 2 * kn, C - time_dim)
    hedge = torch.cat([h_node_in[left] @ we[:kn] + h_node_in[right] @ we[kn:], smear(tte, 0.0, T, toff, tco)], -1)
    te = tte / T
    roff, rco = W.h("RBF_OFFSET", G), W.h("RBF_COEFF", G)
    rel = pos[left] - pos[right]
    dist = torch.sqrt((rel * rel).sum(-1))
    g = smear(dist, rbf_lo, rbf_hi, roff, rco)

    def tables(i, x):
        tb = {}
        a = ln(_fmm(x, W.b(i, "NB_NN1_W", D, D)) + W.b(i, "NB_NN1_B", D), W.b(i, "NB_NN1_G", D), W.b(i, "NB_NN1_BE", D))
        tb["hn"] = _fmm(a, W.b(i, "NB_NN2_W", D, D)) + W.b(i, "NB_NN2_B", D)
        tb["gx"] = _fmm(x, W.b(i, "NB_GX_W", D, D)) + W.b(i, "NB_G1_B", D) + tn[:, None] * W.b(i, "NB_GT_W", D)
        tb["cen"] = _fmm(x, W.b(i, "NB_CEN_W", D, D)) + W.b(i, "NB_CEN_B", D)
        for s, S in (("l", "EL"), ("r", "ER")):
            tb["nl" + s] = _fmm(x, W.b(i, S + "_NL_W", D, 128))
            tb["gn" + s] = _fmm(x, W.b(i, S + "_GN_W", D, 32)) + W.b(i, S + "_G1_B", 32)
        tb["fl"] = _fmm(x, W.b(i, "EB_NFL_W", D, C)) + W.b(i, "EB_NFL_B", C)
        tb["fr"] = _fmm(x, W.b(i, "EB_NFR_W", D, C)) + W.b(i, "EB_NFR_B", C)
        return tb

    saved = []
    for i in range(num_blocks):
        tb = tables(i, x)
        e = _fmm(torch.cat([hedge, g], -1), W.b(i, "EE_W", C + G, C)) + W.b(i, "EE_B", C)
        a = ln(_fmm(e, W.b(i, "NB_EN1_W", C, D)) + W.b(i, "NB_EN1_B", D), W.b(i, "NB_EN1_G", D), W.b(i, "NB_EN1_BE", D))
        m = (_fmm(a, W.b(i, "NB_EN2_W", D, D)) + W.b(i, "NB_EN2_B", D)) * tb["hn"][right]
        a = ln(_fmm(e, W.b(i, "NB_GE_W", C, D)) + tb["gx"][right], W.b(i, "NB_G1_G", D), W.b(i, "NB_G1_BE", D))
        sg = torch.sigmoid(_fmm(a, W.b(i, "NB_G2_W", D, D)) + W.b(i, "NB_G2_B", D))
        agg = scatter((_fmm(m, W.b(i, "NB_MSG_W", D, D)) + W.b(i, "NB_MSG_B", D)) * sg, left, N)
        o = {}
        for s, S, node in (("l", "EL", left), ("r", "ER", right)):
            inter = (_fmm(e, W.b(i, S + "_BL_W", C, 128))) * tb["nl" + s][node]
            a = ln(_fmm(inter, W.b(i, S + "_I1_W", 128, 128)) + W.b(i, S + "_I1_B", 128), W.b(i, S + "_I1_G", 128), W.b(i, S + "_I1_BE", 128))
            i2 = _fmm(a, W.b(i, S + "_I2_W", 128, C)) + W.b(i, S + "_I2_B", C)
            g1 = _fmm(e, W.b(i, S + "_GB_W", C, 32)) + tb["gn" + s][node] + te[:, None] * W.b(i, S + "_GT_W", 32)
            g1 = ln(g1, W.b(i, S + "_G1_G", 32), W.b(i, S + "_G1_BE", 32))
            o[s] = i2 * torch.sigmoid(_fmm(g1, W.b(i, S + "_G2_W", 32, C)) + W.b(i, S + "_G2_B", C))
        SL, SR = scatter(o["l"], right, N), scatter(o["r"], left, N)
        saved.append(dict(x=x, e=e, agg=agg, SL=SL, SR=SR))
        a = ln(tb["cen"] + agg, W.b(i, "NB_LN_G", D), W.b(i, "NB_LN_BE", D))
        x = x + _fmm(a, W.b(i, "NB_OUT_W", D, D)) + W.b(i, "NB_OUT_B", D)
        u = _fmm(e, W.b(i, "EB_SELF_W", C, C)) + W.b(i, "EB_SELF_B", C) + SL[left] + SR[right] + tb["fl"][left] + tb["fr"][right]
        hedge = e + _fmm(ln(u, W.b(i, "EB_LN_G", C), W.b(i, "EB_LN_BE", C)), W.b(i, "EB_OUT_W", C, C)) + W.b(i, "EB_OUT_B", C)
    # decoder forward
    q1, q2 = inv[:nh], inv[nh:]
    hs = hedge[q1] + hedge[q2]
    dect = _fmm(x, W.h("EDEC1N_W", D, C))
    lh, rh = left[q1], right[q1]
    xh1, rs1 = ln_stats(_fmm(hs, W.h("EDEC1_W", C, C)) + W.h("EDEC1_B", C) + dect[lh] + dect[rh])
    r1 = torch.relu(xh1 * W.h("EDEC1_G", C) + W.h("EDEC1_BE", C))
    xh2, rs2 = ln_stats(_fmm(r1, W.h("EDEC2_W", C, C)) + W.h("EDEC2_B", C))
    r2 = torch.relu(xh2 * W.h("EDEC3_G", C) + W.h("EDEC3_BE", C))
    logits = (_fmm(r2, W.h("EDEC3_W", C, 32)) + W.h("EDEC3_B", 32))[:, :ke]
    # ---- bwd_decode ----
    dl = torch.zeros(nh, 32)
    dl[:, :ke] = d_logits
    d = ln_relu_bwd(dl @ W.h("T_EDEC3", 32, C), xh2, rs2, W.h("EDEC3_G", C), W.h("EDEC3_BE", C))
    d = ln_relu_bwd(d @ W.h("T_EDEC2", C, C), xh1, rs1, W.h("EDEC1_G", C), W.h("EDEC1_BE", C))
    ddect = scatter(d, lh, N) + scatter(d, rh, N)
    dhs = d @ W.h("T_EDEC1", C, C)
    dh = torch.zeros(E, C)
    dh[q1] = dhs
    dh[q2] = dhs
    dx = ddect @ W.h("T_EDEC1N", C, D)
    dG = torch.zeros(E, G)
    for i in reversed(range(num_blocks)):
        sv = saved[i]
        xi, e = sv["x"], sv["e"]
        tb = tables(i, xi)
        # phase A: NodeBlock node tail
        xh4, rs4 = ln_stats(tb["cen"] + sv["agg"])
        dc = ln_relu_bwd(dx @ W.b(i, "T_NB_OUT", D, D), xh4, rs4, W.b(i, "NB_LN_G", D), W.b(i, "NB_LN_BE", D))
        dagg = dc
        dx = dx + dc @ W.b(i, "T_NB_CEN", D, D)
        # bwd_edge_tail
        xh7, rs7 = ln_stats(e @ W.b(i, "EB_SELF_W", C, C) + W.b(i, "EB_SELF_B", C) + sv["SL"][left] + sv["SR"][right]
                            + tb["fl"][left] + tb["fr"][right])
        du = ln_relu_bwd(dh @ W.b(i, "T_EB_OUT", C, C), xh7, rs7, W.b(i, "EB_LN_G", C), W.b(i, "EB_LN_BE", C))
        DUL, DUR = scatter(du, left, N), scatter(du, right, N)
        de = du @ W.b(i, "T_EB_SELF", C, C) + dh
        # bwd_edge_nodeblock
        xh2_, rs2_ = ln_stats(_mm(e, W.b(i, "NB_EN1_W", C, D), False) + W.b(i, "NB_EN1_B", D))
        r2_ = torch.relu(xh2_ * W.b(i, "NB_EN1_G", D) + W.b(i, "NB_EN1_BE", D))
        he = _mm(r2_, W.b(i, "NB_EN2_W", D, D), False) + W.b(i, "NB_EN2_B", D)
        msg = _mm((he * tb["hn"][right]), W.b(i, "NB_MSG_W", D, D), False) + W.b(i, "NB_MSG_B", D)
        xh3, rs3 = ln_stats(_mm(e, W.b(i, "NB_GE_W", C, D), False) + tb["gx"][right])
        r3 = torch.relu(xh3 * W.b(i, "NB_G1_G", D) + W.b(i, "NB_G1_BE", D))
        sg = torch.sigmoid(_mm(r3, W.b(i, "NB_G2_W", D, D), False) + W.b(i, "NB_G2_B", D))
        dout = dagg[left]
        dgt, dmsg = dout * msg * sg * (1 - sg), dout * sg
        da3 = ln_relu_bwd(_mm(dgt, W.b(i, "T_NB_G2", D, D), True), xh3, rs3, W.b(i, "NB_G1_G", D), W.b(i, "NB_G1_BE", D))
        dgx = scatter(da3, right, N)
        de = de + _mm(da3, W.b(i, "T_NB_GE", D, C), True)
        dm = _mm(dmsg, W.b(i, "T_NB_MSG", D, D), True)
        dhn = scatter(dm * he, right, N)
        da2 = ln_relu_bwd(_mm((dm * tb["hn"][right]), W.b(i, "T_NB_EN2", D, D), True), xh2_, rs2_, W.b(i, "NB_EN1_G", D), W.b(i, "NB_EN1_BE", D))
        de = de + _mm(da2, W.b(i, "T_NB_EN1", D, C), True)
        # bwd_edge_bondffn
        dnl, dgn = {}, {}
        for s, S, node, other, DU in (("l", "EL", left, right, DUL), ("r", "ER", right, left, DUR)):
            bl = _mm(e, W.b(i, S + "_BL_W", C, 128), False)
            xh5, rs5 = ln_stats(_mm((bl * tb["nl" + s][node]), W.b(i, S + "_I1_W", 128, 128), False) + W.b(i, S + "_I1_B", 128))
            r5 = torch.relu(xh5 * W.b(i, S + "_I1_G", 128) + W.b(i, S + "_I1_BE", 128))
            i2 = _mm(r5, W.b(i, S + "_I2_W", 128, C), False) + W.b(i, S + "_I2_B", C)
            xh6, rs6 = ln_stats(_mm(e, W.b(i, S + "_GB_W", C, 32), False) + tb["gn" + s][node] + te[:, None] * W.b(i, S + "_GT_W", 32))
            r6 = torch.relu(xh6 * W.b(i, S + "_G1_G", 32) + W.b(i, S + "_G1_BE", 32))
            sgg = torch.sigmoid(_mm(r6, W.b(i, S + "_G2_W", 32, C), False) + W.b(i, S + "_G2_B", C))
            do = DU[other]
            dgg, di2 = do * i2 * sgg * (1 - sgg), do * sgg
            da6 = ln_relu_bwd(_mm(dgg, W.b(i, "T_" + S + "_G2", C, 32), True), xh6, rs6, W.b(i, S + "_G1_G", 32), W.b(i, S + "_G1_BE", 32))
            dgn[s] = scatter(da6, node, N)
            de = de + _mm(da6, W.b(i, "T_" + S + "_GB", 32, C), True)
            da5 = ln_relu_bwd(_mm(di2, W.b(i, "T_" + S + "_I2", C, 128), True), xh5, rs5, W.b(i, S + "_I1_G", 128), W.b(i, S + "_I1_BE", 128))
            dinter = _mm(da5, W.b(i, "T_" + S + "_I1", 128, 128), True)
            dnl[s] = scatter(dinter * bl, node, N)
            de = de + _mm((dinter * tb["nl" + s][node]), W.b(i, "T_" + S + "_BL", 128, C), True)
        dh = _mm(de, W.b(i, "T_EEH", C, C), True)
        dG = dG + _mm(de, W.b(i, "T_EEG", C, 32), True)[:, :G]
        # phase B
        dx = dx + DUL @ W.b(i, "T_EB_NFL", C, D) + DUR @ W.b(i, "T_EB_NFR", C, D)
        dx = dx + dnl["l"] @ W.b(i, "T_EL_NL", 128, D) + dnl["r"] @ W.b(i, "T_ER_NL", 128, D)
        dx = dx + dgn["l"] @ W.b(i, "T_EL_GN", 32, D) + dgn["r"] @ W.b(i, "T_ER_GN", 32, D)
        dx = dx + dgx @ W.b(i, "T_NB_GX", D, D)
        xh1_, rs1_ = ln_stats(xi @ W.b(i, "NB_NN1_W", D, D) + W.b(i, "NB_NN1_B", D))
        da1 = ln_relu_bwd(dhn @ W.b(i, "T_NB_NN2", D, D), xh1_, rs1_, W.b(i, "NB_NN1_G", D), W.b(i, "NB_NN1_BE", D))
        dx = dx + da1 @ W.b(i, "T_NB_NN1", D, D)
    # bwd_pos
    u = dist[:, None] - roff[None, :]
    dgdd = torch.exp(rco * (u * u)) * 2.0 * rco * u
    dd = (dG * dgdd).sum(-1)
    mask = (dist >= rbf_lo) & (dist <= rbf_hi) & (dist > 0)
    s = torch.where(mask, dd / dist, torch.zeros_like(dd))
    d_pos = scatter(s[:, None] * rel, left, N) - scatter(s[:, None] * rel, right, N)
    return logits, d_pos
