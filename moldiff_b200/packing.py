"""Weight packing: reference-schema ``state_dict`` -> one fp32 blob in the kernel layout.

The slot lists are parsed from ``include/moldiff_b200.h`` (the X-macro lists are the ABI), so the
Python packer and the CUDA kernels cannot drift apart.  Every matrix is stored ``[K][N]`` row-major
(= ``Linear.weight.T``); first-layer Linears that act on a concatenation ``[edge ; node ; time]`` are
split column-wise so the per-node part can be hoisted out of the per-edge work (exact, by linearity
-- SURVEY.md section 7.1).  State-dict key names cited below are the reference's
(``models/graph.py:12-27,123-131,252-266,378-382``).
"""
from __future__ import annotations

import os
import re

import torch

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "moldiff_b200.h")
NODE_DIM, EDGE_DIM, NUM_RBF, MAX_BLOCKS = 256, 64, 16, 16
_ALIGN = 32  # floats (128 B): keeps every slot 16-byte aligned for cp.async / float4 loads


def _parse_slots(macro):
    src = open(_HEADER).read()
    m = re.search(r"#define\s+" + macro + r"\(X\)(.*?)\n\n", src, re.S)
    if not m:
        raise RuntimeError(f"{macro} not found in {_HEADER}")
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    return re.findall(r"X\((\w+)\)", body)


BLOCK_SLOTS = _parse_slots("MDB_BLOCK_SLOTS")
TC_SLOTS = _parse_slots("MDB_TC_SLOTS")
TC_HEAD_SLOTS = re.findall(r"X\((\w+)\)", re.search(r"#define\s+MDB_TC_HEAD_SLOTS\(X\)(.*)", open(_HEADER).read()).group(1))
HEAD_SLOTS = _parse_slots("MDB_HEAD_SLOTS")


def _t(w):
    """Linear.weight [out][in] -> [in][out] contiguous fp32."""
    return w.detach().to(torch.float32).t().contiguous()


def _pad_cols(m, n):
    out = torch.zeros(m.shape[0], n, dtype=torch.float32)
    out[:, : m.shape[1]] = m
    return out


def _pad_vec(v, n):
    out = torch.zeros(n, dtype=torch.float32)
    out[: v.shape[0]] = v.detach().float()
    return out


def _mlp2(sd, p, tag, out):
    """MLP(in -> hidden -> out): net.0 Linear, net.1 LayerNorm, net.3 Linear (common.py:184-198)."""
    out[f"{tag}1_W"] = _t(sd[f"{p}.net.0.weight"])
    out[f"{tag}1_B"] = sd[f"{p}.net.0.bias"]
    out[f"{tag}1_G"] = sd[f"{p}.net.1.weight"]
    out[f"{tag}1_BE"] = sd[f"{p}.net.1.bias"]
    out[f"{tag}2_W"] = _t(sd[f"{p}.net.3.weight"])
    out[f"{tag}2_B"] = sd[f"{p}.net.3.bias"]


def _bond_ffn(sd, p, tag, out, bond_dim, node_dim):
    """BondFFN (graph.py:123-131); gate.net.0 acts on cat[bond(bond_dim) ; node(node_dim) ; time(1)]."""
    out[f"{tag}_BL_W" if tag != "PU" else "PU_PB_W"] = _t(sd[f"{p}.bond_linear.weight"])
    out[f"{tag}_NL_W" if tag != "PU" else "PU_PN_W"] = _t(sd[f"{p}.node_linear.weight"])
    out[f"{tag}_I1_W"] = _t(sd[f"{p}.inter_module.net.0.weight"])
    out[f"{tag}_I1_B"] = sd[f"{p}.inter_module.net.0.bias"]
    out[f"{tag}_I1_G"] = sd[f"{p}.inter_module.net.1.weight"]
    out[f"{tag}_I1_BE"] = sd[f"{p}.inter_module.net.1.bias"]
    out[f"{tag}_I2_W"] = _t(sd[f"{p}.inter_module.net.3.weight"])
    out[f"{tag}_I2_B"] = sd[f"{p}.inter_module.net.3.bias"]
    g0 = sd[f"{p}.gate.net.0.weight"]  # [32][bond_dim + node_dim + 1]
    out[f"{tag}_GB_W"] = _t(g0[:, :bond_dim])
    out[f"{tag}_GN_W"] = _t(g0[:, bond_dim:bond_dim + node_dim])
    out[f"{tag}_GT_W"] = g0[:, bond_dim + node_dim]
    out[f"{tag}_G1_B"] = sd[f"{p}.gate.net.0.bias"]
    out[f"{tag}_G1_G"] = sd[f"{p}.gate.net.1.weight"]
    out[f"{tag}_G1_BE"] = sd[f"{p}.gate.net.1.bias"]
    out[f"{tag}_G2_W"] = _t(sd[f"{p}.gate.net.3.weight"])
    out[f"{tag}_G2_B"] = sd[f"{p}.gate.net.3.bias"]


def _asis(w):
    """Linear.weight [out][in] kept as stored: the [K][N] operand of dX = dY @ W."""
    return w.detach().to(torch.float32).contiguous()


def _pad_rows(m, n):
    out = torch.zeros(n, m.shape[1], dtype=torch.float32)
    out[: m.shape[0]] = m
    return out


def block_backward_tensors(sd, net_prefix, i):
    """Transposed-use copies for the input-gradient backward kernels (bond predictor guidance)."""
    o = {}
    ee = sd[f"{net_prefix}.edge_embs.{i}.weight"]                  # [64][80]
    o["T_EEH"] = _asis(ee[:, :EDGE_DIM])
    o["T_EEG"] = _pad_cols(_asis(ee[:, EDGE_DIM:]), 32)
    nb = f"{net_prefix}.node_blocks_with_edge.{i}"
    o["T_NB_NN1"] = _asis(sd[nb + ".node_net.net.0.weight"])
    o["T_NB_NN2"] = _asis(sd[nb + ".node_net.net.3.weight"])
    o["T_NB_EN1"] = _asis(sd[nb + ".edge_net.net.0.weight"])
    o["T_NB_EN2"] = _asis(sd[nb + ".edge_net.net.3.weight"])
    o["T_NB_MSG"] = _asis(sd[nb + ".msg_net.weight"])
    g0 = sd[nb + ".gate.net.0.weight"]
    o["T_NB_GE"] = _asis(g0[:, :EDGE_DIM])
    o["T_NB_GX"] = _asis(g0[:, EDGE_DIM:EDGE_DIM + NODE_DIM])
    o["T_NB_G2"] = _asis(sd[nb + ".gate.net.3.weight"])
    o["T_NB_CEN"] = _asis(sd[nb + ".centroid_lin.weight"])
    o["T_NB_OUT"] = _asis(sd[nb + ".out_transform.weight"])
    eb = f"{net_prefix}.edge_blocks.{i}"
    for tag, sub in (("EL", "bond_ffn_left"), ("ER", "bond_ffn_right")):
        p = f"{eb}.{sub}"
        o[f"T_{tag}_BL"] = _asis(sd[p + ".bond_linear.weight"])
        o[f"T_{tag}_NL"] = _asis(sd[p + ".node_linear.weight"])
        o[f"T_{tag}_I1"] = _asis(sd[p + ".inter_module.net.0.weight"])
        o[f"T_{tag}_I2"] = _asis(sd[p + ".inter_module.net.3.weight"])
        g0 = sd[p + ".gate.net.0.weight"]
        o[f"T_{tag}_GB"] = _asis(g0[:, :EDGE_DIM])
        o[f"T_{tag}_GN"] = _asis(g0[:, EDGE_DIM:EDGE_DIM + NODE_DIM])
        o[f"T_{tag}_G2"] = _asis(sd[p + ".gate.net.3.weight"])
    o["T_EB_NFL"] = _asis(sd[eb + ".node_ffn_left.weight"])
    o["T_EB_NFR"] = _asis(sd[eb + ".node_ffn_right.weight"])
    o["T_EB_SELF"] = _asis(sd[eb + ".self_ffn.weight"])
    o["T_EB_OUT"] = _asis(sd[eb + ".out_transform.weight"])
    return o


def block_tensors(sd, net_prefix, i, update_pos):
    """All slot tensors of block i, keyed by slot name."""
    o = {}
    o["EE_W"] = _t(sd[f"{net_prefix}.edge_embs.{i}.weight"])        # [80][64]
    o["EE_B"] = sd[f"{net_prefix}.edge_embs.{i}.bias"]
    nb = f"{net_prefix}.node_blocks_with_edge.{i}"
    _mlp2(sd, nb + ".node_net", "NB_NN", o)
    _mlp2(sd, nb + ".edge_net", "NB_EN", o)
    o["NB_MSG_W"] = _t(sd[nb + ".msg_net.weight"])
    o["NB_MSG_B"] = sd[nb + ".msg_net.bias"]
    g0 = sd[nb + ".gate.net.0.weight"]                              # [256][64 + 256 + 1]   graph.py:22,46
    o["NB_GE_W"] = _t(g0[:, :EDGE_DIM])
    o["NB_GX_W"] = _t(g0[:, EDGE_DIM:EDGE_DIM + NODE_DIM])
    o["NB_GT_W"] = g0[:, EDGE_DIM + NODE_DIM]
    o["NB_G1_B"] = sd[nb + ".gate.net.0.bias"]
    o["NB_G1_G"] = sd[nb + ".gate.net.1.weight"]
    o["NB_G1_BE"] = sd[nb + ".gate.net.1.bias"]
    o["NB_G2_W"] = _t(sd[nb + ".gate.net.3.weight"])
    o["NB_G2_B"] = sd[nb + ".gate.net.3.bias"]
    o["NB_CEN_W"] = _t(sd[nb + ".centroid_lin.weight"])
    o["NB_CEN_B"] = sd[nb + ".centroid_lin.bias"]
    o["NB_LN_G"] = sd[nb + ".layer_norm.weight"]
    o["NB_LN_BE"] = sd[nb + ".layer_norm.bias"]
    o["NB_OUT_W"] = _t(sd[nb + ".out_transform.weight"])
    o["NB_OUT_B"] = sd[nb + ".out_transform.bias"]
    eb = f"{net_prefix}.edge_blocks.{i}"
    _bond_ffn(sd, eb + ".bond_ffn_left", "EL", o, EDGE_DIM, NODE_DIM)
    _bond_ffn(sd, eb + ".bond_ffn_right", "ER", o, EDGE_DIM, NODE_DIM)
    o["EB_NFL_W"] = _t(sd[eb + ".node_ffn_left.weight"])
    o["EB_NFL_B"] = sd[eb + ".node_ffn_left.bias"]
    o["EB_NFR_W"] = _t(sd[eb + ".node_ffn_right.weight"])
    o["EB_NFR_B"] = sd[eb + ".node_ffn_right.bias"]
    o["EB_SELF_W"] = _t(sd[eb + ".self_ffn.weight"])
    o["EB_SELF_B"] = sd[eb + ".self_ffn.bias"]
    o["EB_LN_G"] = sd[eb + ".layer_norm.weight"]
    o["EB_LN_BE"] = sd[eb + ".layer_norm.bias"]
    o["EB_OUT_W"] = _t(sd[eb + ".out_transform.weight"])
    o["EB_OUT_B"] = sd[eb + ".out_transform.bias"]
    if update_pos:
        pb = f"{net_prefix}.pos_blocks.{i}"
        _mlp2(sd, pb + ".left_lin_edge", "PU_LL", o)
        _mlp2(sd, pb + ".right_lin_edge", "PU_RL", o)
        _bond_ffn(sd, pb + ".edge_lin", "PU", o, EDGE_DIM, EDGE_DIM)
        o["PU_I2_W"] = o["PU_I2_W"].reshape(-1)      # Linear(256 -> 1): used as a dot product
        o["PU_G2_W"] = o["PU_G2_W"].reshape(-1)      # Linear(32 -> 1)
    return o


def head_tensors(sd, kind, net_prefix, time_dim):
    o = {"RBF_OFFSET": sd[f"{net_prefix}.distance_expansion.offset"],
         "RBF_COEFF": sd[f"{net_prefix}.distance_expansion.coeff"]}
    if kind == 0:
        return o
    tprefix = "time_emb.0" if kind == 1 else "time_emb"          # model.py:34-36 vs bond_predictor.py:31
    if time_dim > 0:
        o["TIME_OFFSET"] = sd[tprefix + ".offset"]
        o["TIME_COEFF"] = sd[tprefix + ".coeff"]
    o["NODE_EMB_W"] = _t(sd["node_embedder.weight"])             # [Kn][256 - time_dim]
    o["EDGE_EMB_W"] = _t(sd["edge_embedder.weight"])             # [Ke | 2Kn][64 - time_dim]
    if kind == 1:
        _mlp2(sd, "node_decoder", "NDEC", o)
        o["NDEC2_W"] = _pad_cols(o["NDEC2_W"], 32)
        o["NDEC2_B"] = _pad_vec(o["NDEC2_B"], 32)
        _mlp2(sd, "edge_decoder", "EDEC", o)
        o["EDEC2_W"] = _pad_cols(o["EDEC2_W"], 32)
        o["EDEC2_B"] = _pad_vec(o["EDEC2_B"], 32)
    else:
        # edge_decoder = MLP(64 + 256 -> 64 -> 64 -> K, num_layer=3)      bond_predictor.py:34
        w0 = sd["edge_decoder.net.0.weight"]                     # [64][320]
        o["EDEC1_W"] = _t(w0[:, :EDGE_DIM])
        o["EDEC1N_W"] = _t(w0[:, EDGE_DIM:])
        o["EDEC1_B"] = sd["edge_decoder.net.0.bias"]
        o["EDEC1_G"] = sd["edge_decoder.net.1.weight"]
        o["EDEC1_BE"] = sd["edge_decoder.net.1.bias"]
        o["EDEC2_W"] = _t(sd["edge_decoder.net.3.weight"])
        o["EDEC2_B"] = sd["edge_decoder.net.3.bias"]
        o["EDEC3_G"] = sd["edge_decoder.net.4.weight"]
        o["EDEC3_BE"] = sd["edge_decoder.net.4.bias"]
        o["EDEC3_W"] = _pad_cols(_t(sd["edge_decoder.net.6.weight"]), 32)
        o["EDEC3_B"] = _pad_vec(sd["edge_decoder.net.6.bias"], 32)
        o["T_EDEC1"] = _asis(w0[:, :EDGE_DIM])
        o["T_EDEC1N"] = _asis(w0[:, EDGE_DIM:])
        o["T_EDEC2"] = _asis(sd["edge_decoder.net.3.weight"])
        o["T_EDEC3"] = _pad_rows(_asis(sd["edge_decoder.net.6.weight"]), 32)
    return o


def pack_network(sd, *, kind, net_prefix, num_blocks, update_pos, time_dim=0):
    """Returns (blob fp32 1-D CPU tensor, head_off list[int], block_off list[list[int]]); -1 = absent."""
    if num_blocks > MAX_BLOCKS:
        raise ValueError(f"num_blocks {num_blocks} > {MAX_BLOCKS}")
    chunks, cursor = [], 0

    def put(tensor):
        nonlocal cursor
        flat = tensor.detach().to(torch.float32).reshape(-1).cpu()
        off = cursor
        pad = (-flat.numel()) % _ALIGN
        chunks.append(flat)
        if pad:
            chunks.append(torch.zeros(pad, dtype=torch.float32))
        cursor += flat.numel() + pad
        return off

    head = head_tensors(sd, kind, net_prefix, time_dim)
    head_off = [put(head[s]) if s in head else -1 for s in HEAD_SLOTS]
    block_off = []
    for i in range(num_blocks):
        bt = block_tensors(sd, net_prefix, i, update_pos)
        if kind == 2:
            bt.update(block_backward_tensors(sd, net_prefix, i))
        unknown = set(bt) - set(BLOCK_SLOTS)
        if unknown:
            raise RuntimeError(f"packer produced unknown slots {sorted(unknown)}")
        block_off.append([put(bt[s]) if s in bt else -1 for s in BLOCK_SLOTS])
    return torch.cat(chunks), head_off, block_off


# ---------------------------------------------------------------------------------------------------------
# tensor-core operand images (tcgen05 path)
# ---------------------------------------------------------------------------------------------------------
TC_KB = 16   # K columns per weight stage; must equal tc::KB in csrc/tc_pipe.cuh


TC_OPERAND_DTYPE = torch.float16      # must match tc::OPERAND_FMT in csrc/tc_engine.cuh (0 = fp16, 1 = bf16)
TC_ACC_SCALE = 256.0                  # must match tc::ACC_SCALE: images hold 256 x W so that fp16 lo planes stay normal


def split_bf16(w):
    """fp32 -> (hi, lo) 16-bit pair, round-to-nearest-even both times.  With fp16 planes hi + lo carries 22 significant
    bits (|w| < 6e-5 goes subnormal: absolute error <= 3e-8); with bf16 planes 16 bits over fp32's range."""
    hi = w.to(TC_OPERAND_DTYPE)
    lo = (w - hi.to(torch.float32)).to(TC_OPERAND_DTYPE)
    return hi, lo


def _canonical_plane(wt_stage):
    """[N][KB] bf16 (the 'N x K, K-major' B operand of one stage) -> UMMA SWIZZLE_NONE K-major canonical image:
    byte offset of (n, k) = (n % 8) * 16 + (n / 8) * (KB / 8) * 128 + (k / 8) * 128 + (k % 8) * 2."""
    n, kb = wt_stage.shape
    return wt_stage.reshape(n // 8, 8, kb // 8, 8).permute(0, 2, 1, 3).contiguous().reshape(-1)


def tc_image(w_kn):
    """W[K][N] fp32 (y = x @ W) -> int16 tensor holding, for every K stage, [hi plane | lo plane]."""
    k, n = w_kn.shape
    if k % TC_KB or n % 16:
        raise ValueError(f"tc_image: K={k} must be a multiple of {TC_KB} and N={n} of 16")
    w = w_kn.detach().to(torch.float32).cpu().t().contiguous() * TC_ACC_SCALE     # [N][K]
    if TC_OPERAND_DTYPE == torch.float16:
        w = w.clamp(-65504.0, 65504.0)              # |weight| > 255 saturates like the kernels' cvt.satfinite
    hi, lo = split_bf16(w)
    parts = []
    for s in range(k // TC_KB):
        sl = slice(s * TC_KB, (s + 1) * TC_KB)
        parts.append(_canonical_plane(hi[:, sl]).view(torch.int16))
        parts.append(_canonical_plane(lo[:, sl]).view(torch.int16))
    return torch.cat(parts)


def tc_block_tensors(sd, net_prefix, i, update_pos, with_backward):
    """[K][N] fp32 matrices (y = x @ W) of block i that run on tensor cores, keyed by MDB_TC_SLOTS name."""
    nb = f"{net_prefix}.node_blocks_with_edge.{i}"
    g0 = sd[nb + ".gate.net.0.weight"]
    o = {
        "NB_EN1": _t(sd[nb + ".edge_net.net.0.weight"]),       # [64][256]
        "NB_EN2": _t(sd[nb + ".edge_net.net.3.weight"]),       # [256][256]
        "NB_MSG": _t(sd[nb + ".msg_net.weight"]),
        "NB_GE": _t(g0[:, :EDGE_DIM]),                         # [64][256]
        "NB_G2": _t(sd[nb + ".gate.net.3.weight"]),
    }
    o["EE"] = _t(sd[f"{net_prefix}.edge_embs.{i}.weight"])     # [80][64]
    eb = f"{net_prefix}.edge_blocks.{i}"
    for tag, sub in (("EL", "bond_ffn_left"), ("ER", "bond_ffn_right")):
        p = f"{eb}.{sub}"
        gw = sd[p + ".gate.net.0.weight"]
        o[f"{tag}_BL"] = _t(sd[p + ".bond_linear.weight"])                 # [64][128]
        o[f"{tag}_GB"] = _t(gw[:, :EDGE_DIM])                              # [64][32]
        o[f"{tag}_I1"] = _t(sd[p + ".inter_module.net.0.weight"])          # [128][128]
        o[f"{tag}_G2"] = _t(sd[p + ".gate.net.3.weight"])                  # [32][64]
        o[f"{tag}_I2"] = _t(sd[p + ".inter_module.net.3.weight"])          # [128][64]
    # per-node Linears (tc_node_kernel)
    o["NB_OUT"] = _t(sd[nb + ".out_transform.weight"])
    o["NB_NN1"] = _t(sd[nb + ".node_net.net.0.weight"])
    o["NB_NN2"] = _t(sd[nb + ".node_net.net.3.weight"])
    o["NB_GX"] = _t(g0[:, EDGE_DIM:EDGE_DIM + NODE_DIM])
    o["NB_CEN"] = _t(sd[nb + ".centroid_lin.weight"])
    for tag, sub in (("EL", "bond_ffn_left"), ("ER", "bond_ffn_right")):
        p = f"{eb}.{sub}"
        gw = sd[p + ".gate.net.0.weight"]
        o[f"{tag}_NL"] = _t(sd[p + ".node_linear.weight"])                 # [256][128]
        o[f"{tag}_GN"] = _t(gw[:, EDGE_DIM:EDGE_DIM + NODE_DIM])           # [256][32]
    o["EB_NFL"] = _t(sd[eb + ".node_ffn_left.weight"])         # [256][64]
    o["EB_NFR"] = _t(sd[eb + ".node_ffn_right.weight"])
    if update_pos:
        pp = f"{net_prefix}.pos_blocks.{i}"
        o["PU_LL1"] = _t(sd[pp + ".left_lin_edge.net.0.weight"])   # [256][64]
        o["PU_LL2"] = _t(sd[pp + ".left_lin_edge.net.3.weight"])   # [64][64]
        o["PU_RL1"] = _t(sd[pp + ".right_lin_edge.net.0.weight"])
        o["PU_RL2"] = _t(sd[pp + ".right_lin_edge.net.3.weight"])
    o["EB_SELF"] = _t(sd[eb + ".self_ffn.weight"])             # [64][64]
    o["EB_OUT"] = _t(sd[eb + ".out_transform.weight"])         # [64][64]
    if with_backward:                                          # dX = dY @ W  with W stored [out][in] = [K][N]
        o["BT_NB_G2"] = _asis(sd[nb + ".gate.net.3.weight"])
        o["BT_NB_GE"] = _asis(g0[:, :EDGE_DIM])                # [256][64]
        o["BT_NB_MSG"] = _asis(sd[nb + ".msg_net.weight"])
        o["BT_NB_EN2"] = _asis(sd[nb + ".edge_net.net.3.weight"])
        o["BT_NB_EN1"] = _asis(sd[nb + ".edge_net.net.0.weight"])   # [256][64]
        ee = sd[f"{net_prefix}.edge_embs.{i}.weight"]               # [64][80]
        o["BT_EEH"] = _asis(ee[:, :EDGE_DIM])                       # [64][64]
        o["BT_EEG"] = _pad_cols(_asis(ee[:, EDGE_DIM:]), 32)        # [64][16 -> 32]
        for tag, sub in (("EL", "bond_ffn_left"), ("ER", "bond_ffn_right")):
            p = f"{eb}.{sub}"
            gw = sd[p + ".gate.net.0.weight"]
            o[f"BT_{tag}_G2"] = _asis(sd[p + ".gate.net.3.weight"])            # [64][32]
            o[f"BT_{tag}_I2"] = _asis(sd[p + ".inter_module.net.3.weight"])    # [64][128]
            o[f"BT_{tag}_GB"] = _asis(gw[:, :EDGE_DIM])                        # [32][64]
            o[f"BT_{tag}_I1"] = _asis(sd[p + ".inter_module.net.0.weight"])    # [128][128]
            o[f"BT_{tag}_BL"] = _asis(sd[p + ".bond_linear.weight"])           # [128][64]
            o[f"BT_{tag}_NL"] = _asis(sd[p + ".node_linear.weight"])           # [128][256]
            o[f"BT_{tag}_GN"] = _asis(gw[:, EDGE_DIM:EDGE_DIM + NODE_DIM])     # [32][256]
        o["BT_EB_OUT"] = _asis(sd[eb + ".out_transform.weight"])               # [64][64]
        o["BT_EB_SELF"] = _asis(sd[eb + ".self_ffn.weight"])
        o["BT_EB_NFL"] = _asis(sd[eb + ".node_ffn_left.weight"])               # [64][256]
        o["BT_EB_NFR"] = _asis(sd[eb + ".node_ffn_right.weight"])
        o["BT_NB_GX"] = _asis(g0[:, EDGE_DIM:EDGE_DIM + NODE_DIM])             # [256][256]
        o["BT_NB_NN2"] = _asis(sd[nb + ".node_net.net.3.weight"])
        o["BT_NB_NN1"] = _asis(sd[nb + ".node_net.net.0.weight"])
        o["BT_NB_OUT"] = _asis(sd[nb + ".out_transform.weight"])
        o["BT_NB_CEN"] = _asis(sd[nb + ".centroid_lin.weight"])
    if update_pos:
        pb = f"{net_prefix}.pos_blocks.{i}.edge_lin"
        o["PU_PB"] = _t(sd[pb + ".bond_linear.weight"])        # [64][256]
        o["PU_PN"] = _t(sd[pb + ".node_linear.weight"])        # [64][256]
        o["PU_I1"] = _t(sd[pb + ".inter_module.net.0.weight"]) # [256][256]
        pg = sd[pb + ".gate.net.0.weight"]                     # [32][64 + 64 + 1]
        o["PU_GB"] = _t(pg[:, :EDGE_DIM])                      # [64][32]
        o["PU_GN"] = _t(pg[:, EDGE_DIM:2 * EDGE_DIM])          # [64][32]
    return o


def tc_head_tensors(sd, kind):
    if kind == 1:
        return {"NDEC1": _t(sd["node_decoder.net.0.weight"]),
                "NDEC2": _pad_cols(_t(sd["node_decoder.net.3.weight"]), 32)}
    if kind == 2:
        return {"EDEC1N": _t(sd["edge_decoder.net.0.weight"][:, EDGE_DIM:]),       # [256][64]
                "BT_EDEC1N": _asis(sd["edge_decoder.net.0.weight"][:, EDGE_DIM:])}  # [64][256]  (backward: dx = ddect W)
    return {}


def pack_tc(sd, *, net_prefix, num_blocks, update_pos, with_backward, kind=0):
    """Returns (int16 1-D CPU tensor, block_off list[list[int]] in BYTES, head_off list[int]); -1 = absent;
    images 128-byte aligned."""
    chunks, cursor, offs = [], 0, []
    for i in range(num_blocks):
        bt = tc_block_tensors(sd, net_prefix, i, update_pos, with_backward)
        row = []
        for name in TC_SLOTS:
            if name not in bt:
                row.append(-1)
                continue
            img = tc_image(bt[name])
            pad = (-img.numel()) % 64                          # int16 elements -> 128 bytes
            row.append(cursor * 2)
            chunks.append(img)
            if pad:
                chunks.append(torch.zeros(pad, dtype=torch.int16))
            cursor += img.numel() + pad
        offs.append(row)
    ht = tc_head_tensors(sd, kind)
    head = []
    for name in TC_HEAD_SLOTS:
        if name not in ht:
            head.append(-1)
            continue
        img = tc_image(ht[name])
        pad = (-img.numel()) % 64
        head.append(cursor * 2)
        chunks.append(img)
        if pad:
            chunks.append(torch.zeros(pad, dtype=torch.int16))
        cursor += img.numel() + pad
    return torch.cat(chunks), offs, head
