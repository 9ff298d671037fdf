"""Host binding of the C-ABI (include/moldiff_b200.h) -- ctypes, raw device pointers, no torch types
across the boundary.  PyTorch is used here only for device memory, streams and the once-per-batch sort.

There is deliberately NO fallback: if ``libmoldiff_b200.so`` is missing or the tensors are not on a
CUDA device, every entry point raises.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
import threading

import torch

from . import packing

# MDB_LIB_VARIANT=<name> loads libmoldiff_b200_<name>.so, an A/B build made by moldiff_b200.build.build(variant=...)
# (numerics / tuning studies only; the default is the one product library)
_VARIANT = os.environ.get("MDB_LIB_VARIANT", "")
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                         f"libmoldiff_b200{'_' + _VARIANT if _VARIANT else ''}.so")
_lib = None
_lib_lock = threading.Lock()

NUM_HEAD = len(packing.HEAD_SLOTS)
NUM_BLOCK = len(packing.BLOCK_SLOTS)
NUM_TC = len(packing.TC_SLOTS)
NUM_TC_HEAD = len(packing.TC_HEAD_SLOTS)
MAX_BLOCKS = packing.MAX_BLOCKS


class NetDesc(C.Structure):
    _fields_ = [
        ("blob", C.c_void_p),
        ("num_blocks", C.c_int32),
        ("update_pos", C.c_int32),
        ("rbf_start", C.c_float),
        ("rbf_stop", C.c_float),
        ("time_dim", C.c_int32),
        ("num_node_types", C.c_int32),
        ("num_edge_types", C.c_int32),
        ("num_timesteps", C.c_float),
        ("kind", C.c_int32),
        ("head_off", C.c_int64 * NUM_HEAD),
        ("block_off", (C.c_int64 * NUM_BLOCK) * MAX_BLOCKS),
        ("blob_host", C.c_void_p),
        ("tc_blob", C.c_void_p),
        ("tc_block_off", (C.c_int64 * NUM_TC) * MAX_BLOCKS),
        ("tc_head_off", C.c_int64 * NUM_TC_HEAD),
    ]


class Plan(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("n_edges", C.c_int32),
        ("n_half", C.c_int32),
        ("left", C.c_void_p),
        ("right", C.c_void_p),
        ("perm", C.c_void_p),
        ("inv", C.c_void_p),
    ]


class MoldiffB200Error(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load_library():
    """dlopen the in-tree shared library and declare its prototypes.  Raises if it is not built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise MoldiffB200Error(
                f"{_LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "moldiff_b200 has no CPU / eager fallback")
        lib = C.CDLL(_LIB_PATH)
        vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
        lib.mdb_workspace_bytes.restype = sz
        lib.mdb_workspace_bytes.argtypes = [i64, i64, i32, i32]
        lib.mdb_net_forward.restype = C.c_int
        lib.mdb_net_forward.argtypes = [C.POINTER(NetDesc), C.POINTER(Plan)] + [vp] * 9 + [sz, vp]
        lib.mdb_moldiff_forward.restype = C.c_int
        lib.mdb_moldiff_forward.argtypes = [C.POINTER(NetDesc), C.POINTER(Plan)] + [vp] * 10 + [sz, vp]
        lib.mdb_bondpred_forward.restype = C.c_int
        lib.mdb_bondpred_forward.argtypes = [C.POINTER(NetDesc), C.POINTER(Plan)] + [vp] * 6 + [i32, vp, sz, vp]
        lib.mdb_bondpred_backward.restype = C.c_int
        lib.mdb_bondpred_backward.argtypes = [C.POINTER(NetDesc), C.POINTER(Plan)] + [vp] * 8 + [sz, vp]
        lib.mdb_tc_selftest.restype = C.c_int
        lib.mdb_tc_selftest.argtypes = [vp, vp, vp, i32, i32, i32, vp]
        lib.mdb_decode_rows.restype = C.c_int
        lib.mdb_decode_rows.argtypes = [i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp]
        lib.mdb_radius_graph.restype = C.c_int
        lib.mdb_radius_graph.argtypes = [i32, vp, vp, vp, f32, i32, i32, vp, vp, vp]
        lib.mdb_knn_graph.restype = C.c_int
        lib.mdb_knn_graph.argtypes = [i32, vp, vp, vp, i32, i32, vp, vp, vp]
        lib.mdb_operand_amax.restype = C.c_int
        lib.mdb_operand_amax.argtypes = [i64, i64, vp, vp, vp]
        lib.mdb_transition_step.restype = C.c_int
        lib.mdb_transition_step.argtypes = [i32, i32, i32, i32] + [vp] * 26
        lib.mdb_profile_begin.restype = None
        lib.mdb_profile_end.restype = C.c_int
        lib.mdb_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        lib.mdb_kernel_class_name.restype = C.c_char_p
        lib.mdb_kernel_class_name.argtypes = [C.c_int]
        lib.mdb_num_kernel_classes.restype = C.c_int
        lib.mdb_last_error.restype = C.c_char_p
        lib.mdb_version.restype = C.c_int
        lib.mdb_launch_count.restype = i64
        _lib = lib
        return lib


def launch_count():
    return int(load_library().mdb_launch_count())


# As-written (reference) GEMM FLOPs per directed edge executed inside each edge kernel class per launch
# (SURVEY.md 8d / BASELINE.md: edge_emb 10 240 + NodeBlock edge part 2*(64*256*2 + 256*256*3 + ...) etc.).
KERNEL_LOGICAL_FLOP_PER_EDGE = {
    # edge_embs (80->64) + NodeBlock per-edge Linears as written (edge_net 64->256->256, msg 256->256,
    # gate 321->256->256) + two BondFFNs as written (bond 64->128, node 256->128, inter 128->128->64, gate 321->32->64)
    "edge_b": 2.0 * (80 * 64 + 64 * 256 + 256 * 256 + 256 * 256 + 321 * 256 + 256 * 256
                     + 2 * (64 * 128 + 256 * 128 + 128 * 128 + 128 * 64 + 321 * 32 + 32 * 64)),
    # EdgeBlock tail (node_ffn L/R 256->64, self 64->64, out 64->64) + PosUpdate as written
    "edge_d": 2.0 * (2 * 256 * 64 + 64 * 64 + 64 * 64
                     + 2 * (256 * 64 + 64 * 64) + 64 * 256 + 64 * 256 + 256 * 256 + 256 + 129 * 32 + 32),
    # input-gradient backward of the NodeBlock per-edge Linears as autograd executes them (one dX = dY W per Linear)
    "bwd_edge_nodeblock": 2.0 * (64 * 256 + 256 * 256 + 256 * 256 + 321 * 256 + 256 * 256),
    # tensor-core kernels: same reference Linears (as written), executed as 3 split-fp16 MMAs per product after hoisting
    "tc_nodeblock": 2.0 * (64 * 256 + 256 * 256 + 256 * 256 + 321 * 256 + 256 * 256),
    "tc_nodeblock_bwd": 2.0 * (64 * 256 + 256 * 256 + 256 * 256 + 321 * 256 + 256 * 256),
    "tc_bondffn": 2.0 * (80 * 64 + 2 * (64 * 128 + 256 * 128 + 128 * 128 + 128 * 64 + 321 * 32 + 32 * 64)),
    "tc_bondffn_bwd": 2.0 * (80 * 64 + 2 * (64 * 128 + 256 * 128 + 128 * 128 + 128 * 64 + 321 * 32 + 32 * 64)),
    "bwd_edge_bondffn": 2.0 * (80 * 64 + 2 * (64 * 128 + 256 * 128 + 128 * 128 + 128 * 64 + 321 * 32 + 32 * 64)),
}


# The GEMMs each tensor-core kernel EXECUTES per 128-edge tile after per-node hoisting, as (K, N) -- one logical GEMM = 3 MMAs per
# K step (hi*hi + lo*hi + hi*lo).  Executed MMA FLOP per edge = 3 * 2 * sum(K * N)  (roofline.achieved).
_FFN = [(64, 128), (64, 32), (128, 128), (32, 64), (128, 64)]                  # one BondFFN: bond_linear, gate.0, inter.0, gate.3, inter.3
KERNEL_GEMMS = {
    "tc_nodeblock": [(64, 256), (256, 256), (256, 256), (64, 256), (256, 256)],   # edge_net.0/.3, msg_net, gate.0 (edge part), gate.3
    # forward recompute (5, gate.0 edge part twice more for the two LN backwards -> counted below) + 7 backward GEMMs
    "tc_nodeblock_bwd": [(64, 256), (256, 256), (256, 256), (64, 256), (256, 256),
                         (256, 256), (64, 256), (256, 64), (256, 256), (256, 256), (64, 256), (256, 64)],
    "tc_bondffn": [(80, 64)] + _FFN + _FFN,
    "tc_bondffn_bwd": (_FFN + [(64, 32), (64, 128), (32, 64), (128, 128), (128, 64)]) * 2 + [(64, 64), (64, 32)],
    "tc_edge_d": [(64, 64), (64, 64), (64, 256), (64, 256), (256, 256), (64, 32), (64, 32)],   # denoiser (with PosUpdate)
}
KERNEL_MMA_FLOP_PER_EDGE = {k: 3 * 2.0 * sum(a * b for a, b in v) for k, v in KERNEL_GEMMS.items()}


def profile_kernels(fn, reps=1):
    """Run fn() `reps` times with per-kernel CUDA-event timing enabled; returns {class: {ms_total, launches}}."""
    lib = load_library()
    n = lib.mdb_num_kernel_classes()
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    torch.cuda.synchronize()
    lib.mdb_profile_begin()
    try:
        for _ in range(reps):
            fn()
    finally:
        rc = lib.mdb_profile_end(ms, cnt)
    _check(rc, "mdb_profile_end")
    return {lib.mdb_kernel_class_name(i).decode(): {"ms_total": float(ms[i]), "launches": int(cnt[i])}
            for i in range(n) if cnt[i] > 0}


def _check(rc, what):
    if rc != 0:
        msg = load_library().mdb_last_error().decode()
        raise MoldiffB200Error(f"{what} failed (code {rc}): {msg}")


def _dev_f32(t, name):
    if not t.is_cuda:
        raise MoldiffB200Error(f"{name} must be a CUDA tensor: moldiff_b200 has no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _dev_i64(t, name):
    if not t.is_cuda:
        raise MoldiffB200Error(f"{name} must be a CUDA tensor: moldiff_b200 has no CPU path")
    return t.to(torch.int64).contiguous()


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class PackedNet:
    """Device-resident packed weights + the mdb_net_desc that describes them."""

    def __init__(self, state_dict, *, kind, net_prefix, num_blocks, update_pos, cutoff, start=0.0,
                 time_dim=0, num_node_types=0, num_edge_types=0, num_timesteps=1.0, device=None):
        blob, head_off, block_off = packing.pack_network(
            state_dict, kind=kind, net_prefix=net_prefix, num_blocks=num_blocks,
            update_pos=update_pos, time_dim=time_dim)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MoldiffB200Error("PackedNet needs a CUDA device: moldiff_b200 has no CPU path")
        self.blob = blob.to(self.device)
        self.blob_cpu = blob.contiguous()          # host copy: small parameter vectors travel as kernel arguments
        d = NetDesc()
        d.blob = self.blob.data_ptr()
        d.blob_host = self.blob_cpu.data_ptr()
        d.num_blocks = num_blocks
        d.update_pos = int(bool(update_pos))
        d.rbf_start = float(start)
        d.rbf_stop = float(cutoff)
        d.time_dim = int(time_dim)
        d.num_node_types = int(num_node_types)
        d.num_edge_types = int(num_edge_types)
        d.num_timesteps = float(num_timesteps)
        d.kind = int(kind)
        for i, o in enumerate(head_off):
            d.head_off[i] = o
        for b in range(MAX_BLOCKS):
            for s in range(NUM_BLOCK):
                d.block_off[b][s] = block_off[b][s] if b < num_blocks else -1
        # tensor-core operand images (tcgen05 split-fp16 path); MDB_DISABLE_TC=1 keeps the fp32 FFMA kernels
        self.tc_blob = None
        d.tc_blob = None
        for b in range(MAX_BLOCKS):
            for s in range(NUM_TC):
                d.tc_block_off[b][s] = -1
        for s in range(NUM_TC_HEAD):
            d.tc_head_off[s] = -1
        if os.environ.get("MDB_DISABLE_TC", "0") != "1":
            tcb, tco, tch = packing.pack_tc(state_dict, net_prefix=net_prefix, num_blocks=num_blocks,
                                            update_pos=update_pos, with_backward=(kind == 2), kind=kind)
            self.tc_blob = tcb.to(self.device)
            d.tc_blob = self.tc_blob.data_ptr()
            for b in range(num_blocks):
                for s in range(NUM_TC):
                    d.tc_block_off[b][s] = tco[b][s]
            for s in range(NUM_TC_HEAD):
                d.tc_head_off[s] = tch[s]
        self.desc = d
        self.kind = kind
        self.num_blocks = num_blocks


class GraphPlan:
    """CSR edge order for one batch: edges sorted by (left, right).  Constant over all T steps."""

    def __init__(self, edge_index, n_nodes, paired_halves=True):
        if not edge_index.is_cuda:
            raise MoldiffB200Error("edge_index must be a CUDA tensor: moldiff_b200 has no CPU path")
        ei = edge_index.to(torch.int64)
        E = int(ei.shape[1])
        if n_nodes >= 2 ** 31 or E >= 2 ** 31:
            raise MoldiffB200Error("graph too large for int32 indexing")
        key = ei[0] * int(n_nodes) + ei[1]
        perm = torch.argsort(key, stable=True)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(E, device=ei.device)
        self.left = ei[0][perm].to(torch.int32).contiguous()
        self.right = ei[1][perm].to(torch.int32).contiguous()
        self.perm = perm.to(torch.int32).contiguous()
        self.inv = inv.to(torch.int32).contiguous()
        self.n_nodes, self.n_edges = int(n_nodes), E
        self.n_half = E // 2 if (paired_halves and E % 2 == 0) else 0
        self.device = ei.device
        p = Plan()
        p.n_nodes, p.n_edges, p.n_half = self.n_nodes, self.n_edges, self.n_half
        p.left, p.right = self.left.data_ptr(), self.right.data_ptr()
        p.perm, p.inv = self.perm.data_ptr(), self.inv.data_ptr()
        self.c = p
        self._workspace = {}
        self.save_generation = 0     # bumped by every bondpred_forward(save=True): the saved activations live in the workspace

    def workspace(self, with_backward, num_blocks):
        key = (int(with_backward), int(num_blocks))
        ws = self._workspace.get(key)
        if ws is None:
            nbytes = load_library().mdb_workspace_bytes(self.n_nodes, self.n_edges, key[0], key[1])
            # Contract: every accumulator inside the workspace is cleared INSIDE the call that uses it (memsets at the head of
            # run_forward / run_bondpred_backward, the node kernels re-arm the per-block ones), so a call never depends on what
            # the previous one left behind -- tests/test_gpu_parity.py poisons the workspace with NaNs to hold that.  The
            # zero fill only keeps never-read padding rows (node-blocked tables) finite.
            ws = torch.zeros((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
            self._workspace[key] = ws
        return ws


_plan_cache = collections.OrderedDict()
PLAN_CACHE_SIZE = 4        # batches whose plan (sorted edge list + workspaces) stays alive; a training loop alternates a few


def plan_for(edge_index, n_nodes):
    """Plans are cached on the identity + version of the edge_index tensor: MolDiff.sample reuses one
    edge_index for all T steps, so the sort runs once per batch.  Small LRU: interleaved batches (train / validation,
    guided sampling of two batch sizes) keep their plans instead of re-sorting on every call."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(n_nodes), str(edge_index.device))
    hit = _plan_cache.get(key)
    if hit is not None:
        _plan_cache.move_to_end(key)
        return hit[0]
    plan = GraphPlan(edge_index, n_nodes)
    _plan_cache[key] = (plan, edge_index)        # keep edge_index alive so data_ptr cannot be recycled
    while len(_plan_cache) > PLAN_CACHE_SIZE:
        _plan_cache.popitem(last=False)
    return plan


FP16_OPERAND_LIMIT = 65504.0 / 2       # operand planes saturate at 65504; refuse well before that


def check_operand_range(plan: GraphPlan, num_blocks):
    """After a forward on `plan`: max |v| of the residual streams / e / node_net table that feed the split-fp16 operand planes.
    Raises if any of them is not finite or within a factor 2 of the fp16 range (where the planes would saturate silently);
    returns the four values.  One small launch + a 16-byte D2H: MolDiff.sample calls it on its first step only."""
    lib = load_library()
    ws = plan.workspace(0, num_blocks)
    out = torch.zeros(4, dtype=torch.float32, device=plan.device)
    rc = lib.mdb_operand_amax(plan.n_nodes, plan.n_edges, ws.data_ptr(), out.data_ptr(), _stream_ptr(plan.device))
    _check(rc, "mdb_operand_amax")
    vals = dict(zip(("h_node", "h_edge", "e", "node_net"), (float(v) for v in out.cpu())))
    bad = {k: v for k, v in vals.items() if not (v < FP16_OPERAND_LIMIT)}
    if bad:
        raise MoldiffB200Error(f"activations outside the fp16 operand range of the tensor-core path: {bad} "
                               "(set MDB_DISABLE_TC=1 for the fp32 FFMA kernels)")
    return vals


def net_forward(net: PackedNet, plan: GraphPlan, h_node, pos, h_edge, node_time, edge_time):
    lib = load_library()
    h_node, pos, h_edge = _dev_f32(h_node, "h_node"), _dev_f32(pos, "pos_node"), _dev_f32(h_edge, "h_edge")
    node_time = _dev_f32(node_time, "node_time").reshape(-1)
    edge_time = _dev_f32(edge_time, "edge_time").reshape(-1)
    N, E = plan.n_nodes, plan.n_edges
    if h_node.shape != (N, packing.NODE_DIM) or h_edge.shape != (E, packing.EDGE_DIM) or pos.shape != (N, 3):
        raise MoldiffB200Error(f"shape mismatch: h_node {tuple(h_node.shape)}, h_edge {tuple(h_edge.shape)}, "
                               f"pos {tuple(pos.shape)} for N={N}, E={E} (kernels are built for 256/64)")
    out_node, out_pos, out_edge = torch.empty_like(h_node), torch.empty_like(pos), torch.empty_like(h_edge)
    ws = plan.workspace(0, net.num_blocks)
    rc = lib.mdb_net_forward(C.byref(net.desc), C.byref(plan.c), h_node.data_ptr(), pos.data_ptr(),
                             h_edge.data_ptr(), node_time.data_ptr(), edge_time.data_ptr(),
                             out_node.data_ptr(), out_pos.data_ptr(), out_edge.data_ptr(),
                             ws.data_ptr(), ws.numel() * 4, _stream_ptr(h_node.device))
    _check(rc, "mdb_net_forward")
    return out_node, out_pos, out_edge


def moldiff_forward(net: PackedNet, plan: GraphPlan, h_node_pert, pos_pert, h_edge_pert, batch_node, batch_edge, t):
    lib = load_library()
    h_node_pert, pos_pert = _dev_f32(h_node_pert, "h_node_pert"), _dev_f32(pos_pert, "pos_pert")
    h_edge_pert = _dev_f32(h_edge_pert, "h_edge_pert")
    batch_node, batch_edge, t = _dev_i64(batch_node, "batch_node"), _dev_i64(batch_edge, "batch_edge"), _dev_i64(t, "t")
    N, E = plan.n_nodes, plan.n_edges
    kn, ke = net.desc.num_node_types, net.desc.num_edge_types
    if h_node_pert.shape != (N, kn) or h_edge_pert.shape != (E, ke) or pos_pert.shape != (N, 3):
        raise MoldiffB200Error("MolDiff.forward: input shapes do not match the graph plan / type counts")
    dev = pos_pert.device
    pred_node = torch.empty(N, kn, dtype=torch.float32, device=dev)
    pred_pos = torch.empty(N, 3, dtype=torch.float32, device=dev)
    pred_half = torch.empty(E // 2, ke, dtype=torch.float32, device=dev)
    ws = plan.workspace(0, net.num_blocks)
    rc = lib.mdb_moldiff_forward(C.byref(net.desc), C.byref(plan.c), h_node_pert.data_ptr(), pos_pert.data_ptr(),
                                 h_edge_pert.data_ptr(), batch_node.data_ptr(), batch_edge.data_ptr(), t.data_ptr(),
                                 pred_node.data_ptr(), pred_pos.data_ptr(), pred_half.data_ptr(),
                                 ws.data_ptr(), ws.numel() * 4, _stream_ptr(dev))
    _check(rc, "mdb_moldiff_forward")
    return pred_node, pred_pos, pred_half


def transition_step(pos_tr, node_tr, edge_tr, t, batch_node, batch_half, pos, pred_pos, pred_node, log_node,
                    pred_half, log_half, noise=None):
    """One fused reverse-transition step of the sampler (`mdb_transition_step`): Gaussian posterior for the positions,
    categorical posteriors + Gumbel-max for node / half-edge types.  The random variates are drawn here from torch's
    generator in the order the unfused PyTorch path consumes them (positions, node types, half-edge types).
    `noise` = (z_pos ~ N(0,1) [N,3], u_node ~ U[0,1) [N,Kn], u_half ~ U[0,1) [Eh,Ke]) replaces the draws (parity tests).
    Returns (pos_prev, log_node, h_node_prev, log_half, h_edge_prev [2 Eh, Ke], half_type_prev [Eh])."""
    lib = load_library()
    pos, pred_pos = _dev_f32(pos, "pos"), _dev_f32(pred_pos, "pred_pos")
    pred_node, log_node = _dev_f32(pred_node, "pred_node"), _dev_f32(log_node, "log_node")
    pred_half, log_half = _dev_f32(pred_half, "pred_halfedge"), _dev_f32(log_half, "log_halfedge")
    batch_node, batch_half, t = _dev_i64(batch_node, "batch_node"), _dev_i64(batch_half, "batch_halfedge"), _dev_i64(t, "t")
    N, Eh = pos.shape[0], pred_half.shape[0]
    kn, ke = pred_node.shape[1], pred_half.shape[1]
    if kn != node_tr.num_classes or ke != edge_tr.num_classes or kn > 16 or ke > 16:
        raise MoldiffB200Error("transition_step: class counts do not match the transitions (or exceed 16)")
    dev = pos.device
    if noise is None:
        z_pos = torch.randn_like(pos)
        u_node = torch.rand_like(pred_node)
        u_half = torch.rand_like(pred_half)
    else:
        z_pos, u_node, u_half = (_dev_f32(x, "noise") for x in noise)
        if z_pos.shape != pos.shape or u_node.shape != pred_node.shape or u_half.shape != pred_half.shape:
            raise MoldiffB200Error("transition_step: noise shapes do not match the state")
    pos_out = torch.empty_like(pos)
    log_node_out, h_node_out = torch.empty_like(log_node), torch.empty_like(log_node)
    log_half_out = torch.empty_like(log_half)
    h_edge_out = torch.empty(2 * Eh, ke, dtype=torch.float32, device=dev)
    half_type = torch.empty(Eh, dtype=torch.int64, device=dev)
    tabs = [pos_tr.coef_x0, pos_tr.coef_xt, pos_tr.std, node_tr.q_mats, node_tr.transpopse_q_onestep_mats,
            edge_tr.q_mats, edge_tr.transpopse_q_onestep_mats]
    c0, ct, sd, qn, qnt, qe, qet = [_dev_f32(x.detach(), "transition table") for x in tabs]
    rc = lib.mdb_transition_step(
        N, Eh, kn, ke, batch_node.data_ptr(), batch_half.data_ptr(), t.data_ptr(), pos.data_ptr(), pred_pos.data_ptr(),
        z_pos.data_ptr(), c0.data_ptr(), ct.data_ptr(), sd.data_ptr(), pos_out.data_ptr(), pred_node.data_ptr(),
        log_node.data_ptr(), u_node.data_ptr(), qn.data_ptr(), qnt.data_ptr(), log_node_out.data_ptr(), h_node_out.data_ptr(),
        pred_half.data_ptr(), log_half.data_ptr(), u_half.data_ptr(), qe.data_ptr(), qet.data_ptr(), log_half_out.data_ptr(),
        h_edge_out.data_ptr(), half_type.data_ptr(), _stream_ptr(dev))
    _check(rc, "mdb_transition_step")
    return pos_out, log_node_out, h_node_out, log_half_out, h_edge_out, half_type


def bondpred_forward(net: PackedNet, plan: GraphPlan, h_node, pos, batch_node, batch_edge, t, save=False):
    lib = load_library()
    h_node, pos = _dev_f32(h_node, "h_node"), _dev_f32(pos, "pos_node")
    batch_node, batch_edge, t = _dev_i64(batch_node, "batch_node"), _dev_i64(batch_edge, "batch_edge"), _dev_i64(t, "t")
    N, E = plan.n_nodes, plan.n_edges
    if h_node.shape != (N, net.desc.num_node_types) or pos.shape != (N, 3):
        raise MoldiffB200Error("BondPredictor.forward: input shapes do not match the graph plan / type counts")
    logits = torch.empty(E // 2, net.desc.num_edge_types, dtype=torch.float32, device=pos.device)
    ws = plan.workspace(1 if save else 0, net.num_blocks)
    if save:
        plan.save_generation += 1
    rc = lib.mdb_bondpred_forward(C.byref(net.desc), C.byref(plan.c), h_node.data_ptr(), pos.data_ptr(),
                                  batch_node.data_ptr(), batch_edge.data_ptr(), t.data_ptr(), logits.data_ptr(),
                                  1 if save else 0, ws.data_ptr(), ws.numel() * 4, _stream_ptr(pos.device))
    _check(rc, "mdb_bondpred_forward")
    return logits


def bondpred_backward(net: PackedNet, plan: GraphPlan, h_node, pos, batch_node, batch_edge, t, d_logits, generation=None):
    """d(sum(logits * d_logits)) / d pos through the bond predictor (hand-written backward kernels).  Must
    follow bondpred_forward(save=True) with the same inputs: the saved activations live in the plan's workspace.  `generation`
    = plan.save_generation right after that forward; a later save-forward on the same plan has overwritten them and the call
    raises instead of returning the gradient of the wrong inputs."""
    lib = load_library()
    if generation is not None and generation != plan.save_generation:
        raise MoldiffB200Error("bondpred_backward: the activations saved by this forward were overwritten by a later "
                               "BondPredictor forward on the same graph (one backward per forward, in order)")
    h_node, pos = _dev_f32(h_node, "h_node"), _dev_f32(pos, "pos_node")
    d_logits = _dev_f32(d_logits, "d_logits")
    batch_node, batch_edge, t = _dev_i64(batch_node, "batch_node"), _dev_i64(batch_edge, "batch_edge"), _dev_i64(t, "t")
    if d_logits.shape != (plan.n_edges // 2, net.desc.num_edge_types):
        raise MoldiffB200Error("bondpred_backward: d_logits shape mismatch")
    d_pos = torch.empty(plan.n_nodes, 3, dtype=torch.float32, device=pos.device)
    ws = plan.workspace(1, net.num_blocks)
    rc = lib.mdb_bondpred_backward(C.byref(net.desc), C.byref(plan.c), h_node.data_ptr(), pos.data_ptr(),
                                   batch_node.data_ptr(), batch_edge.data_ptr(), t.data_ptr(),
                                   d_logits.data_ptr(), d_pos.data_ptr(),
                                   ws.data_ptr(), ws.numel() * 4, _stream_ptr(pos.device))
    _check(rc, "mdb_bondpred_backward")
    return d_pos


def tc_selftest(x, w_kn, twice=False, cross_first=False):
    """y = x @ w (x [128][K], w [K][N]) through the tcgen05 split-fp16 pipeline (tests only)."""
    lib = load_library()
    x = _dev_f32(x, "x")
    k, n = w_kn.shape
    img = packing.tc_image(w_kn.cpu()).to(x.device)
    y = torch.empty(128, n, dtype=torch.float32, device=x.device)
    rc = lib.mdb_tc_selftest(x.data_ptr(), img.data_ptr(), y.data_ptr(), k, n, (1 if twice else 0) | (2 if cross_first else 0),
                             _stream_ptr(x.device))
    _check(rc, "mdb_tc_selftest")
    return y
