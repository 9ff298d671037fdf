"""Edge builders over 3D atom positions: `radius_graph` / `knn_graph` with torch_geometric's signatures
(`torch_geometric.nn.radius_graph`, `knn_graph`; imported by the reference at models/graph.py:6), on CUDA through
`mdb_radius_graph` / `mdb_knn_graph` (csrc/mdb_graph_build.cuh).  The reference reaches them only from dead code and
torch_cluster is not installable here, so their results are unpinned by the reference: semantics
follow torch_cluster's documented behaviour and are tested against a brute-force PyTorch restatement.

Returned `edge_index` is [2, E] int64 with (source, target) = (neighbour j, centre i) for flow='source_to_target' and
the two rows swapped for 'target_to_source'; edges are grouped by centre in increasing order.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import engine


def _segments(n, batch, device):
    """First / one-past-last node index of every node's graph (graphs must be contiguous, as PyG requires)."""
    if batch is None:
        lo = torch.zeros(n, dtype=torch.int32, device=device)
        return lo, torch.full((n,), n, dtype=torch.int32, device=device)
    batch = batch.to(device=device, dtype=torch.int64)
    if n > 1 and bool((batch[1:] < batch[:-1]).any()):
        raise engine.MoldiffB200Error("graph builders need `batch` sorted (nodes of one graph contiguous)")
    ar = torch.arange(n, device=device)
    first = torch.ones(n, dtype=torch.bool, device=device)
    first[1:] = batch[1:] != batch[:-1]
    starts = ar[first]                                             # start index of each graph
    gid = torch.cumsum(first.to(torch.int64), 0) - 1               # dense graph id per node
    ends = torch.cat([starts[1:], torch.tensor([n], device=device)])
    return starts[gid].to(torch.int32), ends[gid].to(torch.int32)


def _build(kind, pos, param, batch, loop, width, flow):
    if flow not in ("source_to_target", "target_to_source"):
        raise ValueError(flow)
    if not pos.is_cuda:
        raise engine.MoldiffB200Error("pos must be a CUDA tensor: moldiff_b200 has no CPU path")
    if pos.ndim != 2 or pos.shape[1] != 3:
        raise engine.MoldiffB200Error("graph builders take [N, 3] positions")
    lib = engine.load_library()
    pos = pos.detach().float().contiguous()
    n, dev = pos.shape[0], pos.device
    if n == 0:
        return torch.zeros(2, 0, dtype=torch.int64, device=dev)
    lo, hi = _segments(n, batch, dev)
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    nbr = torch.empty(n, width, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    if kind == "radius":
        rc = lib.mdb_radius_graph(n, p(pos), p(lo), p(hi), C.c_float(float(param)), int(bool(loop)), int(width), p(counts),
                                  p(nbr), st)
    else:
        rc = lib.mdb_knn_graph(n, p(pos), p(lo), p(hi), int(param), int(bool(loop)), p(counts), p(nbr), st)
    engine._check(rc, f"mdb_{kind}_graph")
    keep = torch.arange(width, device=dev)[None, :] < counts[:, None]
    centre = torch.arange(n, device=dev)[:, None].expand(n, width)[keep]
    neigh = nbr[keep].to(torch.int64)
    src, dst = (neigh, centre) if flow == "source_to_target" else (centre, neigh)
    return torch.stack([src, dst], dim=0)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target", **_ignored):
    """All pairs (j -> i) of one graph with |x_i - x_j| < r; at most `max_num_neighbors` per centre i (in index order)."""
    return _build("radius", x, r, batch, loop, int(max_num_neighbors), flow)


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", **_ignored):
    """The k nearest nodes j of every centre i inside its graph (fewer in graphs with < k + 1 nodes), k <= 32."""
    return _build("knn", x, k, batch, loop, int(k), flow)
