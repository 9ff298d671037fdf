"""Decode + per-molecule separation of a sampled batch.

Reference: `utils/sample.py:4-30` (`seperate_outputs`: one boolean mask over the whole batch per molecule) followed by
`FeaturizeMol.decode_output` (`utils/transforms.py:65-122`) per molecule, after a synchronous D2H of the logits and of
the full [T+1, ...] trajectory (`scripts/sample_drug3d.py:125-140`).  Here the per-row arithmetic (softmax, argmax,
max) runs on the GPU for the whole batch (`mdb_decode_rows`), 5 bytes per row + the positions cross PCIe from pinned
buffers, and the split uses the contiguous per-molecule offsets instead of n_graphs full-array masks.  Returns the same
dictionaries `decode_output` returns (numpy arrays), one per molecule.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import engine


def _offsets(batch, n_items, n_graphs):
    # contiguous per-molecule segments are what makes the offset split valid (the reference's boolean masks accept any order)
    if batch.numel() > 1 and bool((batch[1:] < batch[:-1]).any()):
        raise engine.MoldiffB200Error("decode_batch needs batch vectors sorted by molecule (as make_data_placeholder / "
                                      "the PyG loader produce them)")
    counts = torch.bincount(batch, minlength=n_graphs)
    off = np.zeros(n_graphs + 1, dtype=np.int64)
    off[1:] = np.cumsum(counts.cpu().numpy())
    if off[-1] != n_items:
        raise engine.MoldiffB200Error("batch vector does not cover every row")
    return off


def decode_batch(pred_node, pred_pos, pred_halfedge, n_graphs, batch_node, halfedge_index, batch_halfedge,
                 atomic_numbers=(6, 7, 8, 9, 15, 16, 17), num_bond_types=4):
    """[decode_output(...) for every molecule of the batch].  `atomic_numbers` / `num_bond_types` are FeaturizeMol's
    (configs: chem.atomic_numbers, len(chem.mol_bond_types)); node classes >= len(atomic_numbers) are mask atoms and edge
    classes outside 1..num_bond_types are "no bond" / mask (transforms.py:80,98)."""
    if not pred_node.is_cuda:
        raise engine.MoldiffB200Error("decode_batch takes CUDA tensors: moldiff_b200 has no CPU path")
    lib = engine.load_library()
    dev = pred_node.device
    pred_node, pred_halfedge = pred_node.float().contiguous(), pred_halfedge.float().contiguous()
    N, Eh = pred_node.shape[0], pred_halfedge.shape[0]
    node_type = torch.empty(N, dtype=torch.uint8, device=dev)
    half_type = torch.empty(Eh, dtype=torch.uint8, device=dev)
    node_prob = torch.empty(N, dtype=torch.float32, device=dev)
    half_prob = torch.empty(Eh, dtype=torch.float32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.mdb_decode_rows(N, pred_node.shape[1], p(pred_node), Eh, pred_halfedge.shape[1], p(pred_halfedge), p(node_type),
                             p(node_prob), p(half_type), p(half_prob), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    engine._check(rc, "mdb_decode_rows")

    def to_host(t):                                   # pinned staging, all copies in flight before one sync
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        return h
    host = [to_host(t) for t in (node_type, node_prob, pred_pos.float().contiguous(), half_type, half_prob,
                                 halfedge_index.contiguous())]
    off_n = _offsets(batch_node, N, n_graphs)
    off_e = _offsets(batch_halfedge, Eh, n_graphs)
    torch.cuda.current_stream(dev).synchronize()
    atom_type, atom_prob, pos, edge_type, edge_prob, he_index = [h.numpy() for h in host]
    ele = np.asarray(atomic_numbers)
    num_element = len(ele)
    out = []
    for m in range(n_graphs):
        n0, n1, e0, e1 = off_n[m], off_n[m + 1], off_e[m], off_e[m + 1]
        n = n1 - n0
        if n * (n - 1) != 2 * (e1 - e0):
            raise engine.MoldiffB200Error("molecule is not a complete graph of half-edges (utils/sample.py:12)")
        at, keep = atom_type[n0:n1].astype(np.int64), None
        keep = at < num_element
        info = {"element": ele[at[keep]], "atom_pos": pos[n0:n1][keep], "atom_prob": atom_prob[n0:n1][keep]}
        et = edge_type[e0:e1].astype(np.int64)
        is_bond = (et > 0) & (et <= num_bond_types)
        bond_type, bond_prob = et[is_bond], edge_prob[e0:e1][is_bond]
        bond_index = he_index[:, e0:e1][:, is_bond] - n0
        if not keep.all():
            changer = -np.ones(n, dtype=np.int64)
            changer[keep] = np.arange(keep.sum())
            bond_index = changer[bond_index]
            bad = (bond_index < 0).any(axis=0)
            bond_index, bond_type, bond_prob = bond_index[:, ~bad], bond_type[~bad], bond_prob[~bad]
        info["bond_type"] = np.concatenate([bond_type, bond_type])
        info["bond_prob"] = np.concatenate([bond_prob, bond_prob])
        info["bond_index"] = np.concatenate([bond_index, bond_index[::-1]], axis=1)
        out.append(info)
    return out
