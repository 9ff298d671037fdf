"""Synthetic batch skeletons for sampling: per-molecule atom counts ~ N(24.92, 5.52) (GEOM-Drugs
statistics) and the complete upper-triangle half-edge list of every molecule.

Host-side mirror of the reference's ``utils/transforms.py:125-156`` (`make_data_placeholder`): same
argument meaning, same use of the *global* numpy RNG (so `np.random.seed(s)` reproduces the reference's
graphs exactly), same output dict of int64 tensors.
"""
from __future__ import annotations

import numpy as np
import torch

GEOM_DRUGS_MEAN_ATOMS = 24.923464980477522
GEOM_DRUGS_STD_ATOMS = 5.516291901819105


def draw_sizes(n_graphs, max_size=None):
    """The atom counts `make_data_placeholder` would draw (same RNG consumption)."""
    if max_size is None:
        sizes = np.random.normal(GEOM_DRUGS_MEAN_ATOMS, GEOM_DRUGS_STD_ATOMS, size=n_graphs)
    else:
        sizes = np.array([max_size] * n_graphs)
    return sizes.astype("int64")


def make_data_placeholder(n_graphs, device=None, max_size=None, sizes=None):
    """`sizes` (extension): explicit atom counts, e.g. one rank's share of a batch (sharding.balanced_shards)."""
    if sizes is not None:
        sizes = np.asarray(sizes)
        n_graphs = len(sizes)
    elif max_size is None:
        sizes = np.random.normal(GEOM_DRUGS_MEAN_ATOMS, GEOM_DRUGS_STD_ATOMS, size=n_graphs)
    else:
        sizes = np.array([max_size] * n_graphs)
    sizes = sizes.astype("int64")
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    mol_of_node = np.repeat(np.arange(n_graphs), sizes)
    rows, cols, mol_of_pair = [], [], []
    for m, (n, o) in enumerate(zip(sizes, offsets)):
        i, j = np.triu_indices(int(n), 1)
        rows.append(i + o)
        cols.append(j + o)
        mol_of_pair.append(np.full(i.shape[0], m))
    out = {
        "batch_node": torch.from_numpy(mol_of_node).long(),
        "halfedge_index": torch.from_numpy(np.stack([np.concatenate(rows), np.concatenate(cols)])).long(),
        "batch_halfedge": torch.from_numpy(np.concatenate(mol_of_pair)).long(),
    }
    if device is not None:
        out = {k: v.to(device) for k, v in out.items()}
    return out
