"""TEST INFRASTRUCTURE ONLY (never imported by the product): numpy restatement of the reference's output decode.

  seperate_outputs   utils/sample.py:4-30       per-molecule boolean masks over the batch, half-edge index re-based
  decode_output      utils/transforms.py:65-122 softmax / argmax / max, masked-atom removal, bond filtering and doubling

Parity pin: `tests/test_oracle_decode.py` checks this file bit-exactly against golden vectors produced by the unmodified
reference function bodies (`tests/golden/make_golden_decode.py` cuts them out of the reference sources with `ast`,
because utils/transforms.py cannot be imported without rdkit / lmdb / torch_geometric).
"""
import numpy as np


def softmax(x, axis=-1):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def seperate_pred(pred, n_graphs, batch_node, halfedge_index, batch_halfedge):
    out = []
    for i in range(n_graphs):
        ind_node, ind_half = batch_node == i, batch_halfedge == i
        assert ind_node.sum() * (ind_node.sum() - 1) == ind_half.sum() * 2
        he = halfedge_index[:, ind_half]
        out.append({"pred": [pred[0][ind_node], pred[1][ind_node], pred[2][ind_half]],
                    "halfedge_index": he - ind_node.nonzero()[0].min()})
    return out


def decode_output(pred_node, pred_pos, pred_halfedge, halfedge_index, atomic_numbers, num_bond_types):
    num_element = len(atomic_numbers)
    pred_atom = softmax(pred_node, axis=-1)
    atom_type, atom_prob = np.argmax(pred_atom, axis=-1), np.max(pred_atom, axis=-1)
    keep = atom_type < num_element
    if not keep.all():
        changer = -np.ones(len(keep), dtype=np.int64)
        changer[keep] = np.arange(keep.sum())
    atom_type, atom_prob = atom_type[keep], atom_prob[keep]
    element = np.array([atomic_numbers[i] for i in atom_type])
    atom_pos = pred_pos[keep]
    ph = softmax(pred_halfedge, axis=-1)
    edge_type, edge_prob = np.argmax(ph, axis=-1), np.max(ph, axis=-1)
    is_bond = (edge_type > 0) & (edge_type <= num_bond_types)
    bond_type, bond_prob, bond_index = edge_type[is_bond], edge_prob[is_bond], halfedge_index[:, is_bond]
    if not keep.all():
        bond_index = changer[bond_index]
        bad = (bond_index < 0).any(axis=0)
        bond_index, bond_type, bond_prob = bond_index[:, ~bad], bond_type[~bad], bond_prob[~bad]
    return {"element": element, "atom_pos": atom_pos, "atom_prob": atom_prob,
            "bond_type": np.concatenate([bond_type, bond_type]), "bond_prob": np.concatenate([bond_prob, bond_prob]),
            "bond_index": np.concatenate([bond_index, bond_index[::-1]], axis=1)}
