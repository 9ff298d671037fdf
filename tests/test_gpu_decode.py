"""GPU parity of decode_batch (mdb_decode_rows + host bookkeeping) against the numpy restatement of the
reference's seperate_outputs + FeaturizeMol.decode_output (oracle/decode_restatement.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ATOMS = (6, 7, 8, 9, 15, 16, 17)


def _batch(B, seed, mask_bias=0.0):
    from moldiff_b200.placeholder import make_data_placeholder
    np.random.seed(2023 + seed)
    ph = make_data_placeholder(B)
    N, Eh = len(ph["batch_node"]), len(ph["batch_halfedge"])
    g = torch.Generator().manual_seed(seed)
    pred_node = 2 * torch.randn(N, 8, generator=g)
    pred_node[:, 7] += mask_bias                    # class 7 = mask atom: mask_bias > 0 makes masked atoms common
    pred_half = 2 * torch.randn(Eh, 6, generator=g)
    pred_half[:, 0] += 2.0                          # mostly "no bond"
    return ph, pred_node, torch.randn(N, 3, generator=g), pred_half


@pytest.mark.parametrize("mask_bias", [-50.0, 0.0, 3.0])
def test_decode_batch_matches_reference_restatement(mask_bias):
    from moldiff_b200.decode import decode_batch
    from oracle import decode_restatement as D
    B = 24
    ph, pred_node, pred_pos, pred_half = _batch(B, seed=7, mask_bias=mask_bias)
    dev = torch.device("cuda:0")
    got = decode_batch(pred_node.to(dev), pred_pos.to(dev), pred_half.to(dev), B, ph["batch_node"].to(dev),
                       ph["halfedge_index"].to(dev), ph["batch_halfedge"].to(dev), atomic_numbers=ATOMS, num_bond_types=4)
    sep = D.seperate_pred([pred_node.numpy(), pred_pos.numpy(), pred_half.numpy()], B, ph["batch_node"].numpy(),
                          ph["halfedge_index"].numpy(), ph["batch_halfedge"].numpy())
    assert len(got) == B
    for m in range(B):
        ref = D.decode_output(*sep[m]["pred"], sep[m]["halfedge_index"], ATOMS, 4)
        for k in ("element", "bond_type", "bond_index"):
            assert np.array_equal(got[m][k], ref[k]), (m, k)
        assert np.array_equal(got[m]["atom_pos"], ref["atom_pos"])
        assert np.allclose(got[m]["atom_prob"], ref["atom_prob"], rtol=2e-6, atol=1e-7)
        assert np.allclose(got[m]["bond_prob"], ref["bond_prob"], rtol=2e-6, atol=1e-7)


def test_decode_rejects_cpu_and_bad_batches():
    from moldiff_b200 import engine
    from moldiff_b200.decode import decode_batch
    ph, pred_node, pred_pos, pred_half = _batch(3, seed=1)
    with pytest.raises(engine.MoldiffB200Error):
        decode_batch(pred_node, pred_pos, pred_half, 3, ph["batch_node"], ph["halfedge_index"], ph["batch_halfedge"])
    dev = torch.device("cuda:0")
    perm = torch.randperm(len(ph["batch_node"]), generator=torch.Generator().manual_seed(0))
    with pytest.raises(engine.MoldiffB200Error):      # rows not grouped by molecule: the offset split would be wrong
        decode_batch(pred_node.to(dev), pred_pos.to(dev), pred_half.to(dev), 3, ph["batch_node"][perm].to(dev),
                     ph["halfedge_index"].to(dev), ph["batch_halfedge"].to(dev))
