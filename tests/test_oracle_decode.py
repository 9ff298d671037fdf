"""CPU: oracle/decode_restatement.py against golden vectors produced by the unmodified reference functions
(tests/golden/make_golden_decode.py: seperate_outputs + FeaturizeMol.decode_output cut out of the reference sources)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ATOMS = (6, 7, 8, 9, 15, 16, 17)


def test_decode_restatement_reproduces_reference_goldens():
    from oracle import decode_restatement as D
    from tests.golden.make_golden_decode import inputs
    gold = np.load(os.path.join(HERE, "golden", "golden_decode.npz"))
    B = int(gold["B"])
    ph, pn, pp, phf = inputs(B)
    sep = D.seperate_pred([pn, pp, phf], B, ph["batch_node"].numpy(), ph["halfedge_index"].numpy(), ph["batch_halfedge"].numpy())
    n_masked = 0
    for m in range(B):
        got = D.decode_output(*sep[m]["pred"], sep[m]["halfedge_index"], ATOMS, 4)
        n_masked += len(sep[m]["pred"][0]) - len(got["element"])
        for k in ("element", "atom_pos", "atom_prob", "bond_type", "bond_prob", "bond_index"):
            assert np.array_equal(got[k], gold[f"m{m}_{k}"]), (m, k)
    assert n_masked > 0          # the fixture exercises the masked-atom / dangling-bond branch
