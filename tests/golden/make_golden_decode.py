"""Golden vectors for the output decode, produced by the UNMODIFIED reference functions.

`utils/transforms.py` / `utils/sample.py` cannot be imported here (rdkit, lmdb, torch_geometric at import time), so the two
function bodies are cut out of the reference sources with `ast` and executed as they are: `seperate_outputs`
(utils/sample.py) and `FeaturizeMol.decode_output` (utils/transforms.py), with scipy's softmax as in the reference.
Run in the build container:  python tests/golden/make_golden_decode.py  ->  tests/golden/golden_decode.npz
"""
import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def reference_functions():
    from scipy.special import softmax
    ns = {"np": np, "softmax": softmax}
    src = open(os.path.join(REF, "utils/sample.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "seperate_outputs")
    exec(compile(ast.Module([fn], []), "utils/sample.py", "exec"), ns)
    src = open(os.path.join(REF, "utils/transforms.py")).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "FeaturizeMol")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "decode_output")
    exec(compile(ast.Module([fn], []), "utils/transforms.py", "exec"), ns)
    return ns["seperate_outputs"], ns["decode_output"]


def inputs(B=6, seed=11, mask_bias=1.5):
    from moldiff_b200.placeholder import make_data_placeholder
    np.random.seed(2023)
    ph = make_data_placeholder(B)
    N, Eh = len(ph["batch_node"]), len(ph["batch_halfedge"])
    g = torch.Generator().manual_seed(seed)
    pred_node = 2 * torch.randn(N, 8, generator=g)
    pred_node[:, 7] += mask_bias
    pred_half = 2 * torch.randn(Eh, 6, generator=g)
    pred_half[:, 0] += 2.0
    return ph, pred_node.numpy(), torch.randn(N, 3, generator=g).numpy(), pred_half.numpy()


def main():
    seperate_outputs, decode_output = reference_functions()
    atoms = [6, 7, 8, 9, 15, 16, 17]
    feat = types.SimpleNamespace(num_element=7, num_bond_types=4, num_edge_types=6,
                                 nodetype_to_ele={i: e for i, e in enumerate(atoms)})
    B = 6
    ph, pn, pp, phf = inputs(B)
    bn, hei, bh = ph["batch_node"].numpy(), ph["halfedge_index"].numpy(), ph["batch_halfedge"].numpy()
    outputs = {"pred": [pn, pp, phf], "traj": [pn[None], pp[None], phf[None]]}
    out = {"B": B}
    for m, o in enumerate(seperate_outputs(outputs, B, bn, hei, bh)):
        info = decode_output(feat, pred_node=o["pred"][0], pred_pos=o["pred"][1], pred_halfedge=o["pred"][2],
                             halfedge_index=o["halfedge_index"])
        for k, v in info.items():
            out[f"m{m}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/golden_decode.npz"), **out)
    print("wrote golden_decode.npz:", len(out) - 1, "arrays")


if __name__ == "__main__":
    main()
