"""Generates tests/golden/golden_loss.pt: MolDiff.get_loss of the UNMODIFIED reference (models/model.py:128-201,
imported through oracle/ref_shim.py) on seeded synthetic molecules, with everything the reference drew at random
recorded -- the time steps (sample_time), the perturbed molecule (pos_pert, one-hot node / half-edge types and their
log-probability encodings) -- so that a test can teacher-force the same perturbation through another implementation
and compare the four loss values.  Also asserts oracle/restatement.loss_terms == reference on what it writes.

    python tests/golden/make_golden_loss.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import restatement as R  # noqa: E402
from oracle.ref_shim import load_reference, load_yaml_config  # noqa: E402

CASES = {"B8_seed5": dict(B=8, seed=5), "B24_seed6": dict(B=24, seed=6)}


def clean_molecules(B, seed):
    """Synthetic 'dataset' batch: graph sizes as make_data_placeholder (utils/transforms.py:125-156), random atom / bond
    types in the real (non-mask) classes, N(0, 1) * 2 positions."""
    np.random.seed(2023 + seed)
    ph = R.make_data_placeholder(B)
    g = torch.Generator().manual_seed(seed)
    n, eh = len(ph["batch_node"]), len(ph["batch_halfedge"])
    return dict(batch_node=ph["batch_node"], halfedge_index=ph["halfedge_index"], batch_halfedge=ph["batch_halfedge"],
                node_type=torch.randint(0, 7, (n,), generator=g), halfedge_type=torch.randint(0, 5, (eh,), generator=g),
                node_pos=torch.randn(n, 3, generator=g) * 2.0)


def main():
    torch.set_num_threads(8)
    ref = load_reference()
    cfg = load_yaml_config("configs/train/train_MolDiff.yml")
    torch.manual_seed(0)
    model = ref.model.MolDiff(cfg.model, 8, 6).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    out = {}
    for name, c in CASES.items():
        mol = clean_molecules(c["B"], c["seed"])
        rec = {}
        orig_forward = model.forward
        orig_time = model.sample_time
        orig_noise = {k: getattr(model, k).add_noise for k in ("pos_transition", "node_transition", "edge_transition")}

        def forward(*a, **k):
            rec["fwd_in"] = [x.detach().clone() for x in a]
            r = orig_forward(*a, **k)
            rec["preds"] = {kk: v.detach().clone() for kk, v in r.items()}
            return r

        def sample_time(*a, **k):
            r = orig_time(*a, **k)
            rec["time_step"] = r[0].clone()
            return r

        def wrap(key):
            def f(*a, **k):
                r = orig_noise[key](*a, **k)
                rec[key] = [x.detach().clone() for x in r] if isinstance(r, (tuple, list)) else r.detach().clone()
                return r
            return f

        model.forward = forward
        model.sample_time = sample_time
        for k in orig_noise:
            getattr(model, k).add_noise = wrap(k)
        torch.manual_seed(100 + c["seed"])
        with torch.no_grad():
            losses = model.get_loss(mol["node_type"], mol["node_pos"], mol["batch_node"], mol["halfedge_type"],
                                    mol["halfedge_index"], mol["batch_halfedge"], c["B"])
        model.forward = orig_forward
        model.sample_time = orig_time
        for k, f in orig_noise.items():
            getattr(model, k).add_noise = f
        h_node_pert, log_node_t, log_node_0 = rec["node_transition"]
        h_half_pert, log_half_t, log_half_0 = rec["edge_transition"]
        pr = rec["preds"]
        mine = R.loss_terms(sd, mol["node_pos"], rec["time_step"], mol["batch_node"], mol["batch_halfedge"],
                            pr["pred_node"], pr["pred_pos"], pr["pred_halfedge"], log_node_t, log_node_0, log_half_t, log_half_0)
        for k in ("loss", "loss_pos", "loss_node", "loss_edge"):
            e = abs(float(mine[k]) - float(losses[k])) / abs(float(losses[k]))
            print(f"  {name}: oracle vs reference {k:10s} {float(losses[k]):.6f} rel err {e:.1e}")
            assert e < 1e-6, (k, e)
        out[name] = dict(args=c, torch_seed=100 + c["seed"], time_step=rec["time_step"], pos_pert=rec["pos_transition"],
                         h_node_pert=h_node_pert, log_node_t=log_node_t, log_node_0=log_node_0,
                         h_half_pert=h_half_pert, log_half_t=log_half_t, log_half_0=log_half_0,
                         preds={k: v.clone() for k, v in pr.items()},
                         losses={k: float(v) for k, v in losses.items()})
    torch.save(out, os.path.join(ROOT, "tests", "golden", "golden_loss.pt"))


if __name__ == "__main__":
    main()
