"""Generates tests/golden/bondpred_ref64.pt: float64 guidance gradients d objective / d pos of the oracle
(oracle/restatement.py, itself pinned bitwise to the unmodified reference by make_golden.py) -- the "truth" that the
crossed tensor-core / fp32 gradient tests (tests/test_gpu_parity.py, tools/tc_numerics.py) measure against.
Raw gradients (not multiplied by -gui_scale), float64, [N, 3] per case and objective.

    python tests/golden/make_golden_ref64.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import restatement as R  # noqa: E402
from tests.helpers import batch_inputs, doubled  # noqa: E402

CASES = {"B16": dict(B=16, t_values=(400,)),                       # = golden.pt bondpred/B16
         "B48": dict(B=48, seed_graph=77, seed_inputs=78, t_values=(990, 700, 400, 150, 20), pos_scale=2.0)}


def main():
    torch.set_num_threads(8)
    from moldiff_b200 import BondPredictor
    from moldiff_b200.config import builtin_config
    torch.manual_seed(0)
    bp = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).eval()
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in bp.state_dict().items()}
    out = {}
    for name, args in CASES.items():
        inp = batch_inputs(**args)
        ei, be, _ = doubled(inp)
        rec = {"args": args}
        for gui in ("uncertainty", "entropy"):
            delta, logits = R.guidance_delta(sd64, inp["h_node"].double(), inp["pos"].double(), inp["batch_node"], ei, be,
                                             inp["t"], gui_type=gui, gui_scale=1.0)
            rec[gui] = (-delta).contiguous()
            rec["logits"] = logits.float()
            print(name, gui, "max |grad|", float(delta.abs().max()))
        out[name] = rec
    torch.save(out, os.path.join(ROOT, "tests", "golden", "bondpred_ref64.pt"))


if __name__ == "__main__":
    main()
