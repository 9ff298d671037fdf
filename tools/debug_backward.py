"""GPU debugging aid: compares the CUDA backward's d_pos against the CPU emulator / autograd per molecule."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moldiff_b200 import BondPredictor  # noqa: E402
from moldiff_b200.config import builtin_config  # noqa: E402
from oracle import restatement as R  # noqa: E402
from tests.helpers import batch_inputs, doubled, per_molecule_rel_err, to_dev  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    bp = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).eval()
    sd = bp.state_dict()
    inp = batch_inputs(B=6, t_values=(999, 400, 0))
    ei, be, _ = doubled(inp)
    g = torch.Generator().manual_seed(3)
    w = torch.randn(ei.shape[1] // 2, 5, generator=g)
    pos = inp["pos"].clone().requires_grad_(True)
    lg = R.bondpred_forward(sd, inp["h_node"], pos, inp["batch_node"], ei, be, inp["t"])
    ref = torch.autograd.grad((lg * w).sum(), pos)[0]
    gbp = bp.to(dev)
    d = to_dev(inp, dev)
    eid, bed, _ = doubled(d)
    pos_in = d["pos"].clone().requires_grad_(True)
    logits = gbp(d["h_node"], pos_in, d["batch_node"], eid, bed, d["t"])
    print("logits rel err", R.rel_err(logits.detach().cpu(), lg.detach()))
    grad = torch.autograd.grad((logits * w.to(dev)).sum(), pos_in)[0].cpu()
    print("d_pos overall rel err", R.rel_err(grad, ref))
    print("per molecule", per_molecule_rel_err(grad, ref, inp["batch_node"]))
    print("sample", grad[:3], ref[:3])


if __name__ == "__main__":
    main()
