"""GPU tool: worst-case parity of MolDiff.forward against the CPU oracle at BASELINE config-2 size (B=256)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moldiff_b200 import MolDiff  # noqa: E402
from moldiff_b200.config import builtin_config  # noqa: E402
from oracle import restatement as R  # noqa: E402
from tests.helpers import batch_inputs, doubled, oracle_moldiff, to_dev  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).eval()
inp = batch_inputs(B=B, seed_graph=2023, seed_inputs=7, t_values=(900, 500, 100))
t0 = time.time()
ref = oracle_moldiff(model.state_dict(), inp)
print(f"oracle B={B}: {time.time() - t0:.1f} s")
gm = model.to(dev)
d = to_dev(inp, dev)
ei, be, he = doubled(d)
with torch.no_grad():
    out = gm(d["h_node"], d["pos"], d["batch_node"], he, ei, be, d["t"])
for k in ref:
    a, b = out[k].cpu().double(), ref[k].double()
    err = (a - b).abs()
    print(f"  {k:14s} max|d|/max|ref| = {float(err.max() / b.abs().max()):.2e}   rms(d)/rms(ref) = "
          f"{float(err.pow(2).mean().sqrt() / b.pow(2).mean().sqrt()):.2e}   max|ref| = {float(b.abs().max()):.3f}")
