"""GPU tool: where do the worst MolDiff.forward errors at config-2 size sit?  Compares the tensor-core path, the fp32
FFMA path and the fp32 CPU oracle against the fp64 oracle, per molecule."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moldiff_b200 import MolDiff  # noqa: E402
from moldiff_b200.config import builtin_config  # noqa: E402
from tests.helpers import batch_inputs, doubled, oracle_moldiff, to_dev  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).eval()
inp = batch_inputs(B=B, seed_graph=2023, seed_inputs=7, t_values=(900, 500, 100))
sd = model.state_dict()
ref32 = oracle_moldiff(sd, inp)
inp64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in inp.items()}
ref64 = oracle_moldiff({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, inp64)


def run(disable_tc):
    os.environ["MDB_DISABLE_TC"] = "1" if disable_tc else "0"
    m = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).eval()
    m.load_state_dict(sd)
    m = m.to(dev)
    d = to_dev(inp, dev)
    ei, be, he = doubled(d)
    with torch.no_grad():
        return {k: v.cpu() for k, v in m(d["h_node"], d["pos"], d["batch_node"], he, ei, be, d["t"]).items()}


tc, ff = run(False), run(True)
bn = inp["batch_node"]
_, be, _ = doubled(inp)
bh = be[: be.numel() // 2]
for k, owner in (("pred_node", bn), ("pred_pos", bn), ("pred_halfedge", bh)):
    r = ref64[k]
    scale = float(r.abs().max())
    print(f"{k}: max|ref| {scale:.3f}")
    for name, o in (("tc", tc[k]), ("ffma", ff[k]), ("oracle32", ref32[k])):
        err = (o.double() - r).abs().reshape(len(owner), -1).max(dim=1).values / scale
        per_mol = torch.zeros(B, dtype=torch.float64).scatter_reduce_(0, owner, err, "amax")
        top = torch.topk(per_mol, 5)
        print(f"   {name:9s} max {float(err.max()):.2e}  median-mol {float(per_mol.median()):.2e}  top mols "
              + " ".join(f"{int(i)}:{float(v):.1e}" for v, i in zip(top.values, top.indices)))
# the worst molecule: its size, time step and closest atom pair
err = (tc["pred_pos"].double() - ref64["pred_pos"]).abs().max(dim=1).values
w = int(bn[err.argmax()])
sel = bn == w
p = inp["pos"][sel].double()
dist = torch.cdist(p, p) + torch.eye(len(p)) * 1e9
print(f"worst molecule {w}: {int(sel.sum())} atoms, t={int(inp['t'][w])}, min pair distance {float(dist.min()):.4f}, "
      f"|pos|max {float(p.abs().max()):.2f}")
