"""GPU tool: per-phase cycle breakdown of tc_nodeblock_bwd_kernel (clock64 stamps of row thread 0; config 2 graph, guided
step).  Stamps: 0 start, 1 set-up done, then alternating "accumulator arrived" / "A planes published"."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moldiff_b200 import BondPredictor, MolDiff, engine  # noqa: E402
from moldiff_b200.config import builtin_config  # noqa: E402
from moldiff_b200.placeholder import make_data_placeholder  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).to(dev).eval()
bond = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).to(dev).eval()
np.random.seed(2023)
ph = make_data_placeholder(256, device=dev)
st = model.sample_begin(256, ph["batch_node"], ph["halfedge_index"], ph["batch_halfedge"])
g = ("uncertainty", 1e-4)
for _ in range(2):
    model.sample_step(st, 500, bond_predictor=bond, guidance=g)
E = st["edge_index"].shape[1]
tiles = (E + 127) // 128
buf = torch.zeros(tiles * 32, dtype=torch.int64, device=dev)
lib = engine.load_library()
lib.mdb_debug_set_buffer.argtypes = [C.c_void_p]
lib.mdb_debug_select(1)
lib.mdb_debug_set_buffer(buf.data_ptr())
model.sample_step(st, 499, bond_predictor=bond, guidance=g)
torch.cuda.synchronize()
lib.mdb_debug_set_buffer(None)
t = buf.view(tiles, 32).cpu().numpy().astype(np.int64)
names = ["set-up", "e tile -> planes",
         "wait EN1", "epi LN(en1)", "wait EN2", "epi he*hn (+he scratch)", "wait MSG+GE", "epi LN(g1)+gx", "wait G2",
         "epi dout: dgt->planes, dmsg->scratch", "wait BT_G2+GE", "epi LN bwd(g1) + RED dgx", "wait BT_GE",
         "de -> regs, wait reload+BT_MSG", "epi dhn RED + d he", "wait BT_EN2+EN1", "epi LN bwd(en1)", "wait BT_EN1", "epi de out"]
NS = len(names)
d = t[:, 1:NS + 1] - t[:, 0:NS]
print(f"tiles {tiles}; per-tile cycles mean / median / p90 (last launch = block 0 of the bond predictor backward)")
wait = epi = 0
for i, n in enumerate(names):
    print(f"  {n:34s} {d[:, i].mean():9.0f} {np.median(d[:, i]):9.0f} {np.percentile(d[:, i], 90):9.0f}")
    if n.startswith("wait"):
        wait += d[:, i].mean()
    else:
        epi += d[:, i].mean()
tot = t[:, NS] - t[:, 0]
print(f"  {'total':34s} {tot.mean():9.0f} {np.median(tot):9.0f} {np.percentile(tot, 90):9.0f}   waits {wait:.0f}  epilogues+setup {epi:.0f}")
