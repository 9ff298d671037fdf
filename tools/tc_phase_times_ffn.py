"""GPU tool: per-phase cycle breakdown of tc_bondffn_fwd_kernel (clock64 stamps of row thread 0; config 2 graph)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moldiff_b200 import MolDiff, engine  # noqa: E402
from moldiff_b200.config import builtin_config  # noqa: E402
from moldiff_b200.placeholder import make_data_placeholder  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).to(dev).eval()
np.random.seed(2023)
ph = make_data_placeholder(256, device=dev)
st = model.sample_begin(256, ph["batch_node"], ph["halfedge_index"], ph["batch_halfedge"])
for _ in range(3):
    model.sample_step(st, 500)
E = st["edge_index"].shape[1]
tiles = (E + 127) // 128
buf = torch.zeros(tiles * 32, dtype=torch.int64, device=dev)
lib = engine.load_library()
lib.mdb_debug_set_buffer.argtypes = [C.c_void_p]
lib.mdb_debug_select(2)
lib.mdb_debug_set_buffer(buf.data_ptr())
model.sample_step(st, 499)
torch.cuda.synchronize()
lib.mdb_debug_set_buffer(None)
t = buf.view(tiles, 32).cpu().numpy().astype(np.int64)
order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 14, 15, 16, 17]
names = ["set-up + input planes (h_edge, rbf)", "wait EE (80x64)", "epi e -> ebuf, E planes",
         "L: wait BL+GB", "L: epi *nl, gate LN", "L: wait I1+G2", "L: epi LN(128), sigmoid", "L: wait I2", "L: epi out, RED SL[r]",
         "R: wait BL+GB", "R: epi *nl, gate LN", "R: wait I1+G2", "R: epi LN(128), sigmoid", "R: wait I2", "R: epi out, run-reduce SR"]
d = np.stack([t[:, order[i + 1]] - t[:, order[i]] for i in range(len(order) - 1)], 1)
print(f"tiles {tiles}; per-tile cycles (mean / median / p90), last block's kernel only")
for i, n in enumerate(names):
    print(f"  {n:38s} {d[:, i].mean():9.0f} {np.median(d[:, i]):9.0f} {np.percentile(d[:, i], 90):9.0f}")
tot = t[:, 17] - t[:, 0]
print(f"  {'total':38s} {tot.mean():9.0f} {np.median(tot):9.0f} {np.percentile(tot, 90):9.0f}")
