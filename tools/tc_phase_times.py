"""GPU tool: per-phase cycle breakdown of tc_nodeblock_fwd_kernel from in-kernel clock64 stamps (config 2 graph)."""
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
lib.mdb_debug_set_buffer(buf.data_ptr())
model.sample_step(st, 499)
torch.cuda.synchronize()
lib.mdb_debug_set_buffer(None)
t = buf.view(tiles, 32).cpu().numpy().astype(np.int64)
order = [0, 1, 6, 2, 7, 3, 8, 4, 9, 5, 15]
names = ["setup(alloc,barriers,sync)", "e tile -> E planes", "wait G1 (64x256)", "epilogue 1 (LN)", "wait G2 (256x256)",
         "epilogue 2 (*hn)", "wait G3+G4", "epilogue 3 (LN+gx)", "wait G5 (256x256)", "epilogue 4 + reduction"]
d = np.stack([t[:, order[i + 1]] - t[:, order[i]] for i in range(len(order) - 1)], 1)
print(f"tiles {tiles}; per-tile cycles (mean / median / p90), last block's kernel only")
for i, n in enumerate(names):
    print(f"  {n:32s} {d[:, i].mean():9.0f} {np.median(d[:, i]):9.0f} {np.percentile(d[:, i], 90):9.0f}")
tot = t[:, 15] - t[:, 0]
print(f"  {'total':32s} {tot.mean():9.0f} {np.median(tot):9.0f} {np.percentile(tot, 90):9.0f}")

# MMA-thread stamps of the LAST sliced GEMM (G5): a_ready group waits and the final commit, relative to the start of epilogue 3
rel = lambda c: (t[:, c] - t[:, 4])
print("G5 (sliced): rows start epi3 = 0; rows end epi3 %.0f; MMA saw group 0/1/2/3 at %.0f / %.0f / %.0f / %.0f; MMA issued last commit %.0f; rows saw done %.0f"
      % (rel(9).mean(), rel(16).mean(), rel(17).mean(), rel(18).mean(), rel(19).mean(), rel(20).mean(), rel(5).mean()))
