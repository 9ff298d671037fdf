"""GPU tool: A/B of tc_bondffn_bwd2_kernel against tc_bondffn_bwd_kernel inside one process (MDB_TC_FFN_BWD2 is read per call):
one FFMA or tensor-core forward (save), then the backward twice per repetition; per-molecule relative difference of d/dpos
between the two kernels and against the fp64 oracle gradient.  usage: python tools/ffn_bwd_ab.py [ff|tc] [reps]"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moldiff_b200 import BondPredictor, engine
from moldiff_b200.config import builtin_config
from tests.helpers import batch_inputs, doubled, to_dev, per_molecule_rel_err

libc = ctypes.CDLL(None)
def setenv(k, v): libc.setenv(k.encode(), v.encode(), 1); os.environ[k] = v

fwd = sys.argv[1] if len(sys.argv) > 1 else "ff"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
torch.manual_seed(0)
bp = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).eval().to(dev)
nets = {}
for name, dis in (("tc", "0"), ("ff", "1")):
    setenv("MDB_DISABLE_TC", dis)
    nets[name] = bp._pack(dev)
setenv("MDB_DISABLE_TC", "0")
ref = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bondpred_ref64.pt"), weights_only=False)["B16"]
inp = batch_inputs(**ref["args"])
d = to_dev(inp, dev)
ei, be, _ = doubled(d)
plan = engine.plan_for(ei, d["h_node"].shape[0])

def objective(lg):                       # tests/test_gpu_parity.py::_objective("entropy")
    prob = torch.softmax(lg, dim=-1)
    return (-torch.sum(prob * torch.log(prob + 1e-12), dim=-1)).log().sum()

for r in range(reps):
    logits = engine.bondpred_forward(nets[fwd], plan, d["h_node"], d["pos"], d["batch_node"], be, d["t"], save=True)
    lg = logits.detach().clone().requires_grad_(True)
    dl = torch.autograd.grad(objective(lg), lg)[0]
    g = {}
    for v in ("1", "0", "1"):
        setenv("MDB_TC_FFN_BWD2", v)
        gr = engine.bondpred_backward(nets["tc"], plan, d["h_node"], d["pos"], d["batch_node"], be, d["t"], dl)
        torch.cuda.synchronize()
        g.setdefault(v, []).append(gr.cpu().double())
    e_new = per_molecule_rel_err(g["1"][0], ref["entropy"], inp["batch_node"])
    e_old = per_molecule_rel_err(g["0"][0], ref["entropy"], inp["batch_node"])
    ab = per_molecule_rel_err(g["1"][0], g["0"][0], inp["batch_node"])
    rep = per_molecule_rel_err(g["1"][0], g["1"][1], inp["batch_node"])
    print(f"rep {r:2d}: median err new {float(e_new.median()):.2e} old {float(e_old.median()):.2e} | new vs old max {float(ab.max()):.2e} "
          f"median {float(ab.median()):.2e} | new vs new (2nd run) max {float(rep.max()):.2e} | #>1e-4 new {int((e_new > 1e-4).sum())} old {int((e_old > 1e-4).sum())}")
