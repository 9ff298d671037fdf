"""GPU debugging aid: per-block error of the CUDA NodeEdgeNet against the oracle trace (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moldiff_b200 import MolDiff, engine  # noqa: E402
from moldiff_b200.config import builtin_config  # noqa: E402
from oracle import restatement as R  # noqa: E402
from tests.helpers import batch_inputs, doubled  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).eval()
    sd = {"net." + k: v for k, v in model.denoiser.state_dict().items()}
    inp = batch_inputs(B=3, t_values=(100, 600, 900))
    ei, be, _ = doubled(inp)
    g = torch.Generator().manual_seed(9)
    N, E = len(inp["batch_node"]), ei.shape[1]
    h_node, h_edge = torch.randn(N, 256, generator=g), torch.randn(E, 64, generator=g)
    nt = (inp["t"][inp["batch_node"]].float() / 1000).unsqueeze(-1)
    et = (inp["t"][be].float() / 1000).unsqueeze(-1)
    trace = []
    with torch.no_grad():
        R.node_edge_net(sd, "net", h_node, inp["pos"], h_edge, ei, nt, et, num_blocks=6, cutoff=15.0, trace=trace)
    plan = engine.GraphPlan(ei.to(dev), N)
    for k in range(1, 7):
        net = engine.PackedNet(sd, kind=0, net_prefix="net", num_blocks=k, update_pos=True, cutoff=15.0, device=dev)
        out = engine.net_forward(net, plan, h_node.to(dev), inp["pos"].to(dev), h_edge.to(dev), nt.to(dev), et.to(dev))
        torch.cuda.synchronize()
        errs = [R.rel_err(a.cpu(), b) for a, b in zip(out, trace[k - 1])]
        print(f"blocks={k}: h_node {errs[0]:.2e}  pos {errs[1]:.2e}  h_edge {errs[2]:.2e}")


if __name__ == "__main__":
    main()
