"""GPU study of the tensor-core numerics behind the guidance-gradient gap.

Part 1  accumulator behaviour of tcgen05.mma kind::f16 (fp32 accumulate) through mdb_tc_selftest: crafted vectors that
        tell round-to-nearest from truncation, count the alignment guard bits, and the signed error statistics of a random
        K = 256 GEMM whose operands are exactly representable in fp16 (so that ONLY the accumulation can err).
Part 2  crossed paths on the reference golden case bondpred/B16 (and a larger batch): forward and backward are run on the
        tensor-core kernels or on the fp32 FFMA kernels independently (two packings of the same weights on one workspace),
        per-molecule gradient error against the fp64 oracle.

    python tools/tc_numerics.py [out.json]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from moldiff_b200 import BondPredictor, engine  # noqa: E402
from moldiff_b200.config import builtin_config  # noqa: E402
from tests.helpers import batch_inputs, doubled, per_molecule_rel_err, to_dev  # noqa: E402

dev = torch.device("cuda:0")
out = {"variant": os.environ.get("MDB_LIB_VARIANT", "")}


XF = False     # accumulation order of the self-test GEMMs (part 1 runs both)


def selftest(x, w, twice=False):
    return engine.tc_selftest(x.to(dev), w, twice=twice, cross_first=XF).double().cpu()


def part1():
    global XF
    K, N = 64, 32
    # y = sum_k x_k w_k.  Rows k % 16 in 1..8 of W are "small slots" (w = 2^-12), the others "big slots" (w = 1), and x is
    # scaled by 2^10, so every operand below is exactly representable in fp16 (normal range) down to products of 2^-33.
    SMALL_W, scale = 2.0 ** -12, 2.0 ** 10
    w = torch.ones(K, N)
    for k in range(K):
        if 1 <= k % 16 <= 8:
            w[k] = SMALL_W
    rows = []

    def case(name, big, small, exact):
        """big: [(k-step, value)] placed in big slots; small: [(k-step, value)] placed in small slots of that step."""
        x = torch.zeros(K)
        used = {}
        for st, v in big:
            j = used.get(("b", st), 0)
            used[("b", st)] = j + 1
            x[st * 16 + (0, 9, 10, 11)[j]] = v
        for st, v in small:
            j = used.get(("s", st), 0)
            used[("s", st)] = j + 1
            x[st * 16 + 1 + j] = v / SMALL_W
        rows.append((name, x, exact))

    u = 2.0 ** -23                              # ulp(1)
    case("same-step 1 + 2 x 2^-24", [(0, 1.0)], [(0, 2.0 ** -24)] * 2, 1 + u)
    case("same-step 1 + 4 x 2^-25", [(0, 1.0)], [(0, 2.0 ** -25)] * 4, 1 + u)
    case("same-step 1 + 8 x 2^-26", [(0, 1.0)], [(0, 2.0 ** -26)] * 8, 1 + u)
    case("same-step 1 + 1.5 x 2^-24", [(0, 1.0)], [(0, 1.5 * 2.0 ** -24)], 1 + 0.75 * u)
    case("same-step -1 - 1.5 x 2^-24", [(0, -1.0)], [(0, -1.5 * 2.0 ** -24)], -(1 + 0.75 * u))
    case("same-step 1 - 2^-25", [(0, 1.0)], [(0, -(2.0 ** -25))], 1 - 2.0 ** -25)
    case("same-step 1 - 1.5 x 2^-25", [(0, 1.0)], [(0, -1.5 * 2.0 ** -25)], 1 - 1.5 * 2.0 ** -25)
    case("same-step 1 + 0.5 x 2^-24 (x8)", [(0, 1.0)], [(0, 0.5 * 2.0 ** -24)] * 8, 1 + 2 * u)
    case("next-step 1 + 1.5 x 2^-24", [(0, 1.0)], [(1, 1.5 * 2.0 ** -24)], 1 + 0.75 * u)
    case("next-step -1 - 1.5 x 2^-24", [(0, -1.0)], [(1, -1.5 * 2.0 ** -24)], -(1 + 0.75 * u))
    case("next-step 1 - 2^-25", [(0, 1.0)], [(1, -(2.0 ** -25))], 1 - 2.0 ** -25)
    case("next-step 1 + 3 x (1.5 x 2^-24)", [(0, 1.0)], [(1, 1.5 * 2.0 ** -24), (2, 1.5 * 2.0 ** -24), (3, 1.5 * 2.0 ** -24)], 1 + 2.25 * u)
    for sft in range(24, 34):
        case(f"same-step 1 - 1 + 2^-{sft}", [(0, 1.0), (0, -1.0)], [(0, 2.0 ** -sft)], 2.0 ** -sft)
    for sft in range(24, 34):
        case(f"next-step (1 + 2^-{sft}) - 1", [(0, 1.0), (1, -1.0)], [(0, 2.0 ** -sft)], 2.0 ** -sft)
    x = torch.zeros(128, K)
    for i, (_, v, _) in enumerate(rows):
        x[i] = v
    y = selftest(x * scale, w)[:, 0] / scale
    res = []
    for i, (name, v, exact) in enumerate(rows):
        xin = (v * scale)
        rep = bool(torch.all(xin.half().float() == xin)) and bool(torch.all((w * 256).half().float() == w * 256))
        chk = float((v.double() * w[:, 0].double()).sum())
        assert abs(chk - exact) <= 1e-12 * max(1.0, abs(exact)), (name, chk, exact)
        res.append({"case": name, "exact": exact, "got": float(y[i]), "err_ulp1": (float(y[i]) - exact) / u, "operands_fp16_exact": rep})
        print(f"{name:36s} exact {exact!r:24} got {float(y[i])!r:24} err/ulp(1) {(float(y[i]) - exact) / u:+.3f} fp16-exact-operands {rep}")
    out[f"crafted_xf{int(XF)}"] = res

    # random GEMMs with fp16-exact operands: only the accumulation errs
    g = torch.Generator().manual_seed(5)
    stats = {}
    print(f"---- random GEMMs, cross-first = {XF}")
    for (K2, N2) in ((256, 256), (64, 256), (128, 128)):
        xr = torch.randn(128, K2, generator=g).half().float()
        wr = ((torch.rand(K2, N2, generator=g) * 2 - 1) / K2 ** 0.5 * 256).half().float() / 256      # w * 256 exact in fp16
        exact = xr.double() @ wr.double()
        for twice in (False, True):
            y = selftest(xr, wr, twice=twice)
            ex = exact * (2 if twice else 1)
            ulp = torch.exp2(torch.floor(torch.log2(ex.abs().clamp_min(1e-30))) - 23)
            e = (y - ex) / ulp
            shrink = ((y.abs() - ex.abs()) / ulp)
            key = f"K{K2}_N{N2}{'_twice' if twice else ''}"
            stats[key] = {"mean_err_ulp": float(e.mean()), "rms_err_ulp": float(e.pow(2).mean().sqrt()),
                          "mean_shrink_ulp": float(shrink.mean()), "max_abs_err_ulp": float(e.abs().max()),
                          "rel_err_max": float((y - ex).abs().max() / ex.abs().max())}
            print(key, stats[key])
        # positive-only operands: partial sums grow monotonically
        xp, wp = xr.abs(), wr.abs()
        exact = xp.double() @ wp.double()
        y = selftest(xp, wp)
        ulp = torch.exp2(torch.floor(torch.log2(exact)) - 23)
        e = (y - exact) / ulp
        stats[f"K{K2}_N{N2}_positive"] = {"mean_err_ulp": float(e.mean()), "rms_err_ulp": float(e.pow(2).mean().sqrt())}
        print(f"K{K2}_N{N2}_positive", stats[f"K{K2}_N{N2}_positive"])
        # generic fp32 operands (hi | lo split active): total error of the 3-MMA scheme
        xg = torch.randn(128, K2, generator=g)
        wg = (torch.rand(K2, N2, generator=g) * 2 - 1) / K2 ** 0.5
        exact = xg.double() @ wg.double()
        y = selftest(xg, wg)
        y32 = (xg.to(dev) @ wg.to(dev)).double().cpu()
        stats[f"K{K2}_N{N2}_generic"] = {"tc_rel_err": float((y - exact).abs().max() / exact.abs().max()),
                                        "tc_mean_signed_shrink_rel": float(((y.abs() - exact.abs()) / exact.abs().max()).mean()),
                                        "fp32_matmul_rel_err": float((y32 - exact).abs().max() / exact.abs().max())}
        print(f"K{K2}_N{N2}_generic", stats[f"K{K2}_N{N2}_generic"])
    out[f"random_xf{int(XF)}"] = stats


def objective(logits, gui):
    if gui == "uncertainty":
        return torch.sigmoid(-torch.logsumexp(logits, dim=-1)).log().sum()
    prob = torch.softmax(logits, dim=-1)
    return (-torch.sum(prob * torch.log(prob + 1e-12), dim=-1)).log().sum()


def part2():
    ref = torch.load(os.path.join(ROOT, "tests", "golden", "bondpred_ref64.pt"), weights_only=False)
    torch.manual_seed(0)
    bp = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).eval().to(dev)
    nets = {}
    for name, dis in (("tc", "0"), ("ff", "1")):
        os.environ["MDB_DISABLE_TC"] = dis
        nets[name] = bp._pack(dev)
    os.environ.pop("MDB_DISABLE_TC")
    res = {}
    for cname, case in ref.items():
        inp = batch_inputs(**case["args"])
        d = to_dev(inp, dev)
        ei, be, _ = doubled(d)
        plan = engine.plan_for(ei, d["h_node"].shape[0])
        for gui in ("uncertainty", "entropy"):
            g64 = case[gui]
            for f in ("tc", "ff"):
                for b in ("tc", "ff"):
                    logits = engine.bondpred_forward(nets[f], plan, d["h_node"], d["pos"], d["batch_node"], be, d["t"], save=True)
                    lg = logits.detach().clone().requires_grad_(True)
                    dl = torch.autograd.grad(objective(lg, gui), lg)[0]
                    grad = engine.bondpred_backward(nets[b], plan, d["h_node"], d["pos"], d["batch_node"], be, d["t"], dl)
                    torch.cuda.synchronize()
                    e = per_molecule_rel_err(grad.cpu().double(), g64, inp["batch_node"])
                    le = float((logits.cpu().double() - case["logits"]).abs().max() / case["logits"].abs().max())
                    key = f"{cname}/{gui}/fwd={f}/bwd={b}"
                    res[key] = {"median": float(e.median()), "max": float(e.max()), "frac_lt_1e-4": float((e < 1e-4).float().mean()),
                                "logits_rel_err": le}
                    print(key, res[key])
    out["crossed"] = res


if __name__ == "__main__":
    part1()
    XF = True
    part1()
    part2()
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)
