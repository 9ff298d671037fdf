"""CPU study of the tensor-core backward's operand quantisation (no GPU needed).

Runs oracle/blob_emulator.bondpred_forward_backward in float64 with every GEMM that the tensor-core backward kernels
(tc_nodeblock_bwd16, tc_bondffn_bwd) execute replaced by the split-fp16 product the hardware computes:
    A = hi + lo (fp16, hi saturating), W * 256 = Whi + Wlo (fp16),   D = hi Whi + lo Whi + hi Wlo   (fp32-exact here)
under a selectable gradient-scaling policy:  "call" (one power of two per backward call, max |d_logits| -> [8, 16): what
round 1 shipped), "row" (one power of two per tile row, chosen from that row's amax at every operand store).
Prints the per-molecule gradient error of each policy against the unquantised float64 chain.

    python tools/emulate_tc_backward.py [B]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from moldiff_b200 import packing  # noqa: E402
from oracle import blob_emulator as BE  # noqa: E402
from tests.helpers import batch_inputs, doubled, per_molecule_rel_err  # noqa: E402

F16_MAX = 65504.0


def split16(x):
    hi = x.clamp(-F16_MAX, F16_MAX).to(torch.float16).to(x.dtype)
    lo = (x - hi).to(torch.float16).to(x.dtype)
    return hi, lo


def chop32(x, rz):
    """float64 -> nearest (rz False) or truncated (rz True) float32 value, kept in float64."""
    if not rz:
        return x.float().double()
    m, e = torch.frexp(x)
    return torch.ldexp(torch.trunc(m * 16777216.0) / 16777216.0, e)


def accumulate(hi, lo, whi, wlo, acc):
    """Sum the three split products K-step by K-step (16 columns per tcgen05.mma) with an fp32 accumulator that rounds
    ('rn32') or truncates ('rz32') after every instruction; 'exact' keeps float64."""
    if acc == "exact":
        return hi @ whi + lo @ whi + hi @ wlo
    if acc.startswith("offset"):
        # fixed-point accumulation: the accumulator starts at 1.5 C (C a power of two per row, >= slack * 2 max|result|), so it
        # stays inside [C, 2C) where every truncation drops a uniform fraction of the constant quantum u = C 2^-23; the
        # expected loss n u / 2 is added back -> zero-mean error
        slack = float(acc[6:] or 1)
        exact = hi @ whi + lo @ whi + hi @ wlo
        Cr = torch.exp2(torch.ceil(torch.log2(exact.abs().amax(-1, keepdim=True).clamp_min(1e-30) * 2 * slack)))
        u = Cr * 2.0 ** -23
        d = 1.5 * Cr.expand_as(exact).clone()
        n = 0
        for k in range(0, hi.shape[1], 16):
            for a, w in ((hi, whi), (lo, whi), (hi, wlo)):
                d = torch.floor((d + a[:, k:k + 16] @ w[k:k + 16]) / u) * u
                n += 1
        return d - 1.5 * Cr + n * u / 2
    rz = acc.startswith("rz32")
    d = torch.zeros(hi.shape[0], whi.shape[1], dtype=hi.dtype)
    if acc.endswith("crossfirst"):      # all cross-term MMAs first (small accumulator), then the 16 hi*hi MMAs
        for k in range(0, hi.shape[1], 16):
            for a, w in ((lo, whi), (hi, wlo)):
                d = chop32(d + a[:, k:k + 16] @ w[k:k + 16], rz)
        for k in range(0, hi.shape[1], 16):
            d = chop32(d + hi[:, k:k + 16] @ whi[k:k + 16], rz)
        return d
    for k in range(0, hi.shape[1], 16):
        for a, w in ((hi, whi), (lo, whi), (hi, wlo)):
            d = chop32(d + a[:, k:k + 16] @ w[k:k + 16], rz)
    return d


class QMM:
    """a @ w with the kernel's operand quantisation.  mode: None (exact), 'call', 'row'."""

    def __init__(self, mode, stats=None, acc="exact"):
        self.mode, self.stats, self.acc = mode, stats, acc

    def __call__(self, a, w, grad):
        if self.mode is None:
            return a @ w
        whi, wlo = split16(w * 256.0)
        if grad and self.mode == "row":
            amax = a.abs().amax(-1, keepdim=True).clamp_min(1e-300)
            s = torch.exp2(torch.floor(torch.log2(1024.0 / amax)))     # row amax -> [512, 1024]
            a = a * s
        else:
            s = None
        if self.stats is not None and grad:
            self.stats.append(float(a.abs().amax(-1).median()))
        hi, lo = split16(a)
        out = accumulate(hi, lo, whi, wlo, self.acc) / 256.0
        return out / s if s is not None else out


def run(B=16, gui="uncertainty"):
    torch.set_default_dtype(torch.float64)
    from moldiff_b200 import BondPredictor
    from moldiff_b200.config import builtin_config
    torch.manual_seed(0)
    torch.set_default_dtype(torch.float32)
    bp = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).eval()
    torch.set_default_dtype(torch.float64)
    sd = bp.state_dict()
    inp = batch_inputs(B=B, t_values=(999, 400, 0, 650))
    ei, be, _ = doubled(inp)
    blob, ho, bo = packing.pack_network(sd, kind=2, net_prefix="encoder", num_blocks=8, update_pos=False, time_dim=20)
    W = BE.Blob(blob.double(), ho, bo)
    kw = dict(num_blocks=8, rbf_lo=0.0, rbf_hi=20.0, time_dim=20, T=1000.0, kn=8, ke=5, h_node_in=inp["h_node"].double(),
              pos=inp["pos"].double(), edge_index=ei, batch_node=inp["batch_node"], batch_edge=be, t=inp["t"])
    # upstream gradient of the guidance objective
    with torch.no_grad():
        logits, _ = BE.bondpred_forward_backward(W, d_logits=torch.zeros(ei.shape[1] // 2, 5), **kw)
    lg = logits.clone().requires_grad_(True)
    if gui == "uncertainty":
        obj = torch.sigmoid(-torch.logsumexp(lg, dim=-1)).log().sum()
    else:
        prob = torch.softmax(lg, dim=-1)
        obj = (-torch.sum(prob * torch.log(prob + 1e-12), dim=-1)).log().sum()
    d_logits = torch.autograd.grad(obj, lg)[0]
    k = 3 - int(torch.floor(torch.log2(d_logits.abs().max())))
    res = {}
    acc = os.environ.get("ACC", "exact")
    qfwd = os.environ.get("QFWD", "0") == "1"
    for mode in (None, "call") + (("row",) if os.environ.get("ROW", "0") == "1" else ()):
        stats = []
        BE.set_backward_mm(QMM(mode, stats, acc))
        if qfwd and mode is not None:
            fq = QMM("call", None, acc)
            fq_exact = QMM("call", None, "exact")
            cats = [c for c in os.environ.get("CAT", "").split(",") if c]
            names = {}
            for i in range(8):
                for j, nm in enumerate(packing.BLOCK_SLOTS):
                    if bo[i][j] >= 0:
                        names[bo[i][j]] = nm
            for j, nm in enumerate(packing.HEAD_SLOTS):
                if ho[j] >= 0:
                    names[ho[j]] = nm

            def fhook(a, w):
                nm = names.get(w.storage_offset(), "?")
                hit = not cats or any(nm.startswith(c) for c in cats)
                return (fq if hit else fq_exact)(a, w, False)
            BE.set_forward_mm(fhook)
        try:
            with torch.no_grad():
                scale = 2.0 ** k if mode == "call" else 1.0
                _, d_pos = BE.bondpred_forward_backward(W, d_logits=d_logits * scale, **kw)
                res[mode] = d_pos / scale
        finally:
            BE.set_backward_mm(None)
            BE.set_forward_mm(None)
        if mode is not None:
            st = torch.tensor(stats)
            print(f"mode {mode}: median row amax of gradient operands, min/median/max over GEMMs: "
                  f"{st.min():.2e} {st.median():.2e} {st.max():.2e}")
    for mode in [m for m in res if m is not None]:
        e = per_molecule_rel_err(res[mode], res[None], inp["batch_node"])
        print(f"{gui} B={B} acc={acc} qfwd={qfwd} policy={mode}: per-molecule gradient error median {e.median():.2e} max {e.max():.2e}")
    torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    for gui in os.environ.get("GUI", "uncertainty,entropy").split(","):
        run(B, gui)
