"""Forward-noising and reverse-posterior operators of the joint diffusion (host-side PyTorch).

`GaussianTransition` handles continuous variables (atom positions, or one-hot types in the
`continuous` categorical space); `CategoricalTransition` handles discrete types with an absorbing /
near-uniform stationary distribution.  They are exported under the reference's class names
(`ContigousTransition`, `GeneralCategoricalTransition`) by ``models/transition.py`` and keep the
reference's ``state_dict`` entries -- frozen ``nn.Parameter`` tables named ``betas, alphas, alphas_bar,
alphas_bar_prev, coef_x0, coef_xt, std`` and ``q_mats, transpopse_q_onestep_mats`` (sic) -- so reference
checkpoints load strictly.  Semantics follow ``models/transition.py:9-69,178-339`` and
``models/diffusion.py:41-92``; per north_star this layer stays PyTorch (it is ~0 FLOPs), the CUDA work
is the denoiser it surrounds.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LOG_FLOOR = -32.0
TINY = 1e-30


def _frozen(x):
    return nn.Parameter(torch.from_numpy(np.asarray(x)).float(), requires_grad=False)


def per_item(table, t, batch):
    """table[t][batch]: per-graph timestep -> per-node/edge coefficient rows (diffusion.py:60-72)."""
    return table[t][batch]


def gumbel_argmax(logits):
    """Sample class ids from unnormalised log-probabilities (diffusion.py:79-85)."""
    u = torch.rand_like(logits)
    g = -torch.log(-torch.log(u + TINY) + TINY)
    return (g + logits).argmax(dim=-1)


def log_onehot(idx, num_classes):
    if int(idx.max()) >= num_classes:
        raise AssertionError(f"Error: {int(idx.max())} >= {num_classes}")
    return torch.log(F.one_hot(idx, num_classes).float().clamp(min=TINY))


def categorical_kl(log_p, log_q):
    return (log_p.exp() * (log_p - log_q)).sum(dim=-1)


def log_categorical(log_x0, log_prob):
    return (log_x0.exp() * log_prob).sum(dim=-1)


class GaussianTransition(nn.Module):
    """x_t = sqrt(abar_t) x_0 + sqrt(1-abar_t) eps, and the DDPM posterior q(x_{t-1} | x_t, x_0)."""

    def __init__(self, betas, num_classes=None, scaling=1.0):
        super().__init__()
        self.num_classes = num_classes
        self.scaling = scaling
        alphas = 1.0 - betas
        abar = np.cumprod(alphas, axis=0)
        abar_prev = np.concatenate([[1.0], abar[:-1]])
        self.betas = _frozen(betas)
        self.alphas = _frozen(alphas)
        self.alphas_bar = _frozen(abar)
        self.alphas_bar_prev = _frozen(abar_prev)
        self.coef_x0 = _frozen(np.sqrt(abar_prev) * betas / (1 - abar))
        self.coef_xt = _frozen(np.sqrt(alphas) * (1 - abar_prev) / (1 - abar))
        self.std = _frozen(np.sqrt((1 - abar_prev) * betas / (1 - abar)))

    def add_noise(self, x, time_step, batch):
        if self.num_classes is not None:
            x = F.one_hot(x, self.num_classes).float()
        x = x / self.scaling
        abar = self.alphas_bar.index_select(0, time_step).index_select(0, batch).unsqueeze(-1)
        eps = torch.zeros_like(x).normal_()
        pert = abar.sqrt() * x + (1 - abar).sqrt() * eps
        return pert if self.num_classes is None else (pert, x)

    def get_prev_from_recon(self, x_t, x_recon, t, batch):
        mu = (per_item(self.coef_x0, t, batch).unsqueeze(-1) * x_recon
              + per_item(self.coef_xt, t, batch).unsqueeze(-1) * x_t)
        sigma = per_item(self.std, t, batch).unsqueeze(-1)
        x_prev = mu + sigma * torch.randn_like(mu)
        return torch.where((t[batch] == 0).unsqueeze(-1), mu, x_prev)

    def sample_init(self, shape):
        # The reference draws from the *CPU* generator and then moves (transition.py:65-69); kept so
        # that seeded runs start from the same x_T on any device.
        if self.num_classes is None:
            return torch.randn(shape).to(self.betas.device)
        return torch.randn([shape, self.num_classes]).to(self.betas.device)


def _stationary(init_prob, k):
    if init_prob is None or (isinstance(init_prob, str) and init_prob == "uniform"):
        return np.ones(k) / k
    if isinstance(init_prob, str):
        if init_prob == "absorb":      # everything drains into class 0 ("no bond")
            p = 0.01 * np.ones(k)
            p[0] = 1
        elif init_prob == "tomask":    # everything drains into the last class (mask type)
            p = 0.001 * np.ones(k)
            p[-1] = 1.0
        else:
            raise ValueError(init_prob)
        return p / np.sum(p)
    p = np.asarray(init_prob, dtype=np.float64)
    return p / np.sum(p)


class CategoricalTransition(nn.Module):
    """D3PM-style categorical diffusion with one-step kernels Q_t = (1-beta_t) I + beta_t 1 pi^T."""

    def __init__(self, betas, num_classes, init_prob=None):
        super().__init__()
        self.eps = TINY
        self.num_classes = num_classes
        self.init_prob = _stationary(init_prob, num_classes)
        self.betas = betas
        self.num_timesteps = len(betas)
        k = num_classes
        one_step = np.stack([b * np.repeat(self.init_prob[None, :], k, axis=0) + np.eye(k) * (1.0 - b)
                             for b in betas], axis=0)                       # (T, K, K)
        cumulative = [one_step[0]]
        for t in range(1, self.num_timesteps):
            cumulative.append(np.tensordot(cumulative[-1], one_step[t], axes=[[1], [0]]))
        self.q_mats = _frozen(np.stack(cumulative, axis=0))
        self.transpopse_q_onestep_mats = _frozen(np.transpose(one_step, axes=[0, 2, 1]))

    # ---- forward process ----
    def onehot_encode(self, v):
        return F.one_hot(v, self.num_classes).float()

    def q_vt_pred(self, log_v0, t, batch):
        q = torch.einsum("...i,...ij->...j", log_v0.exp(), per_item(self.q_mats, t, batch))
        return torch.log(q + self.eps).clamp_min(LOG_FLOOR)

    def q_vt_sample(self, log_v0, t, batch):
        cls = gumbel_argmax(self.q_vt_pred(log_v0, t, batch))
        return cls, log_onehot(cls, self.num_classes)

    def add_noise(self, v, time_step, batch):
        log_v0 = log_onehot(v, self.num_classes)
        cls, log_vt = self.q_vt_sample(log_v0, time_step, batch)
        return self.onehot_encode(cls), log_vt, log_v0

    # ---- reverse process ----
    def q_v_posterior(self, log_v0, log_vt, t, batch, v0_prob):
        """log q(v_{t-1} | v_t, v_0) up to the t==0 convention (returns log_v0 there)."""
        tm1 = torch.clamp(t - 1, min=0)
        step_t = per_item(self.transpopse_q_onestep_mats, t, batch)
        f1 = torch.einsum("bj,bjk->bk", torch.exp(log_vt), step_t)
        cum = per_item(self.q_mats, tm1, batch)
        if v0_prob:
            f2 = torch.einsum("bj,bjk->bk", torch.exp(log_v0), cum)
        else:
            cls0 = log_v0.argmax(dim=-1)
            f2 = cum[torch.arange(len(cls0), device=cls0.device), cls0]
        if log_v0.ndim == 2:
            tt = t[batch].unsqueeze(-1)
        elif log_v0.ndim == 3:
            tt = t[batch].unsqueeze(-1).unsqueeze(-1)
        else:
            raise NotImplementedError("ndim not supported")
        out = torch.log(f1 + self.eps).clamp_min(LOG_FLOOR) + torch.log(f2 + self.eps).clamp_min(LOG_FLOOR)
        out = out - torch.logsumexp(out, dim=-1, keepdim=True)
        return torch.where(tt == 0, log_v0, out)

    def compute_v_Lt(self, log_v_post_true, log_v_post_pred, log_v0, t, batch):
        kl = categorical_kl(log_v_post_true, log_v_post_pred)
        nll = -log_categorical(log_v0, log_v_post_pred)
        mask = (t == 0).float()[batch]
        if log_v_post_true.ndim == 3:
            mask = mask.unsqueeze(-1)
        elif log_v_post_true.ndim != 2:
            raise NotImplementedError("ndim not supported")
        return mask * nll + (1 - mask) * kl

    def sample_init(self, n):
        log_pi = torch.log(torch.from_numpy(self.init_prob) + self.eps).clamp_min(LOG_FLOOR).to(self.q_mats.device)
        log_pi = log_pi.unsqueeze(0).repeat(n, 1)
        cls = gumbel_argmax(log_pi)
        return cls, self.onehot_encode(cls), log_onehot(cls, self.num_classes)
