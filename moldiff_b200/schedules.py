"""Noise schedules for the three diffusion processes (positions, atom types, bond types).

Host-side numpy, float64, evaluated once at model construction.  Behavioural contract (values, clipping,
argument names as they appear in the YAML configs) follows the reference's
``models/diffusion.py:97-192`` (`get_beta_schedule` and the `advance` / `segment` / `cosine` families);
the implementation here is organised around cumulative-alpha curves instead.
"""
from __future__ import annotations

import numpy as np


def _logistic(x):
    return 1.0 / (1.0 + np.exp(-x))


def _betas_from_alpha_bar(alpha_bar):
    """beta_t = 1 - abar_t / abar_{t-1} (abar_{-1} := 1), clipped to [0, 1]."""
    prev = np.concatenate([[1.0], alpha_bar[:-1]])
    return np.clip(1.0 - alpha_bar / prev, 0.0, 1.0)


def advance_alpha_bar(num_timesteps, scale_start, scale_end, width):
    """Sigmoid-shaped cumulative alpha going from `scale_start` (t=0) to `scale_end` (t=T-1)
    (reference: diffusion.py:110-132).  abar(x) = a*sigmoid(-k x) + b on x in [-1, 1]."""
    k = width
    a = (scale_end - scale_start) / (_logistic(-k) - _logistic(k))
    b = 0.5 * (scale_end + scale_start - a)
    x = np.linspace(-1.0, 1.0, num_timesteps)
    return a * _logistic(-k * x) + b


def segment_alpha_bar(num_timesteps, time_segment, segment_diff):
    """Piecewise `advance` curves; each piece is evaluated on len+1 points and its first point dropped
    (reference: diffusion.py:134-151)."""
    if int(np.sum(time_segment)) != num_timesteps:
        raise AssertionError("time_segment must sum to num_timesteps")
    pieces = []
    for seg_len, params in zip(time_segment, segment_diff):
        curve = advance_alpha_bar(seg_len + 1, params["scale_start"], params["scale_end"], params["width"])
        pieces.append(curve[1:])
    return np.concatenate(pieces)


def cosine_betas(num_timesteps, s=0.008):
    """Nichol & Dhariwal cosine schedule (reference: diffusion.py:97-108)."""
    n = num_timesteps + 1
    x = np.linspace(0, n, n)
    abar = np.cos(((x / n) + s) / (1 + s) * np.pi * 0.5) ** 2
    abar = abar / abar[0]
    return np.clip(1.0 - abar[1:] / abar[:-1], 0.0, 0.999)


def get_beta_schedule(beta_schedule, num_timesteps, **kw):
    """Dispatch on the YAML `beta_schedule` key; extra YAML keys (e.g. `init_prob`) are ignored, as in
    the reference where the whole `diff_*` mapping is splatted in (model.py:60-63)."""
    T = num_timesteps
    if beta_schedule == "advance":
        betas = _betas_from_alpha_bar(advance_alpha_bar(
            T, kw.get("scale_start", 0.999), kw.get("scale_end", 0.001), kw.get("width", 2)))
    elif beta_schedule == "segment":
        betas = _betas_from_alpha_bar(segment_alpha_bar(T, kw["time_segment"], kw["segment_diff"]))
    elif beta_schedule == "cosine":
        betas = cosine_betas(T, kw.get("s", 0.008))
    elif beta_schedule == "linear":
        betas = np.linspace(kw["beta_start"], kw["beta_end"], T, dtype=np.float64)
    elif beta_schedule == "quad":
        betas = np.linspace(kw["beta_start"] ** 0.5, kw["beta_end"] ** 0.5, T, dtype=np.float64) ** 2
    elif beta_schedule == "const":
        betas = kw["beta_end"] * np.ones(T, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(T, 1, T, dtype=np.float64)
    elif beta_schedule == "sigmoid":
        s = kw.get("s", 6)
        betas = _logistic(np.linspace(-s, s, T)) * (kw["beta_end"] - kw["beta_start"]) + kw["beta_start"]
    else:
        raise NotImplementedError(beta_schedule)
    if betas.shape != (T,):
        raise AssertionError("schedule length mismatch")
    return betas
