"""Drop-in for the live part of the reference's models/diffusion.py (schedules + categorical helpers)."""
from moldiff_b200.schedules import advance_alpha_bar, cosine_betas, get_beta_schedule, segment_alpha_bar  # noqa: F401
from moldiff_b200.transitions import (  # noqa: F401
    categorical_kl,
    gumbel_argmax as log_sample_categorical,
    log_categorical,
    log_onehot as index_to_log_onehot,
    per_item,
)


def extract(coef, t, batch, ndim=2):
    out = per_item(coef, t, batch)
    for _ in range(ndim - 1):
        out = out.unsqueeze(-1)
    if ndim > 3:
        raise NotImplementedError("ndim > 3")
    return out
