"""Drop-in for the live part of the reference's models/common.py (MLP, GaussianSmearing)."""
from moldiff_b200.nets import MLP, GaussianSmearing  # noqa: F401
