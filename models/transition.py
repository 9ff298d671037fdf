"""Drop-in for the reference's models/transition.py under its class names."""
from moldiff_b200.transitions import (  # noqa: F401
    CategoricalTransition as GeneralCategoricalTransition,
    GaussianTransition as ContigousTransition,
)
