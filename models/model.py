"""Drop-in for the reference's models/model.py: `MolDiff` (B200-native, see moldiff_b200/diffusion_model.py)."""
from moldiff_b200.diffusion_model import MolDiff  # noqa: F401
from moldiff_b200.nets import MLP, GaussianSmearing, NodeEdgeNet  # noqa: F401
from moldiff_b200.transitions import (  # noqa: F401
    CategoricalTransition as GeneralCategoricalTransition,
    GaussianTransition as ContigousTransition,
)
