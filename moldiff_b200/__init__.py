"""moldiff_b200 -- B200-native (sm_100a) implementation of MolDiff's diffusion denoising hot path.

Public surface mirrors the reference's (`models.model.MolDiff`, `models.bond_predictor.BondPredictor`,
`models.graph.NodeEdgeNet`); the top-level `models/` package of this repo re-exports these classes under
the reference's module paths so `scripts/sample_drug3d.py`-style callers run unchanged.
"""
from .bond_model import BondPredictor
from .diffusion_model import MolDiff
from .nets import MLP, GaussianSmearing, NodeEdgeNet

__all__ = ["MolDiff", "BondPredictor", "NodeEdgeNet", "MLP", "GaussianSmearing"]
