"""Drop-in for the reference's models/bond_predictor.py: `BondPredictor` (moldiff_b200/bond_model.py)."""
from moldiff_b200.bond_model import BondPredictor  # noqa: F401
