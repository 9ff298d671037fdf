"""Reference-compatible module paths (`from models.model import MolDiff`, ...): thin re-exports of
moldiff_b200 so that the reference's scripts/sample_drug3d.py and scripts/train_drug3d.py import the
B200-native implementation when run with this repository as the working directory."""
