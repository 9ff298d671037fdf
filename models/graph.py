"""Drop-in for the reference's models/graph.py: the live classes only (NodeEdgeNet and its blocks)."""
from moldiff_b200.nets import BondFFN, EdgeBlock, NodeBlock, NodeEdgeNet, PosUpdate  # noqa: F401
