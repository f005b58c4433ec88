"""xequinet_b200: B200-native (sm_100a) implementation of XequiNet's XPaiNN message-passing
hot path behind the reference's `xequinet.nn` module API.  See DESIGN.md."""
from . import keys, parallel
from .graph import NeighborGraph, NeighborTransform, SkinNeighborTransform, build_graph, radius_graph, radius_graph_pbc
from .nn import XPaiNN, load_model, resolve_model

__all__ = ["keys", "parallel", "NeighborGraph", "NeighborTransform", "SkinNeighborTransform", "build_graph", "radius_graph", "radius_graph_pbc",
           "XPaiNN", "resolve_model", "load_model"]
