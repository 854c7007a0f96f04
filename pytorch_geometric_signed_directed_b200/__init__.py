"""pytorch_geometric_signed_directed_b200 -- B200-native (sm_100a) implementation of the
signed/directed message-passing hot path of PyTorch Geometric Signed Directed, behind the
reference's own conv-layer API.  See DESIGN.md."""
from . import nn  # noqa: F401
from .nn import *  # noqa: F401,F403

__version__ = "0.1.0"
