"""GPU preprocessing utilities with the reference's names and signatures."""
from .directed import (get_appr_directed_adj, get_second_directed_adj,  # noqa: F401
                       directed_features_in_out)

__all__ = ["get_appr_directed_adj", "get_second_directed_adj", "directed_features_in_out"]
