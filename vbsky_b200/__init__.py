"""vbsky_b200 -- B200-native (sm_100a) implementation of VBSKY's data-parallel hot path:
Felsenstein pruning likelihood + gradient, BDSKY tree prior + gradient and the height-transform
chain, behind the C ABI of include/vbsky_b200.h.  The modules mirror the reference's names
(tree_data, substitution, prune, bdsky, util) for exactly that path and nothing else."""
from . import _lib  # noqa: F401  (raises if the CUDA library has not been built)
from .substitution import HKY, JC69, SubstitutionModel, encode_partials  # noqa: F401
from .tree_data import TreeData  # noqa: F401
from .util import TipData  # noqa: F401

__all__ = ["HKY", "JC69", "SubstitutionModel", "TreeData", "TipData", "encode_partials"]
