"""block_b200: B200-native hot path of a DMRG sweep (sigma = H.psi, Davidson, renormalisation) behind the entry
points of sanshar/Block.  Everything numerical lives in block_b200/lib/libblockb200.so (CUDA, sm_100a)."""
from . import _lib  # noqa: F401
from .hotpath import B2DError, BlockSpec, OperatorSpec, SpinBlock, block_spec_from_record, spinblock_from_record  # noqa: F401
