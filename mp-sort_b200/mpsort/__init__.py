"""mpsort (B200): drop-in for the `mpsort` Python package of MP-sort, backed by
libmpsort-b200.so (hand-written sm_100a CUDA + NCCL)."""
from .version import __version__  # noqa: F401
