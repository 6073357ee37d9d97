"""B200-native OS-CNN + feature-level style-transfer training path.

Host side: the reference's own module interface (``OS_CNN.OS_CNN``, ``OS_CNN.OS_CNN_Structure_build``) plus the
two style-transfer operators; compute side: hand-written sm_100a kernels in ``libtsc_b200.so`` (C-ABI in
``include/tsc_b200.h``).  No CPU fallback, no alternate backend.
"""
from . import _lib, ops                                       # noqa: F401
from .functional import adain, gram_style_loss, os_stack      # noqa: F401
from .ops import set_engine, get_engine                       # noqa: F401

__all__ = ["adain", "gram_style_loss", "os_stack", "set_engine", "get_engine", "ops"]
