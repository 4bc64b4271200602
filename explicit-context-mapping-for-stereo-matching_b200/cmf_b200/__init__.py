"""B200-native kernels of the cmfsm hot path: ctypes binding (`lib`), tensor wrappers (`ops`),
differentiable wrappers (`autograd_ops`).  The shared library is built in-tree as libcmfb200.so."""
from . import lib  # noqa: F401
