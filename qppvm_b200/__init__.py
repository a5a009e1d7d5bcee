"""qppvm_b200 — B200-native batched whole-body QP solve behind the qppvm plugin surface.

Only the hot path of the reference lives here: ``csrc/`` holds the sm_100a
kernels and the C-ABI, ``plugin/`` the C++ mirror of the XBotCore plugin classes, and the Python
modules are the host-side plumbing (layout, synthetic states, ctypes front end, sharding).
"""
from .layout import Desc, Layout, CONFIGS  # noqa: F401
