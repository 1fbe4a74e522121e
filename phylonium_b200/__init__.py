"""phylonium_b200 — B200-native distance pipeline with phylonium's process() interface.

Python side: a ctypes binding of the C ABI (capi.py) and a thin mirror of the reference's
host interface (pipeline.py).  All computation happens in libphylonium_b200.so (CUDA,
sm_100a); nothing here computes on the CPU.
"""
from .capi import (  # noqa: F401
    DIST_ANI,
    DIST_JC,
    DIST_RAW,
    HOM_DTYPE,
    PHYLO_FLAG_COMPLETE_DELETION,
    Context,
    PhyloError,
    gc_content,
    load_library,
    min_anchor_length,
    threshold_for,
)
from .pipeline import EvoModel, format_matrix, process  # noqa: F401

__version__ = "0.1"
