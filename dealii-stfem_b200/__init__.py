"""dealii-stfem_b200: B200-native hot path of immaaane/dealii-stfem.

csrc/        CUDA kernels (sm_100a) + the C ABI declared in include/stfem_b200.h
include/     C++ façade mirroring the reference's operator interface (operators.h / stmg.h)
capi.py      ctypes binding used by tests/ and bench.py
driver.py    tp_01-style driver (parameters, level hierarchy, time loop) over the C ABI
Import as `dealii_stfem_b200` (the hyphenated directory is re-exported by that shim).
"""
from .capi import *  # noqa: F401,F403
from . import capi, dist, driver, fe_time_host  # noqa: F401
from .driver import HeatWaveProblem, parse_parameters  # noqa: F401
