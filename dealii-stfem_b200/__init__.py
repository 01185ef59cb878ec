"""dealii-stfem_b200: B200-native hot path of immaaane/dealii-stfem.

csrc/        CUDA kernels (sm_100a) + the C ABI declared in include/stfem_b200.h
include/     C++ façade mirroring the reference's operator interface (operators.h / stmg.h)
capi.py      ctypes binding used by tests/ and bench.py
driver.py    tp_01-style driver (parameters, level hierarchy, time loop) over the C ABI
tp_01.py     command-line front end: the reference's JSON parameter files in, tp_01's printed tables out
problem_host.py  coefficient tables / cut-off initial value of the practical runs (host-side C-ABI helpers)
Import as `dealii_stfem_b200` (the hyphenated directory is re-exported by that shim).
"""
from .capi import *  # noqa: F401,F403
from . import capi, dist, driver, fe_time_host, problem_host  # noqa: F401
from .driver import HeatWaveProblem, parse_parameters  # noqa: F401
