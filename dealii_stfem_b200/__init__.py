"""Importable alias of the hyphen-named package directory `dealii-stfem_b200/`."""
import os as _os

__path__.insert(0, _os.path.normpath(_os.path.join(_os.path.dirname(__file__), "..", "dealii-stfem_b200")))
from .capi import *  # noqa: E402,F401,F403
from . import capi, dist, driver, fe_time_host, problem_host  # noqa: E402,F401
from .driver import HeatWaveProblem, parse_parameters  # noqa: E402,F401
