"""Inputs shared by tests/gpu_shot.py (GPU half) and tests/check_shot.py (oracle half); the same cases as
tests/test_zz_practical_gpu.py.  Test infrastructure: may use the oracle (mesh generator only)."""
import numpy as np

from golden_util import load
from oracle import spatial as S

PRACTICAL = {"spaceTimeMg": "true", "mgTimeBeforeSpace": "false", "timeType": "DG", "nTimestepsAtOnce": "2", "feDegree": "1",
             "extrapolate": "false", "spaceTimeConvergenceTest": "false", "hyperRectLowerLeft": "-1.0,-1.0,-1.0",
             "hyperRectUpperRight": "1.0,1.0,1.0", "subdivisions": "5,5,5", "distortCoeff": "0.5", "sourcePoint": "0.0,0.0,0.0"}
DIAG_CASES = [(3, 4, 0.0, "CGP", 2, None), (3, 3, 0.15, "DG", 2, "q"), (3, 2, 0.0, "DG", 1, "cell"), (2, 2, 0.1, "DG", 1, None),
              (2, 5, 0.0, "CGP", 3, None)]
POINT_CASES = [(3, 0.15, 3), (3, 0.0, 4), (2, 0.1, 2)]
JACOBI_CASES = [("tf03", 2, 3, {}), ("tf04", 2, 3, {"smoother": "chebyshev", "smoothingSteps": "3"}), ("tf03", 3, 2, {})]


def tp01_params(name):
    return dict(load("tp_01")["params"][name])


def box_mesh_vertices(dim, sub, ref, lo, up, distort):
    """vertices [(z,) y, x, dim] of the oracle's perturbed mesh, None for a Cartesian one."""
    if distort == 0.0:
        return None
    return S.Mesh(dim, sub, ref, lo, up, distort=distort).vertices


def practical_2d_case():
    pj = dict(PRACTICAL, problemType="heat", hyperRectLowerLeft="-1.0,-1.0", hyperRectUpperRight="1.0,1.0", subdivisions="5,5",
              sourcePoint="0.0,0.0", distortGrid="0.1", nTimestepsAtOnce="1")
    mesh = S.Mesh(2, [5, 5], 2, [-1.0, -1.0], [1.0, 1.0], distort=0.1)
    V = mesh.vertices.reshape(-1, 2)
    c = V[np.argmin(np.sum(V * V, axis=1))]
    pj["sourcePoint"] = "%.17g,%.17g" % (c[0], c[1])
    return pj, mesh.vertices
