"""Shared helpers of the GPU parity tests: mirror an oracle level hierarchy on the device."""
import numpy as np


def gpu_mesh(st, ctx, mesh):
    return st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper,
                   vertices=None if mesh.cartesian else mesh.vertices.reshape(-1, mesh.dim))


def gpu_levels(st, ctx, lv, number_type, coeff_q=None):
    """lv: dict from oracle.tp_01.build_levels.  Returns (meshes, ops) coarse -> fine."""
    meshes, ops = [], []
    for l, space in enumerate(lv["spaces"]):
        gm = gpu_mesh(st, ctx, space.mesh)
        A, B = lv["fetw"][l][0], lv["fetw"][l][1]
        ops.append(st.Operator(gm, space.k, A, B, number_type=number_type))
        meshes.append(gm)
    return meshes, ops


def gpu_multigrid(st, ctx, p, lv, ops):
    return st.Multigrid(ctx, ops, lv["mg_type_level"], lv["ptypes"], p["timeType"], p["nTimestepsAtOnce"], lv["poly_time"],
                        smoothing_steps=p["smoothingSteps"], relaxation=p["relaxation"], smoothing_range=p["smoothingRange"],
                        eig_n_iterations=p["smoothingEigCgNIterations"], variable=p["variable"],
                        restrict_is_transpose_prolongate=p["restrictIsTransposeProlongate"])


def rel(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-300)


def rand_block(nb, n, seed=42, mask=None, dtype=np.float64):
    v = np.stack([np.random.RandomState(seed + b).uniform(-1, 1, n) for b in range(nb)])
    if mask is not None:
        v[:, mask] = 0
    return v.astype(dtype)
