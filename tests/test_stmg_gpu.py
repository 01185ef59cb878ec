"""Parity of the space-time multigrid pieces (transfers, Vanka, smoother, V-cycle) and of the
preconditioned FGMRES solve with the CPU oracle (oracle/stmg.py), through the C ABI.

Float multigrid levels: 1e-5 relative (north_star) for operator/transfers; the Vanka inverse goes through
an O(cond) amplification, its tolerance is stated per test.  FGMRES iteration counts: +-1."""
import numpy as np
import pytest

from golden_util import load
from gpu_util import gpu_levels, gpu_multigrid, rand_block, rel
from oracle import fe_time as ft
from oracle import stmg, tp_01

pytestmark = pytest.mark.gpu
G = load("tp_01")


def _tau(p, ref):
    return p["endTime"] * 2.0 ** (-(ref + 1))


def _params(name, dim=2, **over):
    p = tp_01.parse_parameters(G["params"].get(name, {}) if name else {}, dim)
    p.update(over)
    return p


CONFIGS = [
    ("tf03", 2, 3, 0, {}),                      # heat DG(1) x Q2, levels h h h k
    ("tf03", 2, 2, 1, {}),                      # DG(2) x Q3
    ("tf01", 2, 3, 0, {}),                      # 2 steps at once, space_and_time, p-multigrid
    ("tf04", 2, 2, 0, {}),                      # CGP(2) x Q3
    ("tf05", 2, 2, 0, {}),                      # wave, 4 steps at once (tau levels)
    ("tf03", 2, 2, 0, {"distortGrid": 0.15}),   # perturbed mesh: one patch matrix per cell
    ("tf03", 3, 1, 0, {"mgTimeBeforeSpace": True}),   # 3D
]


@pytest.mark.parametrize("name,dim,ref,deg_idx,over", CONFIGS)
def test_multigrid_pieces_match_oracle(ctx, name, dim, ref, deg_idx, over):
    import dealii_stfem_b200 as st
    p = _params(name, dim, **over)
    k = p["feDegree"] + deg_idx
    lv = tp_01.build_levels(p, dim, ref, k, _tau(p, ref), np.float32)
    omg = stmg.GMG(p["timeType"], lv["ops"], lv["spaces"], lv["mg_type_level"], lv["poly_time"], p["nTimestepsAtOnce"],
                   lv["ptypes"], np.float32)
    meshes, ops = gpu_levels(st, ctx, lv, st.F32)
    mg = gpu_multigrid(st, ctx, p, lv, ops)
    nl = mg.n_levels
    assert nl == len(lv["ops"])
    for l in range(nl):
        oop, space = lv["ops"][l], lv["spaces"][l]
        nb, n = oop.nb, oop.n
        x = rand_block(nb, n, 10 + l, space.constrained, np.float32)
        dx, dy = ops[l].new_vector().upload(x), ops[l].new_vector()
        # level operator
        mg.level_apply(l, 4, dy, dx)
        assert rel(dy.download(), oop.vmult(x)) < 1e-5
        info = mg.level_info(l)
        assert int(info["smoother"]) == lv["ptypes"][l]
        if lv["ptypes"][l] != 0:
            # Vanka: float inverse of a patch matrix; compare against the oracle's (double inverse cast to float)
            mg.level_apply(l, 0, dy, dx)
            assert rel(dy.download(), omg.vanka[l].vmult(x)) < 2e-4
            lam_o = omg.smoothers[l].lambda_estimate
            assert abs(info["lambda"] - lam_o) < 2e-3 * abs(lam_o)
            mg.level_apply(l, 1, dy, dx)
            assert rel(dy.download(), omg.smoothers[l].vmult(x)) < 5e-4
        if l > 0:
            oc = lv["ops"][l - 1]
            dc = ops[l - 1].new_vector()
            mg.level_apply(l, 2, dc, dx)                      # restrict
            r_o = np.zeros((oc.nb, oc.n), np.float32)
            omg.transfers[l].restrict_and_add(r_o, x)
            assert rel(dc.download(), r_o) < 1e-5
            xc = rand_block(oc.nb, oc.n, 30 + l, lv["spaces"][l - 1].constrained, np.float32)
            dc.upload(xc)
            mg.level_apply(l, 3, dy, dc)                      # prolongate
            p_o = np.zeros((nb, n), np.float32)
            omg.transfers[l].prolongate_and_add(p_o, xc)
            assert rel(dy.download(), p_o) < 1e-5
            # adjointness <P u, v> = <u, R v> (restrict_is_transpose_prolongate)
            lhs = np.vdot(dy.download().astype(np.float64), x.astype(np.float64))
            mg.level_apply(l, 2, dc, dx)
            rhs = np.vdot(xc.astype(np.float64), dc.download().astype(np.float64))
            assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), 1.0)
            dc.free()
        dx.free(); dy.free()
    # the whole V-cycle, double in / double out
    top = lv["ops"][-1]
    b = rand_block(top.nb, top.n, 77, lv["spaces"][-1].constrained)
    db = st.DeviceBlockVector(ctx, top.nb, top.n, st.F64).upload(b)
    dz = st.DeviceBlockVector(ctx, top.nb, top.n, st.F64)
    mg.vmult(dz, db)
    z_o = omg.vmult(b)
    assert rel(dz.download(), z_o) < 2e-3
    db.free(); dz.free(); mg.close()
    for o in ops:
        o.close()
    for m in meshes:
        m.close()


@pytest.mark.parametrize("name,dim,ref,over", [("tf03", 2, 3, {}), ("tf04", 2, 2, {"smoother": "chebyshev", "smoothingSteps": 2}),
                                               ("tf03", 3, 1, {"mgTimeBeforeSpace": True})])
def test_coarse_grid_gmres_matches_oracle(ctx, name, dim, ref, over):
    """coarseGridSmootherType != "Smoother" (include/stmg.h:1240-1302): left-preconditioned GMRES(10) with
    IterationNumberControl(10, 1e-20) on the coarsest level instead of the smoother.  The V-cycle (eager on the first call,
    then replayed as two CUDA graphs around the coarse solve) must reproduce the oracle's."""
    import dealii_stfem_b200 as st
    p = _params(name, dim, **over)
    k = p["feDegree"]
    lv = tp_01.build_levels(p, dim, ref, k, _tau(p, ref), np.float32)
    omg = stmg.GMG(p["timeType"], lv["ops"], lv["spaces"], lv["mg_type_level"], lv["poly_time"], p["nTimestepsAtOnce"],
                   lv["ptypes"], np.float32, smoothing_steps=p["smoothingSteps"], relaxation=p["relaxation"],
                   smoothing_range=p["smoothingRange"], eig_n_iterations=p["smoothingEigCgNIterations"], variable=p["variable"],
                   coarse_gmres=(10, 1e-20))
    meshes, ops = gpu_levels(st, ctx, lv, st.F32)
    mg = st.Multigrid(ctx, ops, lv["mg_type_level"], lv["ptypes"], p["timeType"], p["nTimestepsAtOnce"], lv["poly_time"],
                      smoothing_steps=p["smoothingSteps"], relaxation=p["relaxation"], smoothing_range=p["smoothingRange"],
                      eig_n_iterations=p["smoothingEigCgNIterations"], variable=p["variable"],
                      restrict_is_transpose_prolongate=p["restrictIsTransposeProlongate"], coarse_grid_maxiter=10, coarse_grid_abstol=1e-20)
    top = lv["ops"][-1]
    db = st.DeviceBlockVector(ctx, top.nb, top.n, st.F64)
    dz = st.DeviceBlockVector(ctx, top.nb, top.n, st.F64)
    for seed in (77, 78, 79):                      # call 1 eager, call 2 captures the two graphs, call 3 replays them
        b = rand_block(top.nb, top.n, seed, lv["spaces"][-1].constrained)
        mg.vmult(dz, db.upload(b))
        z_o = omg.vmult(b)
        assert rel(dz.download(), z_o) < 2e-3
    db.free(); dz.free(); mg.close()
    for o in ops:
        o.close()
    for m in meshes:
        m.close()


@pytest.mark.parametrize("name,dim,ref,over", [("tf03", 2, 3, {}), ("tf03", 2, 3, {"mgTimeBeforeSpace": True}),
                                               ("tf04", 2, 3, {}), ("tf01", 2, 3, {}), ("tf07", 2, 3, {}),
                                               ("tf03", 2, 2, {"distortGrid": 0.1, "smoother": "chebyshev", "smoothingSteps": 2}),
                                               ("tf03", 3, 2, {"mgTimeBeforeSpace": True})])
def test_fgmres_iteration_counts_match_oracle(ctx, name, dim, ref, over):
    """One space-time solve A x = b with a random right-hand side: iteration count within +-1 of the
    oracle (float multigrid on both sides), solutions agree to the solver tolerance."""
    import dealii_stfem_b200 as st
    p = _params(name, dim, **over)
    k = p["feDegree"]
    tau = _tau(p, ref)
    lv = tp_01.build_levels(p, dim, ref, k, tau, np.float32)
    sm = {"relaxation": 1, "chebyshev": 2}[p["smoother"].lower()]
    omg = stmg.GMG(p["timeType"], lv["ops"], lv["spaces"], lv["mg_type_level"], lv["poly_time"], p["nTimestepsAtOnce"],
                   lv["ptypes"], np.float32, smoothing_steps=p["smoothingSteps"])
    space = lv["spaces"][-1]
    A, B = lv["fetw"][-1][0], lv["fetw"][-1][1]
    fine = stmg.LevelOperator(space, A, B, np.float64)
    nb = fine.nb
    b = rand_block(nb, space.n_dofs, 5, space.constrained)
    x_o, it_o, hist = stmg.fgmres(fine.vmult, np.zeros_like(b), b, omg.vmult)
    meshes, ops = gpu_levels(st, ctx, lv, st.F32)
    mg = gpu_multigrid(st, ctx, p, lv, ops)
    opA = st.Operator(meshes[-1], space.k, A, B, number_type=st.F64)
    dx = opA.new_vector(); dx.zero()
    db = opA.new_vector().upload(b)
    solver = st.Fgmres()
    it_g = solver.solve(opA, dx, db, mg)
    assert abs(it_g - it_o) <= 1, (it_g, it_o)
    assert solver.final_residual <= 1e-12 * max(solver.initial_residual, 1.0) * 1.0001 or solver.final_residual <= 1e-12
    assert rel(dx.download(), x_o) < 1e-8
    solver.close(); dx.free(); db.free(); opA.close(); mg.close()
    for o in ops:
        o.close()
    for m in meshes:
        m.close()


# Kronecker / fast-diagonalisation form of the patch smoother (csrc/vanka_fd.cuh) against the oracle's dense
# patch inverses (reference algorithm, include/stmg.h:745-872) and against the dense device path (variant 2).
FD_CASES = [
    # k, cells, upper, ttype, r, nts, dirichlet
    (2, [3, 4, 3], [1.0, 1.0, 1.0], "DG", 1, 1, 0x3f),
    (4, [3, 3, 3], [1.0, 1.5, 0.5], "CGP", 2, 1, 0x3f),      # config 2 family, anisotropic cells
    (3, [4, 2, 1], [1.0, 1.0, 1.0], "DG", 2, 1, 0x15),       # nb = 3, single cell in z, partial Dirichlet
    (1, [5, 5, 5], [1.0, 1.0, 1.0], "DG", 1, 2, 0x3f),       # nb = 4
    (3, [3, 3, 3], [1.0, 1.0, 1.0], "DG", 0, 1, 0x00),       # nb = 1, pure Neumann (mass term regularises)
]


@pytest.mark.parametrize("number_type", [0, 1])
@pytest.mark.parametrize("case", FD_CASES, ids=lambda c: "k%d_%s_%s%d_x%d_m%x" % (c[0], "x".join(map(str, c[1])), c[3], c[4], c[5], c[6]))
def test_vanka_kronecker_form_matches_dense_patches(ctx, case, number_type):
    import dealii_stfem_b200 as st
    from oracle import spatial as S
    k, cells, upper, ttype, r, nts, mask = case
    mesh = S.Mesh(3, cells, 0, lower=[0, 0, 0], upper=upper)
    space = S.Space(mesh, k, dirichlet_faces=mask)
    A, B = ft.get_fe_time_weights(ttype, r, 0.05, nts)[:2]
    dt = np.float64 if number_type == 0 else np.float32
    lop = stmg.LevelOperator(space, A, B, dt)
    vanka_o = stmg.PreconditionVanka(lop, dt)
    x = rand_block(lop.nb, lop.n, 3, space.constrained, dt)
    ref = vanka_o.vmult(x)
    gm = st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper, dirichlet_faces=mask)
    outs = []
    for variant in (0, 2):
        op = st.Operator(gm, k, A, B, number_type=number_type, variant=variant)
        mg = st.Multigrid(ctx, [op], "", [1], ttype, nts, [r])
        info = mg.level_info(0)
        assert (info["patch_matrices"] == 0) == (variant == 0)       # 0 stored patch matrices <=> Kronecker form
        dx, dy = op.new_vector().upload(x), op.new_vector()
        mg.level_apply(0, 0, dy, dx)
        outs.append(dy.download())
        tol = 1e-9 if number_type == 0 else 2e-4
        assert rel(outs[-1], ref) < tol, "variant %d" % variant
        assert np.all(outs[-1][:, space.constrained] == 0)
        dx.free(); dy.free(); mg.close(); op.close()
    gm.close()
