"""Oracle AND product host algebra against the reference's own known-answer outputs
(tests/tp_02.output, tests/transfer_02.output, tests/tp04.cc -> tests/golden/*.json) and against
each other at full double precision.  CPU only."""
import re

import numpy as np
import pytest

from golden_util import load, matches_print
from oracle import fe_time as oft

import dealii_stfem_b200.fe_time_host as pft

IMPLS = {"oracle": oft, "product": pft}


def _type(tag):
    return "CGP" if tag == "CG" else "DG"


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_tp_02_time_weights(impl):
    ft = IMPLS[impl]
    d = load("tp_02")
    checked = 0
    i = 0
    while i < len(d):
        h, ms = d[i]["header"], d[i]["matrices"]
        m = re.match(r"^(CG|DG)\((\d+)\)$", h)
        if m:
            t, r = _type(m.group(1)), int(m.group(2))
            # single-step weights with tau = 1: split_lhs_rhs of get_cg_weights / get_dg_weights
            A, B, G, Z = ft.get_fe_time_weights(t, r, 1.0, 1)
            if t == "CGP":
                # printed: matrix (r x r+1) = [-Gamma | Alpha], matrix_der = [-Zeta | Beta]
                mine = [np.hstack([-G, A]), np.hstack([-Z, B])]
                wave_in = (A, B, G, Z)
            else:
                # printed: jump, mass, derivative; get_fe_time_weights returns the jump as Gamma (fe_time.h:403-407)
                mine = [G, A, B]
                wave_in = (A, B, G, None)     # tp_02.cc passes (mass, der, jump, nil)
            for a, g in zip(mine, ms):
                assert matches_print(a, g), (impl, h)
                checked += 1
            wave = ft.get_fe_time_weights_wave(t, *wave_in)
            for a, g in zip(wave, d[i + 1]["matrices"]):
                assert matches_print(a, g), (impl, h, "wave")
                checked += 1
            i += 2
            continue
        m = re.match(r"^(CG|DG)\((\d+)\) - (\d+) timesteps in one system$", h)
        if m:
            t, r, n = _type(m.group(1)), int(m.group(2)), int(m.group(3))
            for a, g in zip(ft.get_fe_time_weights(t, r, 1.0, n), ms):
                assert matches_print(a, g), (impl, h)
                checked += 1
            w1 = ft.get_fe_time_weights(t, r, 1.0, 1)
            for a, g in zip(ft.get_fe_time_weights_wave(t, *w1, n), d[i + 1]["matrices"]):
                assert matches_print(a, g), (impl, h, "wave")
                checked += 1
            i += 2
            continue
        i += 1
    assert checked == 183


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_transfer_02_time_transfer_matrices(impl):
    ft = IMPLS[impl]
    d = load("transfer_02")
    seen, checked, i = {}, 0, 0
    while i < len(d):
        h = d[i]["header"]
        if h in ("- Prolongation", "- Restriction"):
            h2 = d[i + 1]["header"]
            m = re.match(r"^(CG|DG)\((\d+)\)$", h2)
            t, r = _type(m.group(1)), int(m.group(2))
            seen[(h, h2)] = seen.get((h, h2), 0) + 1
            n = 2 if seen[(h, h2)] == 1 else 4      # transfer_02.cc: second pass uses 4 steps at once
            M = ft.get_time_prolongation_matrix(t, r, n) if "Prol" in h else ft.get_time_restriction_matrix(t, r, n)
            assert matches_print(M, d[i + 1]["matrices"][0]), (impl, h, h2, n)
            checked += 1
            i += 2
            continue
        if h == "- Projection":
            m = re.match(r"^(CG|DG) From (\d+) to (\d+)$", d[i + 1]["header"])
            n = int(re.match(r"Timesteps at once: (\d+)", d[i + 2]["header"]).group(1))
            M = ft.get_time_projection_matrix(_type(m.group(1)), int(m.group(2)), int(m.group(3)), n)
            assert matches_print(M, d[i + 2]["matrices"][0]), (impl, d[i + 1]["header"], n)
            checked += 1
            i += 3
            continue
        i += 1
    assert checked == 60


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_tp04_mg_sequences(impl):
    """Exact integer pins.  Cases that pass an EMPTY p_seq together with use_p_multigrid_space=true
    (tp04.cc:244-871) predate today's signature (p_seq.size()-1 would underflow, fe_time.cc:79): their
    expectations are reproduced by letting the space degrees follow k_seq, which is what the driver
    does (k = r+1, tp_01.cc:77)."""
    ft = IMPLS[impl]
    t = load("tp04")
    assert t["output_fail_lines"] == 0
    names = {"tau": "t", "k": "k", "h": "h", "p": "p"}
    for c in t["cases"]:
        p_seq = c["p_seq"]
        if c["use_p_multigrid_space"] and not p_seq:
            p_seq = c["k_seq"]
        r = ft.get_mg_sequence(c["n_sp_lvl"], c["k_seq"], p_seq, c["n_timesteps_at_once"], c["n_timesteps_at_once_min"],
                               c["lower_lvl"], c["coarsening_type"], c["time_before_space"],
                               c["use_p_multigrid_space"], c["zip_from_back"])
        assert r == [names[x] for x in c["expected"]], c
        if "expected_p" in c:
            p = ft.get_precondition_stmg_types(r, c["coarsening_type"], c["time_before_space"], c["p_zip_from_back"])
            assert p == c["expected_p"], c


def test_appendix_b_values():
    """SURVEY.md App. B full-precision values (regenerated independently in the survey)."""
    A, B, G, Z = oft.get_fe_time_weights("DG", 1, 1.0, 1)
    assert np.allclose(A, np.diag([0.75, 0.25]), atol=1e-14)
    assert np.allclose(B, [[1.125, 0.375], [-1.125, 0.625]], atol=1e-14)
    assert np.allclose(G[:, 0], [1.5, -0.5], atol=1e-14)
    A, B, G, Z = oft.get_fe_time_weights("CGP", 2, 1.0, 1)
    assert np.allclose(A, [[2 / 3, 0], [0, 1 / 6]], atol=1e-14)
    assert np.allclose(B, [[4 / 3, 1 / 3], [-4 / 3, 2 / 3]], atol=1e-14)
    assert np.allclose(G[:, 0], [-1 / 3, 1 / 6], atol=1e-14)
    assert np.allclose(Z[:, 0], [5 / 3, -2 / 3], atol=1e-14)
    A, B, G, Z = oft.get_fe_time_weights("DG", 2, 1.0, 1)
    assert np.allclose(np.diag(A), [0.376403062700467, 0.512485826188422, 0.111111111111111], atol=1e-13)
    assert np.allclose(G[:, 0], [1.5580782047249222, -0.8914115380582552, 0.3333333333333331], atol=1e-13)


@pytest.mark.parametrize("ttype,r,nts", [("DG", 0, 1), ("DG", 1, 2), ("DG", 3, 4), ("CGP", 1, 1), ("CGP", 2, 2),
                                         ("CGP", 4, 4), ("DG", 5, 1), ("CGP", 5, 2)])
def test_product_equals_oracle_full_precision(ttype, r, nts):
    tau = 0.0375
    for a, b in zip(oft.get_fe_time_weights(ttype, r, tau, nts), pft.get_fe_time_weights(ttype, r, tau, nts)):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-13)
    w1o = oft.get_fe_time_weights(ttype, r, tau, 1)
    w1p = pft.get_fe_time_weights(ttype, r, tau, 1)
    for a, b in zip(oft.get_fe_time_weights_wave(ttype, *w1o, nts), pft.get_fe_time_weights_wave(ttype, *w1p, nts)):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-10 * max(1.0, np.abs(a).max()))
    if nts >= 2:
        assert np.allclose(oft.get_time_prolongation_matrix(ttype, r, nts), pft.get_time_prolongation_matrix(ttype, r, nts), atol=1e-12)
        assert np.allclose(oft.get_time_restriction_matrix(ttype, r, nts), pft.get_time_restriction_matrix(ttype, r, nts), atol=1e-12)
    lo = 0 if ttype == "DG" else 1
    if r > lo:
        assert np.allclose(oft.get_time_projection_matrix(ttype, r - 1, r, nts), pft.get_time_projection_matrix(ttype, r - 1, r, nts), atol=1e-12)
        assert np.allclose(oft.get_time_projection_matrix(ttype, r, r - 1, nts), pft.get_time_projection_matrix(ttype, r, r - 1, nts), atol=1e-12)


@pytest.mark.parametrize("n", range(1, 9))
def test_quadrature_rules(n):
    from oracle import quadrature as Q
    for kind, fn in (("gauss", Q.gauss), ("lobatto", Q.gauss_lobatto), ("radau", Q.gauss_radau_right)):
        if kind == "lobatto" and n < 2:
            continue
        xo, wo = fn(n)
        xp, wp = pft.quadrature_rule(kind, n)
        assert np.allclose(xo, xp, atol=1e-14) and np.allclose(wo, wp, atol=1e-14)
        assert abs(wo.sum() - 1.0) < 1e-14
        # exactness: Gauss 2n-1, Radau 2n-2, Lobatto 2n-3
        deg = {"gauss": 2 * n - 1, "radau": 2 * n - 2, "lobatto": 2 * n - 3}[kind]
        for p in range(deg + 1):
            assert abs((wp * xp ** p).sum() - 1.0 / (p + 1)) < 1e-13
