"""The product's tp_01 front end (dealii-stfem_b200/tp_01.py): parameter files with the reference's JSON keys and the
printed output format.  The formatter is fed the numbers of tests/tp_01.output (tests/golden/tp_01.json) and must
reproduce the stored text (tests/golden/tp_01_text.json: three of the eight blocks) character by character: run headers, convergence tables with
log2 rates, iteration count tables.  CPU only (no solve here; the GPU run is tests/test_zz_practical_gpu.py /
tests/test_tp01_gpu.py)."""
import io

import pytest

import dealii_stfem_b200 as st
from dealii_stfem_b200 import tp_01 as front
from golden_util import load

G = load("tp_01")
T = load("tp_01_text")


@pytest.mark.parametrize("name", sorted(T))
def test_printed_output_reproduces_reference_text(name):
    out = io.StringIO()
    p = st.parse_parameters(G["params"][name], 2)
    avgs = []
    for tab in G["tables"][name]:
        rows = [dict(run, levels="x" * run["max_level"]) for run in tab["runs"]]
        for row in rows:
            out.write(front.run_header(row))
        out.write("Convergence table k=%d\n" % tab["k"])
        out.write(front.convergence_table(rows, legacy_work=True))
        out.write("\n")
        avgs.append([r["iterations"] / r["timesteps"] for r in rows])
    degrees = list(range(p["feDegree"], p["feDegree"] + p["nDegCycles"]))
    refinements = list(range(p["refinement"], p["refinement"] + p["nRefCycles"]))
    assert degrees == [t["k"] for t in G["tables"][name]]
    out.write("Iteration count table\n")
    out.write(front.iteration_table(degrees, refinements, avgs))
    out.write("\n")
    got = out.getvalue().split("\n")[:-1]
    assert got == T[name]


def test_work_column_follows_current_source():
    """tp_01.cc:715: work = n_dofs * n_blocks * total_gmres_iterations."""
    row = dict(cells=16, s_dofs=81, t_dofs=4, timesteps=4, iterations=28, linf=1.0, l2=1.0, h1=1.0)
    txt = front.convergence_table([row])
    assert txt.splitlines()[1].split()[:5] == ["16", "81", "4", "1296", "9072"]
    nan = front.convergence_table([row, row], st_convergence=False).splitlines()
    assert nan[1].split()[5:] == ["nan", "-"] * 3 and nan[2].split()[5:] == ["nan", "nan"] * 3


def test_parameter_file_keys():
    """Every key the reference's parse() registers (include/parameters.h:92-144) that tp_01.cc reads is understood;
    sourcePoint defaults to the midpoint of the default box (parameters.h:79)."""
    p = st.parse_parameters({"hyperRectLowerLeft": "-1.0,-1.0,-1.0", "hyperRectUpperRight": "1.0,1.0,1.0", "subdivisions": "5,5,5",
                             "distortCoeff": "0.5", "spaceTimeConvergenceTest": "false", "functionalFile": "practical_01.txt"}, 3)
    assert p["sourcePoint"] == [0.5, 0.5, 0.5] and p["subdivisions"] == [5, 5, 5] and p["distortCoeff"] == 0.5
    assert p["spaceTimeConvergenceTest"] is False and p["functionalFile"] == "practical_01.txt"
    p = st.parse_parameters({"sourcePoint": "0.0,0.0,0.0", "timeType": "DG", "feDegree": "0"}, 3)
    assert p["sourcePoint"] == [0.0, 0.0, 0.0] and p["feDegreeMin"] == 0


def test_functional_file_line_format():
    """tp_01.cc:620-630: time in a 16-wide scientific field, every value after a 16-wide blank field."""
    from dealii_stfem_b200.driver import format_functional_row
    ln = format_functional_row(0.0625, [2.229576e-02, -3.5])
    assert ln == "    6.250000e-02                2.229576e-02                -3.500000e+00\n"
