"""tests/cpp/tp01_main.cpp = the reference's tp_01 driver (tests/tp_01.cc:56-848) as a C++ host program on the C ABI: it must
compile as plain C++17 against include/stfem_b200.h (CPU check) and, on the GPU, print the reference's convergence tables:
error norms of tests/tp_01.output to its 6 printed digits, iteration averages within one iteration per solve."""
import json
import os
import re
import subprocess

import pytest

from golden_util import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "tp01_main.cpp")
LIBDIR = os.path.join(ROOT, "dealii-stfem_b200")
G = load("tp_01")


def _compile(out):
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", out, "-L", LIBDIR, "-lstfem_b200",
           "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_tp01_main_compiles_and_links(tmp_path):
    import dealii_stfem_b200 as st
    st.capi.lib()                      # the library must exist (no fallback)
    exe = _compile(str(tmp_path / "tp01_main"))
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tf03", "tf07"])
def test_tp01_main_prints_the_reference_tables(tmp_path, name):
    exe = _compile(str(tmp_path / "tp01_main"))
    pj = dict(G["params"][name])
    pj["nRefCycles"] = "2"                      # the first two refinements of the stored output
    pj["nDegCycles"] = "1"
    pfile = tmp_path / (name + ".json")
    pfile.write_text(json.dumps(pj))
    r = subprocess.run([exe, "--file", str(pfile), "--dim", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    gold = G["tables"][name][0]["runs"][:2]
    # per-run headers (tp_01.cc:101-104, 703-707)
    heads = re.findall(r":: Number of active cells: (\d+)\n:: Number of degrees of freedom: (\d+)\n:: Min Level 0  Max Level (\d+)\n"
                       r"Average GMRES iterations (\S+) \((\d+) gmres_iterations / (\d+) timesteps\)", out)
    assert len(heads) == 2, out
    for h, g in zip(heads, gold):
        assert int(h[0]) == g["cells"] and int(h[1]) == g["s_dofs"] and int(h[5]) == g["timesteps"]
    # the convergence table: cells s-dofs t-dofs st-dofs work | Linf rate | L2 rate | H1 rate
    m = re.search(r"Convergence table k=(\d+)\n(.*?)\n\n", out, re.S)
    assert m, out
    lines = m.group(2).split("\n")
    assert "L2-L2" in lines[0] and "L2-H1_semi" in lines[0] and "st-dofs" in lines[0]
    for ln, g in zip(lines[1:], gold):
        f = ln.split()
        assert int(f[0]) == g["cells"] and int(f[1]) == g["s_dofs"] and int(f[2]) == g["t_dofs"]
        linf, l2, h1 = float(f[5]), float(f[7]), float(f[9])
        for mine, ref in ((linf, g["linf"]), (l2, g["l2"]), (h1, g["h1"])):
            assert abs(mine - ref) <= 1.5e-5 * abs(ref), (ln, g)           # 6 printed digits on either side
    assert lines[1].split()[6] == "-" and re.match(r"^\d+\.\d\d$", lines[2].split()[6])     # convergence rates (log2)
    assert "Iteration count table" in out
