#!/usr/bin/env python3
"""Extract the reference's own known-answer data into small JSON fixtures.

Run in the BUILD container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

Inputs (read-only, never copied verbatim):
  tests/tp_02.output        time weights, 2-decimal pins      -> tp_02.json
  tests/transfer_02.output  time transfer matrices, 2 decimals -> transfer_02.json
  tests/tp04.cc             get_mg_sequence / smoother-type expectations
                            (exact integer pins, stated in the test source)  -> tp04.json
  tests/tp_01.output        cells / dofs / iterations / error tables -> tp_01.json
  tests/transfer_01.output  error tables of the (stale) time-multigrid test: 2x2 cells, tau = 2^-(i+1),
                            i = 2..4, DG(j) / CGP(j+1), j = 1..3, 1 and 2 time steps at once
                            (tests/transfer_01.cc:395-396, 429, 731-737, 772-776) -> transfer_01.json

A printed matrix entry is '%7.2f', blank when |x| < 0.01
(reference tests/tp_02.cc:19-27); blank entries are stored as null.
"""
import json
import os
import re
import sys

REF = os.environ.get("STFEM_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def parse_matrix_rows(lines):
    rows = []
    for ln in lines:
        ln = ln.rstrip("\n")
        n = (len(ln) + 6) // 7
        row = []
        for j in range(n):
            tok = ln[7 * j:7 * j + 7].strip()
            row.append(float(tok) if tok else None)
        rows.append(row)
    # pad ragged rows (trailing blanks may be stripped by editors; they are not here)
    w = max(len(r) for r in rows)
    for r in rows:
        r.extend([None] * (w - len(r)))
    return rows


_NUMLINE = re.compile(r"^[ \-0-9.]*$")


def parse_sections(path):
    """Split a tp_02 / transfer_02 style output into (header, [matrices])."""
    sections = []
    cur = None
    block = []
    with open(path) as f:
        for ln in f:
            s = ln.rstrip("\n")
            if len(s) == 0:
                # empty line terminates a matrix
                if block:
                    cur["matrices"].append(parse_matrix_rows(block))
                    block = []
                continue
            if s.strip() == "" and len(s) % 7 == 0:
                # a row whose entries are all |x| < 0.01 prints as blanks only
                block.append(s)
                continue
            if _NUMLINE.match(s) and re.search(r"\d", s):
                block.append(s)
            else:
                if block:
                    cur["matrices"].append(parse_matrix_rows(block))
                    block = []
                cur = {"header": s.strip(), "matrices": []}
                sections.append(cur)
    if block:
        cur["matrices"].append(parse_matrix_rows(block))
    return sections


def golden_tp_02():
    secs = parse_sections(os.path.join(REF, "tests/tp_02.output"))
    return secs


def golden_transfer_02():
    return parse_sections(os.path.join(REF, "tests/transfer_02.output"))


def golden_tp04():
    src = open(os.path.join(REF, "tests/tp04.cc")).read()
    body = src[src.index("run_tests()"):src.index("run_idx_tests(bool")]
    # split into test blocks: top-level "{ ... }" after "// Test"
    blocks = re.split(r"\n  // Test[^\n]*\n", body)[1:]
    cases = []
    for b in blocks:
        def grab(name, default=None):
            m = re.search(name + r"\s*=\s*([^;]+);", b)
            return m.group(1).strip() if m else default

        def vec(s):
            return [int(x) for x in re.findall(r"\d+", s)] if s else []

        def mgvec(s):
            return re.findall(r"MGType::(\w+)", s)

        call = re.search(r"get_mg_sequence\((.*?)\);", b, re.S).group(1)
        args = [a.strip() for a in re.sub(r"\s+", " ", call).split(",")]
        # re-join the std::vector<unsigned int>{...} initialiser if it was split
        joined = []
        depth = 0
        for a in args:
            if depth > 0:
                joined[-1] += "," + a
            else:
                joined.append(a)
            depth += a.count("{") - a.count("}")
        args = joined
        names = {}
        for key in ("n_sp_lvl", "n_timesteps_at_once", "n_timesteps_at_once_min"):
            names[key] = int(grab(r"unsigned int\s+" + key))
        k_seq = vec(grab(r"k_seq"))
        p_seq_decl = grab(r"p_seq")
        p_seq = vec(p_seq_decl) if p_seq_decl else []
        if len(args) > 2 and "{" in args[2]:
            p_seq = vec(args[2])
        lower = re.search(r"lower_lvl\s*=\s*MGType::(\w+)", b).group(1)
        ctype = re.search(r"coarsening_type\s*=\s*CoarseningType::(\w+)", b).group(1)
        tbs = grab(r"bool\s+time_before_space") == "true"
        extra = [a for a in args[8:]]
        use_pmg = (extra[0] == "true") if len(extra) > 0 else False
        zip_back = (extra[1] == "true") if len(extra) > 1 else True
        expected = mgvec(re.search(r"expected_mg_type_level\s*=\s*\{(.*?)\};", b, re.S).group(1))
        case = dict(names, k_seq=k_seq, p_seq=p_seq, lower_lvl=lower, coarsening_type=ctype,
                    time_before_space=tbs, use_p_multigrid_space=use_pmg, zip_from_back=zip_back,
                    expected=expected)
        mp = re.search(r"get_precondition_stmg_types\((.*?)\);", b, re.S)
        if mp:
            pargs = [a.strip() for a in re.sub(r"\s+", " ", mp.group(1)).split(",")]
            case["p_zip_from_back"] = pargs[3] == "true"
            case["expected_p"] = vec(re.search(r"expected_p\s*=\s*\{(.*?)\};", b, re.S).group(1))
        cases.append(case)
    out_lines = open(os.path.join(REF, "tests/tp04.output")).read().splitlines()
    n_pass = sum(1 for l in out_lines if l.startswith("[PASS]"))
    n_fail = sum(1 for l in out_lines if "[FAIL]" in l)
    return {"cases": cases, "output_pass_lines": n_pass, "output_fail_lines": n_fail}


def golden_tp_01():
    """Per run: cells, dofs, max level, iterations, timesteps and the three error norms."""
    txt = open(os.path.join(REF, "tests/tp_01.output")).read().splitlines()
    # The 8 json configs are run in order tf01..tf08; each prints 3 degrees x 4 refinements.
    runs = []
    tables = []
    i = 0
    cur_runs = []
    while i < len(txt):
        ln = txt[i]
        m = re.match(r":: Number of active cells: (\d+)", ln)
        if m:
            cells = int(m.group(1))
            dofs = int(re.match(r":: Number of degrees of freedom: (\d+)", txt[i + 1]).group(1))
            lv = re.match(r":: Min Level (\d+)\s+Max Level (\d+)", txt[i + 2])
            it = re.match(r"Average GMRES iterations ([\d.]+) \((\d+) gmres_iterations / (\d+) timesteps\)",
                          txt[i + 3])
            cur_runs.append(dict(cells=cells, s_dofs=dofs, max_level=int(lv.group(2)),
                                 iterations=int(it.group(2)), timesteps=int(it.group(3))))
            i += 4
            continue
        m = re.match(r"Convergence table k=(\d+)", ln)
        if m:
            k = int(m.group(1))
            rows = []
            j = i + 2
            while j < len(txt) and txt[j].strip():
                t = txt[j].split()
                # cells s-dofs t-dofs st-dofs work Linf [rate] L2 [rate] H1 [rate]
                vals = [x for x in t if x != "-"]
                cells, sd, td, std, work = (int(v) for v in vals[:5])
                errs = [float(v) for v in vals[5:] if "e" in v]
                rows.append(dict(cells=cells, s_dofs=sd, t_dofs=td, st_dofs=std, work=work,
                                 linf=errs[0], l2=errs[1], h1=errs[2]))
                j += 1
            for r, run in zip(rows, cur_runs):
                assert r["cells"] == run["cells"] and r["s_dofs"] == run["s_dofs"]
                run.update(r)
            tables.append(dict(k=k, runs=cur_runs))
            cur_runs = []
            i = j
            continue
        i += 1
    names = ["tf01", "tf02", "tf03", "tf04", "tf05", "tf06", "tf07", "tf08"]
    out = {}
    assert len(tables) == 24, len(tables)
    for c, name in enumerate(names):
        out[name] = tables[3 * c:3 * c + 3]
    # the JSON parameter sets, so tests can re-create each config without the reference tree
    params = {}
    for name in names:
        params[name] = json.load(open(os.path.join(REF, "tests/json", name + ".json")))
    return {"tables": out, "params": params}


def golden_tp_01_text():
    """The printed text of tests/tp_01.output, one block per parameter file (each block ends after its "Iteration count
    table"): fixture for the product's output formatter (dealii-stfem_b200/tp_01.py)."""
    txt = open(os.path.join(REF, "tests/tp_01.output")).read().splitlines()
    names = ["tf01", "tf02", "tf03", "tf04", "tf05", "tf06", "tf07", "tf08"]
    blocks, cur, i = [], [], 0
    while i < len(txt):
        cur.append(txt[i])
        if txt[i].startswith("Iteration count table"):
            i += 1
            while i < len(txt) and txt[i].strip():
                cur.append(txt[i])
                i += 1
            cur.append("")
            blocks.append(cur)
            cur = []
        i += 1
    assert len(blocks) == 8, len(blocks)
    # three blocks with three different column-width patterns are enough to pin the table writer (the numbers of all
    # eight are in tp_01.json); the fixture stays a small excerpt of the reference's output file
    keep = ("tf03", "tf05", "tf06")
    return {n: b for n, b in zip(names, blocks) if n in keep}


def golden_transfer_01():
    """Tables in the order main() runs them (tests/transfer_01.cc:772-776): DG, CGP with one time step per solve,
    then DG, CGP with two; inside each, degree index j = 1..3 and i = 2..4 (tau = 2^-(i+1))."""
    txt = open(os.path.join(REF, "tests/transfer_01.output")).read().splitlines()
    tables = []
    i = 0
    while i < len(txt):
        if txt[i].startswith("cells s-dofs"):
            rows = []
            j = i + 1
            while j < len(txt) and txt[j].strip():
                vals = [x for x in txt[j].split() if x != "-"]
                cells, sd, td, std = (int(v) for v in vals[:4])
                errs = [float(v) for v in vals[4:] if "e" in v]
                rows.append(dict(cells=cells, s_dofs=sd, t_dofs=td, st_dofs=std, linf=errs[0], l2=errs[1], h1=errs[2]))
                j += 1
            tables.append(rows)
            i = j
            continue
        i += 1
    assert len(tables) == 12 and all(len(t) == 3 for t in tables), [len(t) for t in tables]
    out = []
    for s, (ttype, nts) in enumerate((("DG", 1), ("CGP", 1), ("DG", 2), ("CGP", 2))):
        for j in range(1, 4):
            rows = tables[3 * s + j - 1]
            out.append(dict(timeType=ttype, nTimestepsAtOnce=nts, feDegree=j if ttype == "DG" else j + 1,
                            runs=[dict(r, tau_exponent=-(i + 1)) for i, r in zip(range(2, 5), rows)]))
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at %s (fixtures are committed; nothing to do)" % REF)
    for name, fn in (("tp_02", golden_tp_02), ("transfer_02", golden_transfer_02),
                     ("tp04", golden_tp04), ("tp_01", golden_tp_01), ("transfer_01", golden_transfer_01), ("tp_01_text", golden_tp_01_text)):
        data = fn()
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(data, f, indent=0, separators=(",", ":"))
        print(name, "ok")


if __name__ == "__main__":
    main()
