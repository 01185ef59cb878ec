"""Command-line front end with the reference's parameter files and output format (reference tests/tp_01.cc:727-848):

    python -m dealii_stfem_b200.tp_01 --file tests/json/tf03.json --dim 2 [--precondition_double]

reads a JSON parameter file with the keys of include/parameters.h:92-144, runs the nDegCycles x nRefCycles loop of
convergence tests on the GPU (HeatWaveProblem = the convergence_test lambda) and prints what tp_01 prints on rank 0:
the per-run header (tp_01.cc:101-104, 210, 703-707), one "Convergence table k=..." per degree (:712-760) and the
"Iteration count table" (:761-764), so that the output can be diffed against tests/tp_01.output.
The text tables follow deal.II's TableHandler::write_text(table_with_headers) / ConvergenceTable rules, restated in
`TextTable` below.  Host glue only; no oracle import."""
import argparse
import json
import math
import sys


class TextTable:
    """deal.II TableHandler (table_with_headers) + ConvergenceTable::evaluate_convergence_rates(reduction_rate_log2):
    a column is as wide as its longest entry; a rate column joins its value column in a super-column headed by the key;
    headers are centred over the column ((width - len) / 2 blanks in front, the rest behind, lengths counted in BYTES as
    std::string does, hence the off-centre look of the UTF-8 headers); entries are right-aligned; every column is followed
    by one blank.  Floating-point entries: precision 4, fixed, unless set_precision / set_scientific say otherwise."""

    def __init__(self):
        self.order, self.cols, self.fmt, self.rates = [], {}, {}, set()

    def add_value(self, key, value):
        if key not in self.cols:
            self.order.append(key)
            self.cols[key] = []
        self.cols[key].append(value)

    def set_format(self, key, precision, scientific):
        self.fmt[key] = (precision, scientific)

    def evaluate_convergence_rates(self, key):
        self.rates.add(key)

    def _entry(self, key, v):
        if isinstance(v, str):
            return v
        if isinstance(v, int):
            return str(v)
        prec, sci = self.fmt.get(key, (4, False))
        if math.isnan(v):
            return "nan"
        return ("%.*e" if sci else "%.*f") % (prec, v)

    def write_text(self):
        n_rows = max(len(c) for c in self.cols.values()) if self.cols else 0
        groups = []                                    # (header key, [list of entry columns])
        for key in self.order:
            entries = [self._entry(key, v) for v in self.cols[key]]
            sub = [entries]
            if key in self.rates:
                vals = self.cols[key]
                rate = ["-"]
                for a, b in zip(vals[:-1], vals[1:]):
                    q = a / b if b != 0 else float("nan")
                    rate.append("nan" if (math.isnan(q) or q <= 0) else "%.2f" % math.log2(q))
                sub.append(rate)
            groups.append((key, sub))
        lines = []
        head = ""
        widths = []
        for key, sub in groups:
            w = [max(len(e) for e in col) if col else 0 for col in sub]
            total = sum(w) + len(w) - 1
            klen = len(key.encode("utf-8"))
            if total < klen:                            # header longer than the columns under it: widen the first one
                w[0] += klen - total
                total = klen
            front = (total - klen) // 2
            head += " " * front + key + " " * (total - klen - front) + " "
            widths.append(w)
        lines.append(head)
        for r in range(n_rows):
            ln = ""
            for (key, sub), w in zip(groups, widths):
                for col, wc in zip(sub, w):
                    ln += (col[r] if r < len(col) else "").rjust(wc) + " "
            lines.append(ln)
        return "\n".join(lines) + "\n"


def _g(v):
    """default ostream formatting of a double (6 significant digits)."""
    return "%g" % v


def run_header(row):
    """tp_01.cc:101-104, 210, 703-707."""
    avg = row["iterations"] / row["timesteps"]
    return (":: Number of active cells: %d\n:: Number of degrees of freedom: %d\n:: Min Level 0  Max Level %d\n"
            "Average GMRES iterations %s (%d gmres_iterations / %d timesteps)\n\n"
            % (row["cells"], row["s_dofs"], len(row["levels"]), _g(avg), row["iterations"], row["timesteps"]))


ERR_KEYS = (("L∞-L∞", "linf"), ("L2-L2", "l2"), ("L2-H1_semi", "h1"))


def convergence_table(rows, st_convergence=True, legacy_work=False):
    """tp_01.cc:712-723, 739-757 for the rows of one degree.  legacy_work: the stored tests/tp_01.output predates
    tp_01.cc:715 (work = n_dofs * n_blocks * iterations) and shows st-dofs * iterations in that column."""
    t = TextTable()
    for r in rows:
        t.add_value("cells", int(r["cells"]))
        t.add_value("s-dofs", int(r["s_dofs"]))
        t.add_value("t-dofs", int(r["t_dofs"]))
        t.add_value("st-dofs", int(r["timesteps"] * r["s_dofs"] * r["t_dofs"]))
        t.add_value("work", int(r["s_dofs"] * r["t_dofs"] * r["iterations"] * (r["timesteps"] if legacy_work else 1)))
        for key, name in ERR_KEYS:
            t.add_value(key, float(r[name]) if st_convergence else float("nan"))
    for key, _ in ERR_KEYS:
        t.set_format(key, 5, True)
        t.evaluate_convergence_rates(key)
    return t.write_text()


def iteration_table(degrees, refinements, averages):
    """tp_01.cc:724, 736, 761-764: averages[j][i] = average iterations of degree j, refinement i."""
    t = TextTable()
    for j, k in enumerate(degrees):
        t.add_value("k \\ r", int(k))
        for i, rf in enumerate(refinements):
            t.add_value(str(rf), float(averages[j][i]))
    return t.write_text()


def run(params_json, dim, out=sys.stdout, precondition_float=True, ctx=None, max_steps=None):
    """test<Number, NumberPreconditioner>() of tp_01.cc:727-765 on the GPU."""
    from . import capi
    from .driver import HeatWaveProblem, parse_parameters
    p = parse_parameters(params_json, dim)
    own = ctx is None
    if own:
        ctx = capi.Context(0)
    k0, r0 = p["feDegree"], p["refinement"]
    degrees = list(range(k0, k0 + p["nDegCycles"]))
    refinements = list(range(r0, r0 + p["nRefCycles"]))
    averages, all_rows = [], []
    for k in degrees:
        rows = []
        for rf in refinements:
            prob = HeatWaveProblem(ctx, p, dim, rf, k, mg_number_type=capi.F32 if precondition_float else capi.F64)
            if not p["spaceTimeConvergenceTest"]:
                prob.functional_file = p["functionalFile"]          # appended to like tp_01.cc:620
            row = prob.run(max_steps=max_steps)
            prob.close()
            out.write(run_header(row))
            rows.append(row)
        out.write("Convergence table k=%d\n" % k)
        out.write(convergence_table(rows, p["spaceTimeConvergenceTest"]))
        out.write("\n")
        averages.append([r["iterations"] / r["timesteps"] for r in rows])
        all_rows.append(rows)
    out.write("Iteration count table\n")
    out.write(iteration_table(degrees, refinements, averages))
    out.write("\n")
    if own:
        ctx.close()
    return all_rows


def main(argv=None):
    ap = argparse.ArgumentParser(description="tp_01 of dealii-stfem on the B200 path")
    ap.add_argument("-f", "--file", required=True, help="path to the JSON parameter file")
    ap.add_argument("-d", "--dim", type=int, default=2, help="spatial dimensions")
    ap.add_argument("--precondition_double", action="store_true",
                    help="multigrid levels in double (reference default -p: float)")
    ap.add_argument("--max_steps", type=int, default=None, help="stop every run after this many solves (not a reference option)")
    a = ap.parse_args(argv)
    with open(a.file) as f:
        pj = json.load(f)
    run(pj, a.dim, precondition_float=not a.precondition_double, max_steps=a.max_steps)


if __name__ == "__main__":
    main()
