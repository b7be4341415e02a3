"""Summary of the element-wise parity errors the `-m gpu` tests record (gpurun_out/parity_errors.jsonl, written by
tests/tolerances.py): one line per test case, worst cases per dtype.  Output: profiles/r02_parity_errors.md.

    python tools/parity_table.py gpurun_out/parity_errors.jsonl > profiles/r02_parity_errors.md
"""
import json
import sys

rows = {}
for line in open(sys.argv[1]):
    line = line.strip()
    if line:
        r = json.loads(line)
        rows[r["what"]] = r  # the file is appended to run after run: keep the latest record of a case

def short(w):
    return w.replace("tests/", "").replace(" (call)", "").replace("|", "/")

groups = {"fp64 (bound 1e-12)": [], "fp32 (bound 1e-5; 4e-5 for M > 512; 3e-5 for max|W| > 0.5)": []}
for w, r in rows.items():
    (groups["fp64 (bound 1e-12)"] if r["tol"] <= 1e-9 else groups[list(groups)[1]]).append((w, r))
print("# Element-wise relative errors of E_loc against the float64 oracle, as recorded by the GPU tests (round 2)\n")
print("`rel = |x - ref| / max(|ref|, floor)`, floor = 1 % (fp64) / 10 % (fp32) of `max|ref|` (tests/tolerances.py).\n")
for name, g in groups.items():
    if not g:
        continue
    worst = max(r["max_rel"] for _, r in g)
    print(f"## {name}: {len(g)} cases, worst {worst:.2e}\n")
    print("| case | elements | max rel | median rel | bound |\n|---|---|---|---|---|")
    for w, r in sorted(g, key=lambda t: -t[1]["max_rel"]):
        print(f"| {short(w)} | {r.get('n', '')} | {r['max_rel']:.2e} | {r.get('median_rel', 0):.2e} | {r['tol']:.0e} |")
    print()
