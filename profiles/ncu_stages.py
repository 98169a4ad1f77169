#!/usr/bin/env python
"""Per-stage share of instructions / stall samples of the Jaccard kernel in an .ncu-rep: source
lines are grouped by the stage markers in csrc/jaccard.cu (the .so must be the profiled build).

    python profiles/ncu_stages.py <rep> <libnsm_b200.so> <mangled-kernel-substring>
"""
import csv
import io
import pathlib
import subprocess
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
from ncu_lines import line_table  # noqa: E402

MARKS = [("warp_intersect_count", "uint32_t warp_intersect_count("),
         ("warp_intersect_steps (nested levels)", "ulonglong2 warp_intersect_steps("),
         ("bound_intersection", "uint32_t bound_intersection("),
         ("kernel prologue", "jaccard_allpairs_kernel(const JaccardParams p)"),
         ("flush_out", "auto flush_out"), ("level accessors", "auto left_level"),
         ("unit staging (TMA, any words)", "while (unit < n_units) {"),
         ("stage A any", "// ---- stage A:"), ("stage B bound", "// ---- stage B:"),
         ("stage C exact", "// ---- stage C:"), ("compaction", "// ---- threshold compaction"),
         ("epilogue", "    if (!packets && !coded) flush_out();")]


def main():
    rep, so, ksub = sys.argv[1:4]
    src = (pathlib.Path(__file__).resolve().parents[1] / "napkon-string-matching_b200" / "csrc" /
           "jaccard.cu").read_text().splitlines()
    starts = [(name, next(i for i, l in enumerate(src, 1) if pat in l)) for name, pat in MARKS]
    ranges = [(n, a, starts[i + 1][1] - 1 if i + 1 < len(starts) else 10 ** 9)
              for i, (n, a) in enumerate(starts)]
    table = line_table(so, ksub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ai, si = hdr.index("Address"), hdr.index("# Samples")
    ei, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    body = [r for r in rows[2:] if len(r) > ti and r[ai].startswith("0x")]
    base = min(int(r[ai], 16) for r in body)
    agg = {n: [0.0, 0.0, 0.0] for n, _, _ in ranges}
    agg["inlined headers (popc, ldg, ddiv, shfl, ...)"] = [0.0, 0.0, 0.0]
    for r in body:
        key = table.get(int(r[ai], 16) - base)
        name = "inlined headers (popc, ldg, ddiv, shfl, ...)"
        if key and key[0][0] == "jaccard.cu":
            name = next((n for n, a, b in ranges if a <= key[0][1] <= b), name)
        a = agg[name]
        a[0] += float(r[si] or 0); a[1] += float(r[ei] or 0); a[2] += float(r[ti] or 0)
    ts, te = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
    for n, a in agg.items():
        print(f"{n:46s} inst {100*a[1]/te:5.1f}%  samples {100*a[0]/ts:5.1f}%  active lanes {a[2]/max(a[1],1):5.1f}")
    print(f"total warp instructions {te:.3e}")


if __name__ == "__main__":
    main()
