#!/usr/bin/env python
"""Attributes the per-SASS-instruction metrics of an .ncu-rep to CUDA source lines, using the
line table nvdisasm prints for the kernel's cubin (the .so must be the one that was profiled).

    python profiles/ncu_lines.py <rep> <libnsm_b200.so> <kernel-substring> [top]
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import pathlib
from collections import defaultdict


def line_table(so, kernel_sub):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(pathlib.Path(so).resolve())], cwd=d, check=True,
                       capture_output=True)
        table = {}
        for cubin in pathlib.Path(d).glob("*.cubin"):
            txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
            in_kernel, cur = False, None
            for ln in txt.splitlines():
                m = re.match(r"\s*\.text\.(\S+):", ln)
                if m:
                    in_kernel = kernel_sub in m.group(1)
                    continue
                if not in_kernel:
                    continue
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = (pathlib.Path(m.group(1)).name, int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m and cur:
                    table[int(m.group(1), 16)] = (cur, m.group(2).strip())
            if table:
                return table
    return {}


def main():
    rep, so, ksub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(so, ksub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ai, si = hdr.index("Address"), hdr.index("# Samples")
    ei, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    body = [r for r in rows[2:] if len(r) > ti and r[ai].startswith("0x")]
    base = min(int(r[ai], 16) for r in body)
    agg = defaultdict(lambda: [0.0, 0.0, 0.0])
    for r in body:
        key = table.get(int(r[ai], 16) - base, (("?", 0), ""))[0]
        a = agg[key]
        a[0] += float(r[si] or 0); a[1] += float(r[ei] or 0); a[2] += float(r[ti] or 0)
    tot_s = sum(a[0] for a in agg.values()) or 1
    tot_e = sum(a[1] for a in agg.values()) or 1
    print(f"total samples {tot_s:.0f}, warp instructions {tot_e:.3e}")
    src_cache = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        f, n = key
        if f not in src_cache:
            cands = list(pathlib.Path(__file__).resolve().parents[1].rglob(f)) if f != "?" else []
            src_cache[f] = cands[0].read_text().splitlines() if cands else []
        text = src_cache[f][n - 1].strip() if 0 < n <= len(src_cache[f]) else ""
        print(f"{100*a[1]/tot_e:5.1f}% inst {100*a[0]/tot_s:5.1f}% smp  thr/inst {a[2]/max(a[1],1):4.1f}  {f}:{n}  {text[:90]}")


if __name__ == "__main__":
    main()
