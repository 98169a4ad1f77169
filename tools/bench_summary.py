"""Prints the key figures of bench.py JSON lines: python tools/bench_summary.py <log> [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        print(path, "unreadable:", exc)
        continue
    r = d["roofline"]
    print(f"{d['config']['workload']:14s} value {d['value']:.4g} {d['unit']}  {d['ms_per_step']:.2f} ms/step  "
          f"kernel {r['kernel_ms_per_launch']:.2f} ms x {r['launches_per_step']:.0f}  frac {r['frac']:.3f}  "
          f"e2e {d['e2e']['value']:.4g} ({d['e2e']['ms_per_step']:.1f} ms)  kept {d['config']['kept_pairs_per_step']}  "
          f"stats {d['kernel_stats_per_step']}")
