"""CPU suite: bench.py's accounting — the evaluation and algorithmic-op counts that the headline
metric and the roofline numerator are built from — against a brute-force count
over every item pair and compare_terms step, and the JSON contract of the reference arm."""
import json
import subprocess
import sys

import numpy as np

from conftest import ROOT
from napkon_string_matching import synthetic as syn
from napkon_string_matching.gpu import pack

sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def brute_force(left, right, sizes_l, sizes_r, fuzzy):
    kl, kr = left.levels_per_item(), right.levels_per_item()
    ol, orr = left.item_level_off.astype(np.int64), right.item_level_off.astype(np.int64)
    evals = ops = 0
    for i in range(left.n_items):
        for j in range(right.n_items):
            if kl[i] == 0 or kr[j] == 0:
                continue
            for t in range(1, max(kl[i], kr[j]) + 1):
                a = int(sizes_l[ol[i] + min(t, kl[i] - 1)])
                b = int(sizes_r[orr[j] + min(t, kr[j] - 1)])
                evals += 1
                ops += 8 * -(-min(a, b) // 64) * max(a, b) if fuzzy else a + b
    return evals, ops


def test_jaccard_counts_equal_brute_force():
    for maker, args in ((syn.token_id_level_sets, (60, 3)), (syn.term_level_sets, (50, 4))):
        a, b = maker(*args), maker(args[0] + 7, args[1] + 1)
        if maker is syn.token_id_level_sets:
            pl, pr = pack.pack_suffix_id_sets(*a, 30000), pack.pack_suffix_id_sets(*b, 30000)
        else:
            rank = pack.frequency_rank([a[1], b[1]], 20000)
            pl, pr = pack.pack_part_id_sets(*a, 20000, rank), pack.pack_part_id_sets(*b, 20000, rank)
        evals, ops = bench.schedule_counts(pl, pr)
        want = brute_force(pl, pr, pl.level_sizes(), pr.level_sizes(), fuzzy=False)
        assert (evals, ops) == tuple(map(float, want))


def test_flat_string_counts_equal_brute_force():
    rng = np.random.default_rng(0)
    mk = lambda n: [["x" * int(k)] for k in rng.integers(0, 200, size=n)]  # noqa: E731
    pl, pr = pack.pack_strings(mk(40), mk(55))
    evals, ops = bench.schedule_counts(pl, pr)
    # flat scoring evaluates every pair once, whatever the lengths
    ll, lr = pl.level_lengths(), pr.level_lengths()
    want_ops = sum(8 * -(-min(int(a), int(b)) // 64) * max(int(a), int(b)) for a in ll for b in lr)
    assert evals == 40.0 * 55.0 and ops == float(want_ops)


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["workload"] == "tokenids50k"
    from oracle import ref_arm

    assert line["cpu_baseline"]["kind"] == ("reference" if ref_arm.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["scaling"] == "strong"
    assert {"workload", "item_pairs_per_step", "pair_scores_per_step", "threshold"} <= set(line["config"])
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
