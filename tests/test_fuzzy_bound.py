"""CPU suite: soundness of what qratio_flat.cu prunes with, restated in numpy / pure Python from
the packed arrays — (1) the character-counter distance bound never exceeds the Indel distance
(bucket merging, saturation at 255 and the one-sided code collapse included), (2) the integer keep
test `dist <= dmax[lensum]` equals `score >= threshold` on the oracle's own float64 map, so a pair
the kernel skips is PROVEN below the threshold and a pair it keeps is exactly the oracle's."""
import math

import numpy as np

from oracle import reference_port as port
from napkon_string_matching.gpu import pack


def counters(p, g):
    """The 32 byte counters of level g as integers."""
    return p.level_hist[g].view(np.uint8).astype(np.int64)


def level_text(p, g):
    return bytes(p.level_string_codes(g).tolist())


def dmax_table(thr, top):
    """dmax[s] = largest dist with the oracle's QRatio/100 map >= thr (-1: none), by full scan."""
    out = []
    for s in range(top + 1):
        ok = [d for d in range(s + 1) if (0.0 if s == 0 else port.indel_ratio_from_counts(d, s)) >= thr]
        out.append(max(ok) if ok else -1)
        if s:   # the map is monotone in dist, which is what makes one table entry per s enough
            vals = [port.indel_ratio_from_counts(d, s) for d in range(s + 1)]
            assert all(a >= b for a, b in zip(vals, vals[1:]))
    return out


def test_counter_distance_is_a_lower_bound_of_the_indel_distance():
    rng = np.random.default_rng(1)
    small = list("abcdefghijklmnopqrstuvwxyzäöüß 0123456789")          # 41 codes: buckets are shared
    wide_l = small + [chr(0x400 + i) for i in range(150)]              # > 255 over both sides:
    wide_r = small + [chr(0xE00 + i) for i in range(150)]              # one-sided code points collapse
    cases = [(small, small, 120), (small[:6], small[:6], 700), (wide_l, wide_r, 150)]
    for al, ar, hi in cases:
        L = [["".join(rng.choice(al, size=int(rng.integers(0, hi))))] for _ in range(40)]
        R = [["".join(rng.choice(ar, size=int(rng.integers(0, hi))))] for _ in range(40)]
        L.append(["a" * 300 + "b" * 280]); R.append(["a" * 290 + "c" * 3])   # counters saturate at 255
        pl, pr = pack.pack_strings(L, R)
        assert pl.level_hist.shape == (pl.n_levels, 8)
        for gl in range(pl.n_levels):
            a = level_text(pl, gl)
            assert counters(pl, gl).sum() <= len(a)   # saturation only lowers
            for gr in range(0, pr.n_levels, 3):
                b = level_text(pr, gr)
                dist = len(a) + len(b) - 2 * port.lcs_length(a, b)
                d1 = int(np.abs(counters(pl, gl) - counters(pr, gr)).sum())
                assert d1 <= dist, (a, b, d1, dist)


def test_integer_keep_test_equals_the_float_threshold():
    for thr in (-0.5, 0.0, 0.1, 0.35, 0.5, 0.7, 0.7000000000000001, 0.9, 1.0, 1.0000001, float("nan")):
        table = dmax_table(thr, 90)
        for s in range(91):
            for d in range(s + 1):
                score = 0.0 if s == 0 else port.indel_ratio_from_counts(d, s)
                assert (d <= table[s]) == (score >= thr), (thr, s, d)
        if math.isnan(thr):
            assert set(table) == {-1}


def test_bound_never_prunes_a_pair_that_reaches_the_threshold():
    rng = np.random.default_rng(2)
    alpha = list("abcdefgh ")
    base = ["".join(rng.choice(alpha, size=int(rng.integers(5, 60)))) for _ in range(30)]
    L = [[s] for s in base]
    # near copies: a few edits each, so that many pairs are at or just above / below the thresholds
    R = []
    for s in base:
        t = list(s)
        for _ in range(int(rng.integers(0, 12))):
            pos = int(rng.integers(0, len(t) + 1))
            if rng.random() < 0.5 and t:
                del t[min(pos, len(t) - 1)]
            else:
                t.insert(pos, str(rng.choice(alpha)))
        R.append(["".join(t)])
    pl, pr = pack.pack_strings(L, R)
    for thr in (0.5, 0.7, 0.85):
        table = dmax_table(thr, 200)
        pruned = kept = 0
        for gl in range(pl.n_levels):
            a = level_text(pl, gl)
            for gr in range(pr.n_levels):
                b = level_text(pr, gr)
                s = len(a) + len(b)
                d1 = int(np.abs(counters(pl, gl) - counters(pr, gr)).sum())
                score = port.qratio_processed(a.decode("latin1"), b.decode("latin1")) if a and b else 0.0
                if d1 > table[s]:
                    pruned += 1
                    assert score < thr, (a, b, score, thr)
                kept += score >= thr
        assert pruned > 0 and kept > 0
