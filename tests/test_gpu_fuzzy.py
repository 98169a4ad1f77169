"""GPU suite: the fuzzy_match kernels (qratio_flat.cu: distance bound + integer keep test;
qratio.cu: compare_terms over levels; qratio_long.cu: one warp per pair for strings beyond 512
characters) against the C oracle (textbook dynamic programme).  Bit-exact triples."""
import numpy as np
import pytest

from conftest import assert_same_triples
from oracle import c_oracle
from napkon_string_matching.gpu import lib as nsmlib
from napkon_string_matching.gpu import pack

pytestmark = pytest.mark.gpu

ALPHA = list("abcdefghijklmnopqrstuvwxyzäöüß0123456789 ")


def strings(rng, n, lo, hi, alpha=ALPHA, max_k=1, min_k=1):
    return [["".join(rng.choice(alpha, size=int(rng.integers(lo, hi + 1)))).strip()
             for _ in range(int(rng.integers(min_k, max_k + 1)))] for _ in range(n)]


def check(engine, pl, pr, thr, flat=False, **kw):
    got = engine.all_pairs(engine.upload(pl), engine.upload(pr), thr, flat=flat, **kw)
    info = engine.last_info
    want, oflags = c_oracle.all_pairs(pl, pr, thr, flat=flat)
    assert_same_triples((got["left"], got["right"], got["score"]), (want["left"], want["right"], want["score"]))
    assert bool(info["flags"] & nsmlib.FLAG_EMPTY_ITEM) == bool(oflags & c_oracle.FLAG_INDEX_ERROR)
    return got, info


@pytest.mark.parametrize("thr", [-1.0, 0.0, 0.2, 0.5, 0.7, 0.9, 1.0, 1.5, float("nan")])
@pytest.mark.parametrize("flat", [True, False])
def test_flat_kernel_threshold_regimes(engine, thr, flat):
    """Every threshold region of the integer keep test (dist <= dmax[m + n]) and of the distance
    bound, with empty strings, items without a level, and sizes that straddle the 64-item tile
    and the 512-thread block."""
    rng = np.random.default_rng(11)
    L, R = strings(rng, 150, 0, 90, max_k=1, min_k=0), strings(rng, 700, 0, 120, max_k=1, min_k=0)
    L[3], L[64], R[0], R[511], R[512] = [""], ["abc abc"], ["abc abc"], [""], ["x"]
    L += [list(v) for v in R[100:120]]        # exact copies: score 1.0 (0.5 through compare_terms)
    pl, pr = pack.pack_strings(L, R)
    got, info = check(engine, pl, pr, thr, flat=flat)
    if thr == 0.7:
        assert info["stats"]["bound_pairs"] == len(L) * len(R)
        assert info["stats"]["candidates"] < 0.5 * info["stats"]["bound_pairs"]   # the bound prunes
    if thr != thr or thr > 1.0:
        assert len(got) == 0


def test_small_right_side_against_many_left_items(engine):
    """Fewer right items than one tile of left items (a launch then runs with few threads): the
    tile staging must not depend on the thread count (round-1 advisor finding)."""
    rng = np.random.default_rng(12)
    for n_r in (1, 20, 32, 33):
        L, R = strings(rng, 97, 1, 70), strings(rng, n_r, 1, 70)
        check(engine, *pack.pack_strings(L, R), 0.3, flat=True)
        L, R = strings(rng, 45, 1, 40, max_k=3, min_k=0), strings(rng, n_r, 1, 40, max_k=3, min_k=0)
        check(engine, *pack.pack_strings(L, R), 0.3)


def test_long_strings_every_pass(engine):
    """Level strings beyond 512 characters on the left, on the right and on both sides: the
    swapped pass and the warp-per-pair kernel; flat and levelled."""
    rng = np.random.default_rng(13)
    alpha = list("abcde fgh")

    def side(n_short, n_long, k):
        s = strings(rng, n_short, 0, 300, alpha, max_k=k)
        s += strings(rng, n_long, 513, 2100, alpha, max_k=k)
        order = rng.permutation(len(s))
        return [s[i] for i in order]

    for k in (1, 3):
        L, R = side(70, 5, k), side(40, 4, k)
        R[0] = [L[0][0]] if k == 1 else list(L[0])         # an exact long/any copy
        pl, pr = pack.pack_strings(L, R)
        assert int(pl.class_end[-1]) < pl.n_items and int(pr.class_end[-1]) < pr.n_items
        for thr in (0.0, 0.4):
            check(engine, pl, pr, thr, flat=(k == 1))
        # only one side long
        check(engine, *pack.pack_strings(L, strings(rng, 33, 0, 200, alpha, max_k=k)), 0.3, flat=(k == 1))
        check(engine, *pack.pack_strings(strings(rng, 33, 0, 200, alpha, max_k=k), R), 0.3, flat=(k == 1))
    # a row block that cuts through the long items
    pl, pr = pack.pack_strings(side(70, 9, 1), side(40, 6, 1))
    got = engine.all_pairs(engine.upload(pl), engine.upload(pr), 0.3, flat=True, rows=(60, 76))
    want, _ = c_oracle.all_pairs(pl, pr, 0.3, flat=True, l_begin=60, l_end=76)
    assert_same_triples((got["left"], got["right"], got["score"]), (want["left"], want["right"], want["score"]))


def test_alphabet_of_400_code_points(engine):
    """More than 255 distinct code points over both sides: code points of one side only share a
    code; scores are unchanged."""
    rng = np.random.default_rng(14)
    common = [chr(c) for c in range(0x61, 0x61 + 26)] + [chr(0x4E00 + i) for i in range(150)]
    only_l = [chr(0x0400 + i) for i in range(120)]
    only_r = [chr(0x0E00 + i) for i in range(110)]
    L = strings(rng, 80, 1, 90, common + only_l)
    R = strings(rng, 90, 1, 90, common + only_r)
    pl, pr = pack.pack_strings(L, R)
    assert pl.n_alphabet <= 255 and len(set("".join(x for v in L + R for x in v))) > 255
    for thr in (0.0, 0.3):
        check(engine, pl, pr, thr, flat=True)
    with pytest.raises(pack.PackError):     # more than 253 code points on BOTH sides: genuinely unsupported
        wide = [chr(0x4E00 + i) for i in range(300)]
        pack.pack_strings([["".join(wide)]], [["".join(wide)]])


@pytest.mark.parametrize("k", [1, 3])
def test_category_masks_through_the_fuzzy_kernels(engine, k):
    rng = np.random.default_rng(15)
    L, R = strings(rng, 90, 1, 60, max_k=k), strings(rng, 140, 1, 60, max_k=k)
    pl, pr = pack.pack_strings(L, R)
    lm = rng.integers(0, 8, size=len(L)).astype(np.uint64)
    rm = rng.integers(0, 8, size=len(R)).astype(np.uint64)
    for mode in (nsmlib.CAT_LIST_LIST, nsmlib.CAT_MEMBER):
        # masks are indexed by stored position, like the cohort
        got = engine.all_pairs(engine.upload(pl), engine.upload(pr), 0.2, flat=False,
                               l_cat=engine.upload_masks(lm[pl.perm]), r_cat=engine.upload_masks(rm[pr.perm]),
                               cat_mode=mode)
        want, _ = c_oracle.all_pairs(pl, pr, 0.2, l_cat=lm[pl.perm], r_cat=rm[pr.perm], cat_mode=mode)
        assert_same_triples((got["left"], got["right"], got["score"]), (want["left"], want["right"], want["score"]))
        assert 0 < len(got)


def test_levels_deeper_than_four(engine):
    rng = np.random.default_rng(16)
    L, R = strings(rng, 60, 1, 50, max_k=9, min_k=0), strings(rng, 75, 1, 80, max_k=7, min_k=0)
    pl, pr = pack.pack_strings(L, R)
    for thr in (0.0, 0.3):
        check(engine, pl, pr, thr)
