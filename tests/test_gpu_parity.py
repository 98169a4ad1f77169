"""GPU suite: the CUDA path (through the C ABI) against the oracle and the reference-generated
golden vectors.  Bit-exact: same pair set, same float64 scores."""
import numpy as np
import pytest

from conftest import assert_same_triples, load_golden
from oracle import c_oracle
from napkon_string_matching import synthetic as syn
from napkon_string_matching.gpu import lib as nsmlib
from napkon_string_matching.gpu import pack
from napkon_string_matching.text.tokenize import gen_comp_value

pytestmark = pytest.mark.gpu


def run_gpu(engine, pl, pr, thr, **kw):
    out = engine.all_pairs(engine.upload(pl), engine.upload(pr), thr, **kw)
    return out, engine.last_info


def check_against_oracle(engine, pl, pr, thr, flat=False, **kw):
    out, info = run_gpu(engine, pl, pr, thr, flat=flat, **kw)
    want, _ = c_oracle.all_pairs(pl, pr, thr, flat=flat)
    assert info["count"] == len(out)
    assert_same_triples((out["left"], out["right"], out["score"]),
                        (want["left"], want["right"], want["score"]))
    return out, info


def _golden_levels(name, column):
    meta, inputs, arrays = load_golden(name)
    L = [gen_comp_value(v) if v is not None else None for v in inputs["left"][column]]
    R = [gen_comp_value(v) if v is not None else None for v in inputs["right"][column]]
    lkeep = np.array([i for i, v in enumerate(L) if v is not None])
    rkeep = np.array([i for i, v in enumerate(R) if v is not None])
    return meta, arrays, [L[i] for i in lkeep], [R[i] for i in rkeep], lkeep, rkeep


@pytest.mark.parametrize("name,column", [("cfg1_400_term_jaccard", "Term"),
                                         ("cfg2_300_tokenids_jaccard", "TokenIds"),
                                         ("variable_80_jaccard", "Variable")])
def test_jaccard_matches_reference_golden(engine, name, column):
    meta, arrays, L, R, lkeep, rkeep = _golden_levels(name, column)
    pl, pr = pack.pack_sets(L, R)
    out, info = run_gpu(engine, pl, pr, meta["kwargs"]["score_threshold"])
    assert info["flags"] == 0
    assert_same_triples((lkeep[out["left"]], rkeep[out["right"]], out["score"]),
                        (arrays["left_pos"], arrays["right_pos"], arrays["score"]))


@pytest.mark.parametrize("name,column", [("fuzzy_150_term", "Term"),
                                         ("fuzzy_60_question_str", "Question")])
def test_fuzzy_matches_golden(engine, name, column):
    meta, arrays, L, R, lkeep, rkeep = _golden_levels(name, column)
    pl, pr = pack.pack_strings(pack.fuzzy_level_strings(L), pack.fuzzy_level_strings(R))
    out, info = run_gpu(engine, pl, pr, meta["kwargs"]["score_threshold"])
    assert_same_triples((lkeep[out["left"]], rkeep[out["right"]], out["score"]),
                        (arrays["left_pos"], arrays["right_pos"], arrays["score"]))


@pytest.mark.parametrize("thr", [0.0, 0.05, 0.3, 0.5, 0.9375, 1.5])
def test_jaccard_thresholds_vs_oracle(engine, thr):
    vocab = syn.vocabulary(3000)
    fl = syn.questionnaire_frame(300, 21, vocab, "hap")
    fr = syn.questionnaire_frame(517, 22, vocab, "pop")
    pl, pr = pack.pack_sets([gen_comp_value(t) for t in fl["Term"]],
                            [gen_comp_value(t) for t in fr["Term"]])
    check_against_oracle(engine, pl, pr, thr)


def test_jaccard_token_ids_2000_vs_oracle(engine):
    lens, flat = syn.token_id_level_sets(2000, syn.SEED_LEFT)
    pl = pack.pack_suffix_id_sets(lens, flat, 30000)
    lens, flat = syn.token_id_level_sets(2000, syn.SEED_RIGHT)
    pr = pack.pack_suffix_id_sets(lens, flat, 30000)
    out, info = check_against_oracle(engine, pl, pr, 0.1)
    assert info["stats"]["candidates"] >= len(out)
    assert info["stats"]["candidates"] < info["item_pairs"]  # the filter did prune


def test_jaccard_exact_signature_vocabulary(engine):
    rng = np.random.default_rng(3)
    mk = lambda n: [[[f"w{int(x)}" for x in rng.integers(0, 120, size=int(rng.integers(1, 9)))]
                     for _ in range(int(rng.integers(1, 5)))] for _ in range(n)]
    pl, pr = pack.pack_sets(mk(257), mk(300))
    assert pl.exact_bits
    out, info = check_against_oracle(engine, pl, pr, 0.2)
    assert info["stats"]["level_merges"] == 0  # popcount path only


def test_jaccard_flat_and_ragged_edges(engine):
    # flat = the scalar score function on K = 1 items
    L = [[["a", "b", "c"]], [["x"]], [["a"]], [["q", "r", "s", "t", "u"]]]
    R = [[["b", "c", "d"]], [["x"]], [["y"]]]
    pl, pr = pack.pack_sets(L, R)
    out, _ = check_against_oracle(engine, pl, pr, 0.0, flat=True)
    assert len(out) == 12
    got = {(int(a), int(b)): s for a, b, s in zip(out["left"], out["right"], out["score"])}
    assert got[(0, 0)] == 2 / 4 and got[(1, 1)] == 1.0 and got[(2, 2)] == 0.0
    # unequal level counts, deep clamp, one very long set
    big = [f"t{i}" for i in range(700)]
    L = [[["x"], ["a"], ["a", "b"]], [["a"]], [big[:5], big[:300], big], [["p"]] * 9]
    R = [[["y"], ["a"]], [big[100:400]], [["p"], ["p", "q"]]]
    pl, pr = pack.pack_sets(L, R)
    check_against_oracle(engine, pl, pr, 0.0)
    check_against_oracle(engine, pl, pr, 0.26)


def test_jaccard_empty_inputs_and_flags(engine):
    pl, pr = pack.pack_sets([], [[["a"]]])
    out, info = run_gpu(engine, pl, pr, 0.1)
    assert len(out) == 0 and info["count"] == 0
    pl, pr = pack.pack_sets([[["a"], []]], [[["b"], []]])
    out, info = run_gpu(engine, pl, pr, 0.0)
    assert info["flags"] & nsmlib.FLAG_ZERO_UNION and len(out) == 0
    pl, pr = pack.pack_sets([[], [["a"]]], [[["a"]], []])
    out, info = run_gpu(engine, pl, pr, 0.0)
    assert info["flags"] & nsmlib.FLAG_EMPTY_ITEM
    got = {(int(a), int(b)): s for a, b, s in zip(out["left"], out["right"], out["score"])}
    assert got == {(0, 1): 0.0, (1, 0): 0.5}


def test_jaccard_overflow_reruns_with_exact_capacity(engine):
    lens, flat = syn.token_id_level_sets(600, 5, n_ids=200)
    p = pack.pack_suffix_id_sets(lens, flat, 200)
    want, _ = c_oracle.all_pairs(p, p, 0.05)
    from napkon_string_matching.gpu.engine import Engine

    fresh = Engine()  # arenas are grow-only: a new engine starts with none
    out = fresh.all_pairs(fresh.upload(p), fresh.upload(p), 0.05, capacity=64)
    assert fresh.last_info["reruns"] >= 1
    assert_same_triples((out["left"], out["right"], out["score"]),
                        (want["left"], want["right"], want["score"]))


def test_jaccard_row_blocks_partition_the_result(engine):
    lens, flat = syn.token_id_level_sets(900, 8)
    pl = pack.pack_suffix_id_sets(lens, flat, 30000)
    dl = engine.upload(pl)
    whole = engine.all_pairs(dl, dl, 0.2)
    parts = [engine.all_pairs(dl, dl, 0.2, rows=(b, e)) for b, e in ((0, 123), (123, 124), (124, 900))]
    cat = np.concatenate(parts)
    assert_same_triples((cat["left"], cat["right"], cat["score"]),
                        (whole["left"], whole["right"], whole["score"]))


def test_jaccard_category_masks(engine):
    rng = np.random.default_rng(9)
    lens, flat = syn.token_id_level_sets(400, 31, n_ids=500)
    p = pack.pack_suffix_id_sets(lens, flat, 500)
    lm = rng.integers(0, 8, size=400).astype(np.uint64)
    rm = rng.integers(0, 8, size=400).astype(np.uint64)
    d = engine.upload(p)
    for mode in (nsmlib.CAT_LIST_LIST, nsmlib.CAT_MEMBER):
        out = engine.all_pairs(d, d, 0.1, l_cat=engine.upload_masks(lm),
                               r_cat=engine.upload_masks(rm), cat_mode=mode)
        want, _ = c_oracle.all_pairs(p, p, 0.1, l_cat=lm, r_cat=rm, cat_mode=mode)
        assert_same_triples((out["left"], out["right"], out["score"]),
                            (want["left"], want["right"], want["score"]))


@pytest.mark.parametrize("thr", [0.0, 0.4, 0.7])
def test_fuzzy_flat_strings_vs_oracle(engine, thr):
    from napkon_string_matching.text.process import default_process

    vocab = syn.vocabulary(2000)
    sl = [[default_process(s)] for s in syn.question_strings(300, 41, vocab)]
    sr = [[default_process(s)] for s in syn.question_strings(333, 42, vocab)]
    sl[7], sr[11], sr[12] = [""], [""], ["a"]
    pl, pr = pack.pack_strings(sl, sr)
    check_against_oracle(engine, pl, pr, thr, flat=True)


def test_fuzzy_multiword_lengths_vs_oracle(engine):
    rng = np.random.default_rng(17)
    alpha = list("abcdefghij klmn")
    mk = lambda n, lo, hi: [["".join(rng.choice(alpha, size=int(rng.integers(lo, hi))))] for _ in range(n)]
    for hi in (64, 65, 129, 200, 300, 500):
        pl, pr = pack.pack_strings(mk(40, 1, hi), mk(70, max(1, hi - 70), hi + 1))
        check_against_oracle(engine, pl, pr, 0.3, flat=True)


def test_fuzzy_levels_term_shape_vs_oracle(engine):
    vocab = syn.vocabulary(3000)
    fl = syn.questionnaire_frame(150, 51, vocab, "hap")
    fr = syn.questionnaire_frame(290, 52, vocab, "pop")
    L = pack.fuzzy_level_strings([gen_comp_value(t) for t in fl["Term"]])
    R = pack.fuzzy_level_strings([gen_comp_value(t) for t in fr["Term"]])
    pl, pr = pack.pack_strings(L, R)
    check_against_oracle(engine, pl, pr, 0.45)
    check_against_oracle(engine, pl, pr, 0.0)


def test_fuzzy_empty_items_and_flags(engine):
    pl, pr = pack.pack_strings([[], ["abc"]], [["abc"], []])
    out, info = run_gpu(engine, pl, pr, 0.0)
    assert info["flags"] & nsmlib.FLAG_EMPTY_ITEM
    got = {(int(a), int(b)): s for a, b, s in zip(out["left"], out["right"], out["score"])}
    assert got == {(0, 1): 0.0, (1, 0): 0.5}


def test_jaccard_full_size_properties(engine):
    """BASELINE configs[1] scale (50k-item cohorts): too big for the oracle to enumerate, so
    check properties that do not depend on size: row blocks partition the result, the kept set
    is symmetric under swapping the sides, kept scores equal the oracle's on a sample, and a
    sample of pairs that were not kept score below the threshold in the oracle."""
    thr = 0.1
    lens, flat = syn.token_id_level_sets(50_000, syn.SEED_LEFT)
    pl = pack.pack_suffix_id_sets(lens, flat, 30000)
    lens, flat = syn.token_id_level_sets(50_000, syn.SEED_RIGHT)
    pr = pack.pack_suffix_id_sets(lens, flat, 30000)
    dl, dr = engine.upload(pl), engine.upload(pr)
    rows = (10_000, 30_000)
    whole = engine.all_pairs(dl, dr, thr, rows=rows)
    n_whole = len(whole)
    assert n_whole == engine.last_info["count"] > 10_000_000
    assert whole["left"].min() >= rows[0] and whole["left"].max() < rows[1]
    key = whole["left"].astype(np.uint64) * np.uint64(pr.n_items) + whole["right"]
    order = np.argsort(key, kind="stable")
    key, score = key[order], whole["score"][order]
    assert np.all(key[1:] != key[:-1]), "a pair was emitted twice"
    # (1) row blocks partition the result
    parts = 0
    for b, e in ((10_000, 10_001), (10_001, 17_333), (17_333, 30_000)):
        engine.all_pairs(dl, dr, thr, rows=(b, e), to_host=False)
        parts += engine.last_info["count"]
    assert parts == n_whole
    # (2) symmetry: right x left keeps the transposed set with the same scores
    swapped = engine.all_pairs(dr, dl, thr)
    sel = (swapped["right"] >= rows[0]) & (swapped["right"] < rows[1])
    skey = swapped["right"][sel].astype(np.uint64) * np.uint64(pr.n_items) + swapped["left"][sel]
    sorder = np.argsort(skey, kind="stable")
    assert np.array_equal(skey[sorder], key)
    assert np.array_equal(swapped["score"][sel][sorder].view(np.uint64), score.view(np.uint64))
    del swapped
    # (3) kept scores are the oracle's, bit for bit, on a sample
    rng = np.random.default_rng(1)
    pick = rng.choice(n_whole, size=50_000, replace=False)
    want, _ = c_oracle.score_pairs(pl, pr, whole["left"][pick], whole["right"][pick])
    assert np.array_equal(want.view(np.uint64), whole["score"][pick].view(np.uint64))
    assert np.all(whole["score"] >= thr)
    # (4) pairs that were not kept are below the threshold
    li = rng.integers(rows[0], rows[1], size=300_000).astype(np.uint32)
    ri = rng.integers(0, pr.n_items, size=300_000).astype(np.uint32)
    probe = li.astype(np.uint64) * np.uint64(pr.n_items) + ri
    pos = np.searchsorted(key, probe)
    kept = (pos < len(key)) & (key[np.minimum(pos, len(key) - 1)] == probe)
    got, _ = c_oracle.score_pairs(pl, pr, li, ri)
    assert np.array_equal(got >= thr, kept)


def _random_side(rng, n, vocab, max_k, max_size):
    items = []
    for _ in range(n):
        k = int(rng.integers(0, max_k + 1))
        items.append([[f"w{int(x)}" for x in rng.zipf(1.4, size=int(rng.integers(0, max_size + 1))) % vocab]
                      for _ in range(k)])
    return items


@pytest.mark.parametrize("seed", range(6))
def test_jaccard_random_ragged_inputs_vs_oracle(engine, seed):
    """Ragged level counts (0 .. 14: beyond the 10 staged slots), empty levels, empty items,
    vocabularies below and above the exact-bit limit; results AND exception flags must match."""
    rng = np.random.default_rng(1000 + seed)
    vocab = [30, 100, 129, 4000, 4000, 60000][seed]
    L = _random_side(rng, int(rng.integers(1, 300)), vocab, [3, 14, 6, 14, 2, 5][seed], [6, 12, 40, 8, 3, 30][seed])
    R = _random_side(rng, int(rng.integers(1, 400)), vocab, [3, 14, 6, 12, 2, 5][seed], [6, 12, 40, 8, 3, 30][seed])
    pl, pr = pack.pack_sets(L, R)
    for thr in (0.0, 0.08, 0.3, 0.55):
        got, info = run_gpu(engine, pl, pr, thr)
        want, oflags = c_oracle.all_pairs(pl, pr, thr)
        assert_same_triples((got["left"], got["right"], got["score"]),
                            (want["left"], want["right"], want["score"]))
        if thr == 0.0:  # every pair reaches exact scoring: flags are complete
            assert bool(info["flags"] & nsmlib.FLAG_ZERO_UNION) == bool(oflags & c_oracle.FLAG_ZERO_UNION)
            assert bool(info["flags"] & nsmlib.FLAG_EMPTY_ITEM) == bool(oflags & c_oracle.FLAG_INDEX_ERROR)


@pytest.mark.parametrize("seed", range(4))
def test_fuzzy_random_ragged_inputs_vs_oracle(engine, seed):
    rng = np.random.default_rng(2000 + seed)
    alpha = list("abcdefghijklmnopqrstuvwxyzäöüß0123456789 ")[: [8, 20, 41, 41][seed]]
    max_len = [40, 90, 200, 500][seed]

    def side(n, max_k):
        return [["".join(rng.choice(alpha, size=int(rng.integers(0, max_len + 1)))).strip()
                 for _ in range(int(rng.integers(0, max_k + 1)))] for _ in range(n)]

    L, R = side(int(rng.integers(1, 120)), [1, 4, 3, 2][seed]), side(int(rng.integers(1, 150)), [1, 4, 2, 3][seed])
    pl, pr = pack.pack_strings(L, R)
    for thr in (0.0, 0.35, 0.6):
        got, info = run_gpu(engine, pl, pr, thr)
        want, oflags = c_oracle.all_pairs(pl, pr, thr)
        assert_same_triples((got["left"], got["right"], got["score"]),
                            (want["left"], want["right"], want["score"]))
        assert bool(info["flags"] & nsmlib.FLAG_EMPTY_ITEM) == bool(oflags & c_oracle.FLAG_INDEX_ERROR)


def test_jaccard_nested_levels_single_intersection_and_its_fallbacks(engine):
    """gen_comp_value's levels are nested, which the kernel uses to intersect the tail ids once
    per pair.  Cover the cases where it must fall back to per-level intersections: more than 16
    steps, more than 255 tail ids, or one side not nested."""
    rng = np.random.default_rng(77)

    def suffix_items(n, max_k, per_part, vocab):
        out = []
        for _ in range(n):
            parts = [[f"w{int(x)}" for x in rng.zipf(1.3, size=int(rng.integers(1, per_part + 1))) % vocab]
                     for _ in range(int(rng.integers(1, max_k + 1)))]
            out.append([sorted({w for part in parts[-j:] for w in part}) for j in range(1, len(parts) + 1)])
        return out

    for max_k, per_part, vocab in ((6, 4, 3000), (22, 2, 3000), (3, 420, 9000)):
        L, R = suffix_items(150, max_k, per_part, vocab), suffix_items(170, max_k, per_part, vocab)
        pl, pr = pack.pack_sets(L, R)
        assert pl.nested and pr.nested and not pl.exact_bits
        for thr in (0.0, 0.2):
            _, info = check_against_oracle(engine, pl, pr, thr)
        assert info["stats"]["level_merges"] > 0
    # one side not nested: per-level path
    L = suffix_items(100, 5, 4, 3000)
    R = [[[f"w{int(x)}" for x in rng.integers(0, 3000, size=5)] for _ in range(3)] for _ in range(120)]
    pl, pr = pack.pack_sets(L, R)
    assert pl.nested and not pr.nested
    check_against_oracle(engine, pl, pr, 0.0)


def test_unsupported_inputs_raise_instead_of_falling_back(engine):
    # strings beyond 512 characters are scored (tests/test_gpu_fuzzy.py); what is refused: a level
    # string beyond the generic kernel's 16384 characters on BOTH sides of a pair
    pl, pr = pack.pack_strings([["ab" * 9000]], [["ba" * 9000]])
    with pytest.raises(nsmlib.NsmError, match="exceed the generic kernel"):
        engine.all_pairs(engine.upload(pl), engine.upload(pr), 0.1, flat=True)
    # flat scoring needs one level per item
    pl, pr = pack.pack_sets([[["a"], ["b"]]], [[["a"]]])
    with pytest.raises(nsmlib.NsmError, match="one level"):
        engine.all_pairs(engine.upload(pl), engine.upload(pr), 0.1, flat=True)
    # sets against strings
    ql, = pack.pack_strings([["abc"]])
    with pytest.raises(TypeError):
        engine.all_pairs(engine.upload(pl), engine.upload(ql), 0.1)
    # maximum sizes that ARE supported: a 60000-token level, a 512-character string
    big = [f"t{i}" for i in range(60000)]
    pl, pr = pack.pack_sets([[big]], [[big[:30000] + ["zz"]], [["t5"]]])
    out, _ = check_against_oracle(engine, pl, pr, 0.0, flat=True)
    assert sorted(out["score"].tolist()) == [1 / 60000, 30000 / 60001]
    pl, pr = pack.pack_strings([["ab" * 256]], [["ba" * 256], ["a" * 512]])
    check_against_oracle(engine, pl, pr, 0.0, flat=True)


@pytest.mark.parametrize("vocab,max_k,per_part", [(90, 4, 5), (400, 4, 9), (3000, 3, 30), (20000, 4, 6)])
def test_jaccard_two_shared_bits_filter_vs_oracle(engine, vocab, max_k, per_part):
    """High thresholds with depth D = 1 run the kernels whose stage A asks for TWO shared
    signature bits unless a side is "wild" (tail ids sharing a signature bit, or a step-1 level so
    small that one shared id could suffice).  Dense vocabularies (many shared ids, many folded
    tail bits), tiny levels and every threshold region around 1/2 must match the oracle."""
    rng = np.random.default_rng(vocab + per_part)

    def suffix_items(n):
        out = []
        for _ in range(n):
            parts = [[f"w{int(x)}" for x in rng.zipf(1.2, size=int(rng.integers(1, per_part + 1))) % vocab]
                     for _ in range(int(rng.integers(1, max_k + 1)))]
            out.append([sorted({w for part in parts[-j:] for w in part}) for j in range(1, len(parts) + 1)])
        return out

    pl, pr = pack.pack_sets(suffix_items(500), suffix_items(620))
    kept = 0
    for thr in (0.4376, 0.5, 0.5001, 0.5625, 0.6, 0.75, 0.875, 0.9375, 0.95):
        out, info = check_against_oracle(engine, pl, pr, thr)
        kept += len(out)
    assert kept > 0


@pytest.mark.parametrize("vocab,max_k,per_part,nested", [(90, 9, 2, True), (400, 10, 3, True), (3000, 6, 8, True),
                                                          (20000, 11, 2, True), (500, 8, 3, False)])
def test_jaccard_coarse_first_half_vs_oracle(engine, vocab, max_k, per_part, nested):
    """Low thresholds (depth D >= 3) over nested levels that all lie in the slots run stage B with
    the COARSE first half (jaccard.cu, SPLIT == 0) and the stage A tiles drawn from a shared
    counter; deep schedules (K up to 11: one level beyond the slots' ten steps), dense and sparse
    vocabularies and a cohort whose levels are NOT nested (coarse half off) must match the oracle."""
    rng = np.random.default_rng(3 * vocab + max_k)

    def items(n):
        out = []
        for _ in range(n):
            parts = [[f"w{int(x)}" for x in rng.zipf(1.15, size=int(rng.integers(1, per_part + 1))) % vocab]
                     for _ in range(int(rng.integers(1, max_k + 1)))]
            levels = [sorted({w for part in parts[-j:] for w in part}) for j in range(1, len(parts) + 1)]
            if not nested and len(levels) > 2:
                levels[1] = sorted(set(levels[1]) ^ {f"w{int(rng.integers(0, vocab))}", levels[0][0]}) or levels[1]
            out.append(levels)
        return out

    pl, pr = pack.pack_sets(items(700), items(640))
    assert pl.nested == nested and max(pl.max_levels, pr.max_levels) <= min(pl.n_slots, pr.n_slots) + 1
    kept = 0
    for thr in (0.01, 0.03, 0.0625, 0.1, 0.12, 0.2, 0.24):
        out, info = check_against_oracle(engine, pl, pr, thr)
        kept += len(out)
    assert kept > 1000


def test_pipelined_row_blocks_give_the_same_records(engine):
    """Results that go to the host are produced as a probe block plus row blocks sized by the
    probe's density (Engine._run_jobs), so that the copy-out overlaps the remaining kernels.  The
    records must be those of the single launch, for dense and for sparse results."""
    lens, flat = syn.token_id_level_sets(20000, 31)
    pl = pack.pack_suffix_id_sets(lens, flat, 30000)
    lens, flat = syn.token_id_level_sets(3000, 32)
    pr = pack.pack_suffix_id_sets(lens, flat, 30000)
    dl, dr = engine.upload(pl), engine.upload(pr)
    key = lambda a: a[np.lexsort((a["right"], a["left"]))]
    engine.PIPELINE_MIN_PAIRS = 1     # instance overrides of the class constants, dropped below
    engine.direct_host = False        # the device-arena pipeline (Engine._run_jobs)
    try:
        for thr in (0.1, 0.7):
            engine.pipeline_d2h = False
            want = key(engine.all_pairs(dl, dr, thr))
            assert engine.last_info["blocks"] == 1
            engine.pipeline_d2h = True
            for block_bytes in (1 << 20, 512 << 20):
                engine.PIPELINE_BLOCK_BYTES = block_bytes
                got = key(engine.all_pairs(dl, dr, thr))
                assert engine.last_info["blocks"] >= 2 and engine.last_info["count"] == len(want)
                assert np.array_equal(got, want)
            # a row range of the left cohort goes through the same path
            engine.PIPELINE_BLOCK_BYTES = 1 << 20
            part = key(engine.all_pairs(dl, dr, thr, rows=(1000, 19000)))
            assert np.array_equal(part, want[(want["left"] >= 1000) & (want["left"] < 19000)])
    finally:
        engine.pipeline_d2h = engine.direct_host = True
        for name in ("PIPELINE_BLOCK_BYTES", "PIPELINE_MIN_PAIRS"):
            engine.__dict__.pop(name, None)


def _keyed(rec, n_right):
    key = rec["left"].astype(np.uint64) * np.uint64(n_right) + rec["right"]
    order = np.argsort(key, kind="stable")
    return key[order], rec["score"][order]


def check_term_threshold_regimes(engine, n_items, block):
    """cfg5 shape (Term items, K 2-4) at thr 0.5: stage A asks for TWO shared signature bits.
    Properties that do not depend on size: the kept set is the part of the thr-0.45 result (where
    nearly every item is "wild" and stage A falls back to one shared bit) with score >= 0.5; every
    kept score is the oracle's; on a sub-block the oracle's full enumeration finds the same pairs."""
    L = syn.term_level_sets(n_items, syn.SEED_LEFT)
    R = syn.term_level_sets(n_items, syn.SEED_RIGHT)
    rank = pack.frequency_rank([L[1], R[1]], 20000)
    pl, pr = pack.pack_part_id_sets(*L, 20000, rank), pack.pack_part_id_sets(*R, 20000, rank)
    dl, dr = engine.upload(pl), engine.upload(pr)
    hi = engine.all_pairs(dl, dr, 0.5)
    assert engine.last_info["stats"]["bound_pairs"] < 0.55 * n_items * n_items   # the TWO filter is on
    lo = engine.all_pairs(dl, dr, 0.45)
    n_bound_lo = engine.last_info["stats"]["bound_pairs"]
    assert n_bound_lo > 0.6 * n_items * n_items                                    # one-bit regime
    khi, shi = _keyed(hi, pr.n_items)
    klo, slo = _keyed(lo, pr.n_items)
    assert len(khi) == len(np.unique(khi)) and len(klo) == len(np.unique(klo))
    sel = slo >= 0.5
    assert np.array_equal(klo[sel], khi)
    assert np.array_equal(slo[sel].view(np.uint64), shi.view(np.uint64))
    want, _ = c_oracle.score_pairs(pl, pr, lo["left"], lo["right"])
    assert np.array_equal(want.view(np.uint64), lo["score"].view(np.uint64))
    # full enumeration of a sub-block by the oracle
    b0, b1 = block
    sub, _ = c_oracle.all_pairs(pl, pr.rows(0, b1 - b0), 0.45, l_begin=b0, l_end=b1)
    in_block = (lo["left"] >= b0) & (lo["left"] < b1) & (lo["right"] < b1 - b0)
    assert_same_triples((lo["left"][in_block], lo["right"][in_block], lo["score"][in_block]),
                        (sub["left"], sub["right"], sub["score"]))
    return len(hi), len(lo)


def test_term_full_size_threshold_regimes(engine):
    n_hi, n_lo = check_term_threshold_regimes(engine, 200_000, (150_000, 156_000))
    assert 0 < n_hi < n_lo


def check_fuzzy_flat_properties(engine, n_items, thr=0.7):
    """cfg3 shape (flat fuzzy_match on ~60-character strings): kept scores equal the oracle's on a
    sample, a sample of pairs that were not kept is below the threshold, swapping the sides keeps
    the transposed set with the same scores."""
    from napkon_string_matching.text.process import default_process

    vocab = syn.vocabulary()
    sl = [[default_process(s)] for s in syn.question_strings(n_items, syn.SEED_LEFT, vocab)]
    sr = [[default_process(s)] for s in syn.question_strings(n_items + 77, syn.SEED_RIGHT, vocab)]
    pl, pr = pack.pack_strings(sl, sr)
    dl, dr = engine.upload(pl), engine.upload(pr)
    rec = engine.all_pairs(dl, dr, thr, flat=True)
    assert engine.last_info["count"] == len(rec) > 0
    key, score = _keyed(rec, len(sr))
    assert len(np.unique(key)) == len(key)
    swapped = engine.all_pairs(dr, dl, thr, flat=True)
    skey = swapped["right"].astype(np.uint64) * np.uint64(len(sr)) + swapped["left"]
    sorder = np.argsort(skey, kind="stable")
    assert np.array_equal(skey[sorder], key)
    assert np.array_equal(swapped["score"][sorder].view(np.uint64), score.view(np.uint64))
    # records and c_oracle.score_pairs both speak the caller's item indices
    rng = np.random.default_rng(3)
    pick = rng.choice(len(rec), size=min(len(rec), 20_000), replace=False)
    want, _ = c_oracle.score_pairs(pl, pr, rec["left"][pick], rec["right"][pick], flat=True)
    assert np.array_equal(want.view(np.uint64), rec["score"][pick].view(np.uint64))
    li = rng.integers(0, len(sl), size=100_000).astype(np.uint32)
    ri = rng.integers(0, len(sr), size=100_000).astype(np.uint32)
    probe = li.astype(np.uint64) * np.uint64(len(sr)) + ri
    pos = np.searchsorted(key, probe)
    kept = (pos < len(key)) & (key[np.minimum(pos, len(key) - 1)] == probe)
    got, _ = c_oracle.score_pairs(pl, pr, li, ri, flat=True)
    assert np.array_equal(got >= thr, kept)
    return len(rec)


def test_fuzzy_flat_20k_properties(engine):
    assert check_fuzzy_flat_properties(engine, 20_000) > 50_000


def test_jaccard_two_bit_filter_with_folded_shared_ids(engine):
    """The crafted pair of tests/test_filter_soundness.py — its two shared step-1 ids fold onto
    one signature bit — must come out of the kernel with the oracle's score at thr 0.5."""
    import test_filter_soundness as fs

    pl, pr = fs.folded_pair_packs()
    out, info = check_against_oracle(engine, pl, pr, 0.5)
    assert len(out) == 1 and info["flags"] == 0


def test_definitions_1m_x_20k_full_size(engine):
    """BASELINE configs[3] at full size: 1M cohort items x 20k GECCO/KDS-style definitions
    (Term-shaped token sets, K <= 3), thr 0.5, packed on the device.  Size-independent checks:
    every kept score is the oracle's; the thr-0.5 set is the thr-0.45 set cut at 0.5; two
    sub-blocks (the first rows and rows around the middle) are enumerated completely by the
    oracle."""
    from napkon_string_matching.gpu import device_pack as dp

    L = syn.term_level_sets(1_000_000, syn.SEED_LEFT)
    R = syn.definition_level_sets(20_000, syn.SEED_DEFS)
    dl, dr = engine.device_packer.pack([dp.raw_from_parts(*L), dp.raw_from_parts(*R)], 20000, rank="frequency")
    hi = engine.all_pairs(dl, dr, 0.5)
    assert engine.last_info["flags"] == 0 and len(hi) > 1000
    lo = engine.all_pairs(dl, dr, 0.45)
    khi, shi = _keyed(hi, dr.n_items)
    klo, slo = _keyed(lo, dr.n_items)
    assert len(khi) == len(np.unique(khi))
    sel = slo >= 0.5
    assert np.array_equal(klo[sel], khi) and np.array_equal(slo[sel].view(np.uint64), shi.view(np.uint64))
    # the oracle works on host packs with the SAME id ranking the device used
    rank = engine.device_packer.last_rank
    pl, pr = pack.pack_part_id_sets(*L, 20000, rank), pack.pack_part_id_sets(*R, 20000, rank)
    want, _ = c_oracle.score_pairs(pl, pr, lo["left"], lo["right"])
    assert np.array_equal(want.view(np.uint64), lo["score"].view(np.uint64))
    for b0, b1 in ((0, 3000), (499_000, 502_000)):
        sub, _ = c_oracle.all_pairs(pl, pr, 0.45, l_begin=b0, l_end=b1)
        in_block = (lo["left"] >= b0) & (lo["left"] < b1)
        assert_same_triples((lo["left"][in_block], lo["right"][in_block], lo["score"][in_block]),
                            (sub["left"], sub["right"], sub["score"]))


def test_term_1m_x_1m_sampled_properties(engine):
    """BASELINE configs[4] at full size (1M x 1M Term items, thr 0.5; about 6 s of kernel time):
    kept scores equal the oracle's, no pair is kept twice, a random sample of 2e5 pairs is kept
    exactly when the oracle scores it >= 0.5, and a 2000 x 2000 sub-block is enumerated by the
    oracle.  Skipped when the GPU has less than 40 GB free."""
    import time

    import torch

    from napkon_string_matching.gpu import device_pack as dp

    if torch.cuda.mem_get_info()[0] < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    n = 1_000_000
    L, R = syn.term_level_sets(n, syn.SEED_LEFT), syn.term_level_sets(n, syn.SEED_RIGHT)
    dl, dr = engine.device_packer.pack([dp.raw_from_parts(*L), dp.raw_from_parts(*R)], 20000, rank="frequency")
    t0 = time.perf_counter()
    rec = engine.all_pairs(dl, dr, 0.5)
    assert time.perf_counter() - t0 < 60, "1M x 1M took more than a minute"
    assert engine.last_info["flags"] == 0 and engine.last_info["reruns"] == 0
    key, score = _keyed(rec, n)
    assert len(key) == len(np.unique(key)) and len(key) > 100_000
    rank = engine.device_packer.last_rank
    pl, pr = pack.pack_part_id_sets(*L, 20000, rank), pack.pack_part_id_sets(*R, 20000, rank)
    want, _ = c_oracle.score_pairs(pl, pr, rec["left"], rec["right"])
    assert np.array_equal(want.view(np.uint64), rec["score"].view(np.uint64))
    rng = np.random.default_rng(5)
    li = rng.integers(0, n, size=200_000).astype(np.uint32)
    ri = rng.integers(0, n, size=200_000).astype(np.uint32)
    # random pairs almost never reach 0.5: add pairs around kept ones (same left, neighbouring right)
    li = np.concatenate([li, rec["left"][:50_000], rec["left"][:50_000]])
    ri = np.concatenate([ri, rec["right"][:50_000], (rec["right"][:50_000] + 1) % n]).astype(np.uint32)
    probe = li.astype(np.uint64) * np.uint64(n) + ri
    pos = np.searchsorted(key, probe)
    kept = (pos < len(key)) & (key[np.minimum(pos, len(key) - 1)] == probe)
    got, _ = c_oracle.score_pairs(pl, pr, li, ri)
    assert np.array_equal(got >= 0.5, kept)
    b0 = 700_000
    sub, _ = c_oracle.all_pairs(pl, pr.rows(0, 2000), 0.5, l_begin=b0, l_end=b0 + 2000)
    in_block = (rec["left"] >= b0) & (rec["left"] < b0 + 2000) & (rec["right"] < 2000)
    assert_same_triples((rec["left"][in_block], rec["right"][in_block], rec["score"][in_block]),
                        (sub["left"], sub["right"], sub["score"]))
