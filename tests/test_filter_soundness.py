"""CPU suite: soundness of the Jaccard kernel's stage A filters, restated in numpy.

csrc/jaccard.cu prunes a pair in stage A when the signatures of the first D steps share no bit
(or, in the TWO kernels, fewer than two bits while neither side is "wild").  A pruned pair must be
PROVEN below the threshold.  Here the host-side parameter choice (depth D, two_small) and the
per-pair predicates are restated from the packed arrays, and every pair whose exact score (C
oracle) reaches the threshold must pass them — over random nested items, dense and sparse
vocabularies, tiny and large levels and thresholds around 1/2."""
import numpy as np
import pytest

from oracle import c_oracle
from napkon_string_matching.gpu import pack

U64 = np.uint64


def filter_threshold(thr: float) -> np.float32:          # api.cu:filter_threshold
    if not thr > 0.0:
        return np.float32(-np.inf)
    return np.nextafter(np.float32(thr * (1.0 - 1e-9)), np.float32(-np.inf))


def stage_a_parameters(pl, pr, thr):
    """(any_depth D, two_small or None) as nsm_jaccard_allpairs chooses them."""
    thr_lo = filter_threshold(thr)
    if not thr_lo > 0:
        return 0, None
    kmax = max(pl.max_levels, pr.max_levels)
    w_last = np.float32(2.0 ** -kmax) if kmax < 120 else np.float32(0)
    d = 1
    while d < 64 and not (np.float32(2.0 ** -d) - w_last < thr_lo):
        d += 1
    l_deep, r_deep = pl.max_levels > pl.n_slots + 1, pr.max_levels > pr.n_slots + 1
    depth = min(d, 10) if (d <= pl.n_slots or not l_deep) and (d <= pr.n_slots or not r_deep) else 0
    two = None
    if depth == 1 and d <= 2:
        jmin = 2.0 * (float(thr_lo) - 0.5 + float(w_last))
        if jmin > 0.0:
            t = np.floor((1.0 + 1.0 / jmin) * (1.0 + 1e-9))
            if t < 64.0:
                two = int(t) // 2
    return depth, two


def any_words(p, depth):
    n = p.n_items
    if depth == 0:
        return p.item_any[:, 0].copy(), p.item_any[:, 1].copy()
    h, t = np.zeros(n, U64), np.zeros(n, U64)
    for sl in range(min(depth, p.n_slots)):
        h |= p.slot_ht[sl, :n, 0]
        t |= p.slot_ht[sl, :n, 1]
    return h, t


def two_words(p, small):
    """(head, folded 62-bit tail, wild) of every item's step-1 level: jaccard.cu:two_word."""
    n = p.n_items
    head, tail, info = p.slot_ht[0, :n, 0], p.slot_ht[0, :n, 1], p.slot_info[0, :n]
    low = tail & U64(0x3FFFFFFFFFFFFFFF)
    top = (tail >> U64(62)) << U64(60)
    wild = ((info >> 16) != 0) | ((low & top) != 0) | ((info & 0xFFFF) <= small) | (p.item_k == 0)
    return head, low | top, wild


def popcount(x):
    return np.bitwise_count(x).astype(np.int64)


def suffix_items(rng, n, max_k, per_part, vocab, zipf):
    out = []
    for _ in range(n):
        parts = [[f"w{int(x)}" for x in rng.zipf(zipf, size=int(rng.integers(1, per_part + 1))) % vocab]
                 for _ in range(int(rng.integers(1, max_k + 1)))]
        out.append([sorted({w for part in parts[-j:] for w in part}) for j in range(1, len(parts) + 1)])
    return out


@pytest.mark.parametrize("vocab,max_k,per_part,zipf", [(70, 4, 5, 1.2), (400, 4, 9, 1.2), (3000, 3, 30, 1.3),
                                                        (20000, 4, 6, 1.1), (500, 12, 2, 1.3)])
def test_stage_a_never_prunes_a_pair_that_reaches_the_threshold(vocab, max_k, per_part, zipf):
    rng = np.random.default_rng(vocab + max_k)
    pl, pr = pack.pack_sets(suffix_items(rng, 260, max_k, per_part, vocab, zipf),
                            suffix_items(rng, 300, max_k, per_part, vocab, zipf))
    everything, _ = c_oracle.all_pairs(pl, pr, 0.0)
    score = np.zeros((pl.n_items, pr.n_items))
    score[everything["left"], everything["right"]] = everything["score"]
    used_two = 0
    for thr in (0.05, 0.1, 0.26, 0.4376, 0.5, 0.5001, 0.5625, 0.6, 0.75, 0.9, 0.9375):
        depth, two = stage_a_parameters(pl, pr, thr)
        if two is None:
            (hl, tl), (hr, tr) = any_words(pl, depth), any_words(pr, depth)
            passes = ((hl[:, None] & hr[None, :]) | (tl[:, None] & tr[None, :])) != 0
        else:
            used_two += 1
            (hl, tl, wl), (hr, tr, wr) = two_words(pl, two), two_words(pr, two)
            z, w = hl[:, None] & hr[None, :], tl[:, None] & tr[None, :]
            bits = popcount(z) + popcount(w)
            passes = (bits >= 1) & ((bits >= 2) | wl[:, None] | wr[None, :])
        reaches = score >= thr
        assert not np.any(reaches & ~passes), (thr, depth, two, np.argwhere(reaches & ~passes)[:3])
        assert reaches.sum() <= passes.sum()
    assert used_two >= 3   # the thresholds around 1/2 exercised the two-bit regime


def folded_pair_packs():
    """(left, right): one K = 4 item each whose step-1 levels share exactly two tail ids that hash
    to the SAME signature bit (plus fillers on other bits), and whose deeper levels overlap so
    much that compare_terms reaches 0.5."""
    n_vocab = 5000
    ids = np.arange(pack.HEAD_IDS, n_vocab)
    bit = pack.tail_bits(ids, exact_bits=False)
    order = np.argsort(bit, kind="stable")
    same = np.nonzero(bit[order][1:] == bit[order][:-1])[0][0]
    p, q = int(ids[order][same]), int(ids[order][same + 1])           # two tail ids on one bit
    free = [int(i) for i in ids if bit[i - pack.HEAD_IDS] != bit[p - pack.HEAD_IDS] and i not in (p, q)]
    seen, fill = set(), []
    for i in free:                                                    # six fillers on six other bits
        b = int(bit[i - pack.HEAD_IDS])
        if b not in seen:
            seen.add(b); fill.append(i)
        if len(fill) == 6:
            break
    shared = list(range(0, 50)) + free[100:150]                        # enters at level 2 only

    def item(first, fillers):
        l1 = sorted({p, q, *fillers})
        l2 = sorted({*l1, *shared})
        return [[first], l1, l2, l2]

    def packed(levels):
        sizes = [len(l) for l in levels]
        return pack.finish_sets(np.array([0, len(levels)]), np.concatenate([[0], np.cumsum(sizes)]),
                                np.concatenate(levels), n_vocab)

    return packed(item(p, fill[:3])), packed(item(q, fill[3:]))


def test_two_bit_filter_keeps_a_pair_whose_shared_ids_fold_onto_one_signature_bit():
    """Two shared tail ids that hash to the SAME signature bit show up as one shared bit.  The
    items are then "wild" (fold count > 0) and must be tested for one shared bit only — a pair
    built to reach the threshold exactly this way has to pass stage A."""
    pl, pr = folded_pair_packs()
    out, _ = c_oracle.all_pairs(pl, pr, 0.5)
    assert len(out) == 1 and out["score"][0] >= 0.5                    # the pair reaches thr 0.5
    depth, two = stage_a_parameters(pl, pr, 0.5)
    assert depth == 1 and two == 4
    (hl, tl, wl), (hr, tr, wr) = two_words(pl, two), two_words(pr, two)
    bits = popcount(hl & hr) + popcount(tl & tr)
    assert bits[0] == 1 and wl[0] and wr[0]      # one shared bit, and only the fold flag saves it
    assert (pl.slot_info[0, 0] >> 16) == 1 and (pl.slot_info[0, 0] & 0xFFFF) == 5 > two


def stage_b_bound(pl, pr, unroll=6):
    """Real-valued version of the kernel's stage B: upper bound of compare_terms' score of every
    pair from the slot summaries (jaccard.cu:bound_intersection / bound_step; the kernel evaluates
    the same expressions in fp32 with every operation rounded up)."""
    nl, nr = pl.n_items, pr.n_items
    kmax = np.maximum(pl.item_k[:, None].astype(np.int64), pr.item_k[None, :].astype(np.int64))
    exact = pl.exact_bits and pr.exact_bits
    ub = np.zeros((nl, nr))
    for t in range(1, unroll + 1):
        sl, sr = min(t, pl.n_slots) - 1, min(t, pr.n_slots) - 1
        ha, ta, ia = pl.slot_ht[sl, :nl, 0], pl.slot_ht[sl, :nl, 1], pl.slot_info[sl, :nl].astype(np.int64)
        hb, tb, ib = pr.slot_ht[sr, :nr, 0], pr.slot_ht[sr, :nr, 1], pr.slot_info[sr, :nr].astype(np.int64)
        both = ta[:, None] & tb[None, :]
        fold = np.minimum(ia[:, None] >> 16, ib[None, :] >> 16)
        extra = np.where(fold == 255, 1 << 16, fold) if not exact else 0
        it = np.where(both != 0, popcount(both) + extra, 0)
        a, b = (ia & 0xFFFF)[:, None], (ib & 0xFFFF)[None, :]
        ih = np.minimum(popcount(ha[:, None] & hb[None, :]) + it, np.minimum(a, b))
        union = a + b - ih
        ub += np.where(t <= kmax, np.where(union > 0, ih / np.maximum(union, 1), 0.0) * 2.0 ** -t, 0.0)
    return ub + np.where(kmax > unroll, 2.0 ** -unroll - 2.0 ** -kmax.astype(float), 0.0)


@pytest.mark.parametrize("vocab,max_k,per_part,zipf", [(70, 4, 5, 1.2), (400, 4, 9, 1.2), (3000, 3, 30, 1.3),
                                                        (20000, 4, 6, 1.1), (500, 9, 2, 1.3)])
def test_stage_b_bound_is_an_upper_bound_of_the_exact_score(vocab, max_k, per_part, zipf):
    rng = np.random.default_rng(7 * vocab + max_k)
    pl, pr = pack.pack_sets(suffix_items(rng, 220, max_k, per_part, vocab, zipf),
                            suffix_items(rng, 260, max_k, per_part, vocab, zipf))
    assert max(pl.max_levels, pr.max_levels) <= min(pl.n_slots, pr.n_slots) + 1   # slots hold every level
    everything, _ = c_oracle.all_pairs(pl, pr, 0.0)
    score = np.zeros((pl.n_items, pr.n_items))
    score[everything["left"], everything["right"]] = everything["score"]
    ub = stage_b_bound(pl, pr)
    assert np.all(ub >= score - 1e-12), np.argwhere(ub < score - 1e-12)[:3]
    # and it is a useful bound, not a trivial one: it separates most pairs from a threshold of 0.3
    assert (ub < 0.3).mean() > 0.3


def coarse_first_half(pl, pr, thr, unroll=6):
    """Stage B's coarse first half (jaccard.cu, SPLIT == 0) restated in real arithmetic: a pair goes
    on to the per-step bound when it shares a signature bit at step 2, or when
    (2^-2 - 2^-min(T, kmax)) * min(1, ih_T / umin) + [kmax > T](2^-T - 2^-kmax) reaches the
    threshold, ih_T being the intersection bound of step T's summaries and umin = a_3 + b_3 -
    min(ih_T, a_3, b_3) a lower bound of every union from step 3 on."""
    nl, nr = pl.n_items, pr.n_items
    slot = lambda t, s: min(t, s) - 1
    kmax = np.maximum(pl.item_k[:, None].astype(np.int64), pr.item_k[None, :].astype(np.int64))
    a2, b2 = pl.slot_ht[slot(2, pl.n_slots), :nl], pr.slot_ht[slot(2, pr.n_slots), :nr]
    share2 = ((a2[:, None, 0] & b2[None, :, 0]) | (a2[:, None, 1] & b2[None, :, 1])) != 0
    sl, sr = slot(unroll, pl.n_slots), slot(unroll, pr.n_slots)
    ha, ta, ia = pl.slot_ht[sl, :nl, 0], pl.slot_ht[sl, :nl, 1], pl.slot_info[sl, :nl].astype(np.int64)
    hb, tb, ib = pr.slot_ht[sr, :nr, 0], pr.slot_ht[sr, :nr, 1], pr.slot_info[sr, :nr].astype(np.int64)
    both = ta[:, None] & tb[None, :]
    fold = np.minimum(ia[:, None] >> 16, ib[None, :] >> 16)
    extra = np.where(fold == 255, 1 << 16, fold) if not (pl.exact_bits and pr.exact_bits) else 0
    it = np.where(both != 0, popcount(both) + extra, 0)
    ih = np.minimum(popcount(ha[:, None] & hb[None, :]) + it, np.minimum(ia & 0xFFFF, 1 << 30)[:, None])
    ih = np.minimum(ih, (ib & 0xFFFF)[None, :])
    a3 = (pl.slot_info[slot(3, pl.n_slots), :nl] & 0xFFFF).astype(np.int64)[:, None]
    b3 = (pr.slot_info[slot(3, pr.n_slots), :nr] & 0xFFFF).astype(np.int64)[None, :]
    umin = a3 + b3 - np.minimum(ih, np.minimum(a3, b3))
    jc = np.where(umin == 0, 1.0, np.minimum(1.0, ih / np.maximum(umin, 1)))
    wsum = np.maximum(0.0, 0.25 - 2.0 ** -np.minimum(unroll, kmax).astype(float))
    grant = np.where(kmax > unroll, 2.0 ** -unroll - 2.0 ** -kmax.astype(float), 0.0)
    return share2 | (jc * wsum + grant >= float(filter_threshold(thr)))


@pytest.mark.parametrize("vocab,max_k,per_part,zipf", [(70, 4, 5, 1.2), (400, 5, 9, 1.2), (3000, 9, 3, 1.3),
                                                        (20000, 8, 2, 1.1), (500, 9, 2, 1.3), (150, 10, 1, 1.05)])
def test_coarse_first_half_never_prunes_a_pair_that_reaches_the_threshold(vocab, max_k, per_part, zipf):
    rng = np.random.default_rng(11 * vocab + max_k)
    pl, pr = pack.pack_sets(suffix_items(rng, 240, max_k, per_part, vocab, zipf),
                            suffix_items(rng, 280, max_k, per_part, vocab, zipf))
    assert pl.nested and pr.nested
    assert max(pl.max_levels, pr.max_levels) <= min(pl.n_slots, pr.n_slots) + 1   # the kernel's precondition
    everything, _ = c_oracle.all_pairs(pl, pr, 0.0)
    score = np.zeros((pl.n_items, pr.n_items))
    score[everything["left"], everything["right"]] = everything["score"]
    pruned_some = False
    for thr in (0.01, 0.03, 0.0625, 0.1, 0.12, 0.2, 0.24):
        depth, _ = stage_a_parameters(pl, pr, thr)
        passes = coarse_first_half(pl, pr, thr)
        reaches = score >= thr
        assert not np.any(reaches & ~passes), (thr, np.argwhere(reaches & ~passes)[:3])
        pruned_some |= bool(np.any(~passes & (score > 0)))
    assert pruned_some     # and it does prune pairs with a non-zero score
