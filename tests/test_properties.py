"""CPU suite, property-based: random ragged inputs through the Python port (the reference
restated line by line), the packers and the C oracle must agree bit for bit."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from conftest import assert_same_triples
from oracle import c_oracle, reference_port as port
from napkon_string_matching.gpu import pack
from napkon_string_matching.text.process import default_process, join_sorted

TOKENS = st.sampled_from(["a", "b", "c", "Haus", "haus", "Größe", "x1", "ß", "D000001", "D000002",
                          "und", "_", "é", "Ä", "zz"] + [f"t{i}" for i in range(150)])
LEVEL = st.lists(TOKENS, min_size=0, max_size=9)
ITEM = st.lists(LEVEL, min_size=0, max_size=5)
SIDE = st.lists(ITEM, min_size=0, max_size=7)
COMMON = dict(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])


def port_pairs(L, R, func, thr):
    """Reference semantics pair by pair; pairs on which the reference raises are left out and
    reported as flags, like the kernels do."""
    out, raised = [], set()
    for li, lv in enumerate(L):
        for ri, rv in enumerate(R):
            try:
                s = port.compare_terms(lv, rv, port.SCORE_FUNCS[func])
            except (IndexError, ZeroDivisionError) as e:
                raised.add(type(e).__name__)
                continue
            if s >= thr:
                out.append((li, ri, float(s)))
    return out, raised


@settings(**COMMON)
@given(SIDE, SIDE, st.sampled_from([0.0, 0.05, 0.25, 0.5, 0.75]))
def test_jaccard_port_equals_c_oracle_on_packed(L, R, thr):
    want, raised = port_pairs(L, R, "intersection_vs_union", thr)
    pl, pr = pack.pack_sets(L, R)
    got, flags = c_oracle.all_pairs(pl, pr, thr)
    assert_same_triples((got["left"], got["right"], got["score"]),
                        tuple(zip(*want)) if want else ([], [], []))
    assert bool(flags & c_oracle.FLAG_ZERO_UNION) == ("ZeroDivisionError" in raised)
    assert bool(flags & c_oracle.FLAG_INDEX_ERROR) == ("IndexError" in raised)


@settings(**COMMON)
@given(SIDE, SIDE, st.sampled_from([0.0, 0.3, 0.6]))
def test_fuzzy_port_equals_c_oracle_on_packed(L, R, thr):
    want, raised = port_pairs(L, R, "fuzzy_match", thr)
    pl, pr = pack.pack_strings(pack.fuzzy_level_strings(L), pack.fuzzy_level_strings(R))
    got, flags = c_oracle.all_pairs(pl, pr, thr)
    assert_same_triples((got["left"], got["right"], got["score"]),
                        tuple(zip(*want)) if want else ([], [], []))
    assert bool(flags & c_oracle.FLAG_INDEX_ERROR) == ("IndexError" in raised)


@settings(**COMMON)
@given(SIDE)
def test_pack_sets_roundtrip_and_summaries(items):
    (p,) = pack.pack_sets(items)
    k = p.levels_per_item()
    assert list(k) == [len(it) for it in items] and list(p.item_k) == list(k)
    sizes = p.level_sizes()
    flat_levels = [lv for it in items for lv in it]
    assert list(sizes) == [len(set(lv)) for lv in flat_levels]
    for g in range(p.n_levels):
        toks = p.tok[p.level_tok_off[g]:p.level_tok_off[g + 1]]
        assert np.all(toks[1:] > toks[:-1])
        assert int(p.level_head[g]) == sum(1 << int(t) for t in toks if t < 64)
        assert (int(p.level_info[g]) & 0xFFFF) == len(toks)
    # slots follow compare_terms' schedule: slot t-1 = level min(t, K-1)
    for i, it in enumerate(items):
        for t in range(1, p.n_slots + 1):
            if len(it) == 0:
                assert p.slot_info[t - 1, i] == 0
            else:
                g = int(p.item_level_off[i]) + min(t, len(it) - 1)
                assert p.slot_ht[t - 1, i, 0] == p.level_head[g]
                assert p.slot_ht[t - 1, i, 1] == p.level_tail[g]
                assert p.slot_info[t - 1, i] == p.level_info[g]
    assert p.slot_stride % 128 == 0 and not p.slot_ht[:, p.n_items:].any()
    # same token string <-> same id across levels and items
    (q,) = pack.pack_sets(items + items)
    assert np.array_equal(q.level_sizes()[:p.n_levels], sizes)


@settings(**COMMON)
@given(st.lists(st.lists(st.text(alphabet="abcÄß 1_-!é", max_size=150), min_size=0, max_size=3),
                min_size=0, max_size=8))
def test_pack_strings_roundtrip(items):
    processed = [[default_process(s) for s in it] for it in items]
    (p,) = pack.pack_strings(processed)
    alphabet = sorted(set("".join(s for it in processed for s in it)))
    assert p.n_alphabet == len(alphabet)
    assert sorted(int(x) for x in p.perm) == list(range(len(items)))
    words = [max(1, -(-max((len(s) for s in processed[i]), default=0) // 64)) for i in p.perm]
    assert words == sorted(words)
    assert int(p.class_end[-1]) == len(items)
    for pos, i in enumerate(p.perm):
        g0, g1 = int(p.item_level_off[pos]), int(p.item_level_off[pos + 1])
        got = ["".join(alphabet[c] for c in p.level_string_codes(g)) for g in range(g0, g1)]
        assert got == processed[i]
        assert all(int(p.level_chr_off[g]) % 8 == 0 for g in range(g0, g1))


@settings(**COMMON)
@given(st.lists(TOKENS, max_size=8), st.lists(TOKENS, max_size=8))
def test_scalar_port_properties(a, b):
    if a or b:
        j = port.intersection_vs_union(a, b)
        assert j == port.intersection_vs_union(b, a) and 0.0 <= j <= 1.0
        assert (j == 1.0) == (set(a) == set(b))
    f = port.fuzzy_match(a, b)
    assert f == port.fuzzy_match(b, a) and 0.0 <= f <= 1.0
    sa = default_process(join_sorted(a))
    assert port.fuzzy_match(a, a) == (1.0 if sa else 0.0)
