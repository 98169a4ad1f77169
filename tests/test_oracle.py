"""CPU suite: the oracle (oracle/) against the reference-generated golden vectors, and the
packers against the oracle.  Nothing here touches the GPU."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, assert_same_triples, load_golden
from oracle import c_oracle, reference_port as port
from napkon_string_matching.gpu import pack
from napkon_string_matching.text.tokenize import gen_comp_value, stop_words, word_tokenize

SCALARS = json.loads((GOLDEN / "scalar_cases.json").read_text(encoding="utf-8"))


def _check_scalar(fn, case):
    want = case["result"]
    if want.startswith("raises:"):
        with pytest.raises(Exception) as ei:
            fn(case["left"], case["right"])
        assert type(ei.value).__name__ == want.split(":")[1]
    else:
        assert float(fn(case["left"], case["right"])).hex() == want


@pytest.mark.parametrize("case", SCALARS["jaccard_hand"]["cases"])
def test_port_intersection_vs_union(case):
    _check_scalar(port.intersection_vs_union, case)


@pytest.mark.parametrize("case", SCALARS["fuzzy_hand"]["cases"])
def test_port_fuzzy_match(case):
    _check_scalar(port.fuzzy_match, case)


def test_qratio_known_answers_from_rapidfuzz_docs():
    # fuzz.ratio("this is a test", "this is a test!") == 96.55172413793103 (no processor)
    a, b = "this is a test", "this is a test!"
    lcs = port.lcs_length(a, b)
    assert (1.0 - (len(a) + len(b) - 2 * lcs) / (len(a) + len(b))) * 100 == 96.55172413793103
    # QRatio 2.x processes first: '!' -> ' ' -> stripped -> identical -> 100
    assert port.fuzzy_match(a, b) == 1.0
    assert port.fuzzy_match("", "") == 0.0 and port.fuzzy_match("abc", "") == 0.0
    # Indel, not Levenshtein: a substitution costs 2
    assert port.fuzzy_match("abcd", "abxd") == ((1.0 - 2 / 8) * 100) / 100


@pytest.mark.parametrize("func", ["jaccard", "fuzzy"])
def test_port_compare_terms(func):
    fn = port.intersection_vs_union if func == "jaccard" else port.fuzzy_match
    for case in SCALARS[f"compare_terms_{func}"]["cases"]:
        _check_scalar(lambda l, r: port.compare_terms(l, r, fn), case)


def test_gen_comp_value_matches_reference():
    for case in SCALARS["gen_comp_value"]["cases"]:
        assert gen_comp_value(case["value"]) == case["result"]
        assert port.gen_comp_value(case["value"], word_tokenize, stop_words()) == case["result"]


def _levels(frame, column):
    return [gen_comp_value(v) if v is not None else None for v in frame[column]]


def _golden_levels(name, column):
    meta, inputs, arrays = load_golden(name)
    L, R = _levels(inputs["left"], column), _levels(inputs["right"], column)
    lkeep = [i for i, v in enumerate(L) if v is not None]
    rkeep = [i for i, v in enumerate(R) if v is not None]
    return meta, arrays, [L[i] for i in lkeep], [R[i] for i in rkeep], np.array(lkeep), np.array(rkeep)


@pytest.mark.parametrize("name,column", [("cfg1_400_term_jaccard", "Term"),
                                         ("cfg2_300_tokenids_jaccard", "TokenIds"),
                                         ("variable_80_jaccard", "Variable")])
def test_c_oracle_jaccard_matches_reference_run(name, column):
    meta, arrays, L, R, lkeep, rkeep = _golden_levels(name, column)
    pl, pr = pack.pack_sets(L, R)
    out, flags = c_oracle.all_pairs(pl, pr, meta["kwargs"]["score_threshold"])
    assert flags == 0
    assert_same_triples((lkeep[out["left"]], rkeep[out["right"]], out["score"]),
                        (arrays["left_pos"], arrays["right_pos"], arrays["score"]))


@pytest.mark.parametrize("name,column", [("fuzzy_150_term", "Term"),
                                         ("fuzzy_60_question_str", "Question")])
def test_c_oracle_fuzzy_matches_shimmed_reference_run(name, column):
    meta, arrays, L, R, lkeep, rkeep = _golden_levels(name, column)
    pl, pr = pack.pack_strings(pack.fuzzy_level_strings(L), pack.fuzzy_level_strings(R))
    out, flags = c_oracle.all_pairs(pl, pr, meta["kwargs"]["score_threshold"])
    assert_same_triples((lkeep[out["left"]], rkeep[out["right"]], out["score"]),
                        (arrays["left_pos"], arrays["right_pos"], arrays["score"]))


def test_python_port_equals_c_oracle_on_a_block():
    meta, arrays, L, R, lkeep, rkeep = _golden_levels("cfg1_400_term_jaccard", "Term")
    L, R = L[:60], R[:70]
    want = port.all_pairs(L, R, "intersection_vs_union", 0.1)
    pl, pr = pack.pack_sets(L, R)
    out, _ = c_oracle.all_pairs(pl, pr, 0.1)
    assert_same_triples((out["left"], out["right"], out["score"]), tuple(zip(*want)))
    wantf = port.all_pairs(L[:25], R[:25], "fuzzy_match", 0.3)
    ql, qr = pack.pack_strings(pack.fuzzy_level_strings(L[:25]), pack.fuzzy_level_strings(R[:25]))
    outf, _ = c_oracle.all_pairs(ql, qr, 0.3)
    assert_same_triples((outf["left"], outf["right"], outf["score"]), tuple(zip(*wantf)))


def test_oracle_flags_reference_exceptions():
    pl, pr = pack.pack_sets([[["a"], []]], [[["b"], []]])
    _, flags = c_oracle.all_pairs(pl, pr, 0.0)
    assert flags & c_oracle.FLAG_ZERO_UNION          # ZeroDivisionError in the reference
    pl, pr = pack.pack_sets([[]], [[["a"]]])
    _, flags = c_oracle.all_pairs(pl, pr, 0.0)
    assert flags & c_oracle.FLAG_INDEX_ERROR         # IndexError in the reference
    pl, pr = pack.pack_sets([[]], [[]])
    out, flags = c_oracle.all_pairs(pl, pr, 0.0)
    assert flags == 0 and len(out) == 1 and out["score"][0] == 0.0


def test_pack_sets_layout_and_signatures():
    L = [[["b", "a", "a"], ["a", "b", "c"]], [], [["z"]]]
    R = [[["c"], []]]
    pl, pr = pack.pack_sets(L, R)
    assert pl.n_items == 3 and pl.n_levels == 3 and pl.max_levels == 2
    assert list(pl.item_level_off) == [0, 2, 2, 3]
    assert list(pl.level_sizes()) == [2, 3, 1]
    assert pr.n_levels == 2 and list(pr.level_sizes()) == [1, 0]
    assert pl.exact_bits and pl.n_vocab == 4
    for p in (pl, pr):
        for g in range(p.n_levels):
            toks = p.tok[p.level_tok_off[g]:p.level_tok_off[g + 1]]
            assert list(toks) == sorted(set(toks))
            assert int(p.level_head[g]) == sum(1 << int(t) for t in toks) and p.level_tail[g] == 0
            assert p.level_info[g] & 0xFFFF == len(toks) and p.level_info[g] >> 16 == 0
    # ids are ranked by frequency over both sides: "a" (3 uses incl. duplicates) gets id 0
    assert pl.tok[pl.level_tok_off[0]] == 0
    # item_any = OR over the levels compare_terms can use (1..K-1, or 0 when K == 1)
    assert int(pl.item_any[0, 0]) == int(pl.level_head[1]) and int(pl.item_any[1, 0]) == 0
    assert int(pl.item_any[2, 0]) == int(pl.level_head[2])
    assert int(pr.item_any[0, 0]) == 0  # K == 2: only level 1, which is empty


def test_pack_hashed_signature_invariants():
    rng = np.random.default_rng(5)
    items = [[[f"t{int(x)}" for x in rng.zipf(1.2, size=int(rng.integers(0, 40)))]
              for _ in range(int(rng.integers(1, 5)))] for _ in range(300)]
    (p,) = pack.pack_sets(items)
    assert not p.exact_bits and p.n_vocab > 128
    sizes = p.level_sizes()
    n_head, n_tail_bits = np.bitwise_count(p.level_head), np.bitwise_count(p.level_tail)
    assert np.array_equal((p.level_info >> 16) & 0xFF, np.minimum(sizes - n_head - n_tail_bits, 255))
    assert np.array_equal(p.level_info & 0xFFFF, sizes)
    for g in range(p.n_levels):
        toks = p.tok[p.level_tok_off[g]:p.level_tok_off[g + 1]]
        assert int(p.level_head[g]) == sum(1 << int(t) for t in toks if t < 64)
        assert (p.level_tail[g] == 0) == (not any(t >= 64 for t in toks))
    # head ids are the most frequent ones
    counts = np.bincount(p.tok, minlength=p.n_vocab)
    assert counts[:64].min() >= counts[64:].max()
    k = p.levels_per_item()
    for i in np.nonzero(k >= 2)[0][:50]:
        g0 = int(p.item_level_off[i])
        want = np.bitwise_or.reduce(p.level_head[g0 + 1:g0 + k[i]])
        assert int(p.item_any[i, 0]) == int(want)


def test_pack_suffix_id_sets_equals_generic_packer():
    from napkon_string_matching import synthetic as syn

    lens, flat = syn.token_id_level_sets(500, 11, n_ids=300)
    fast = pack.pack_suffix_id_sets(lens, flat, 300)
    lists, pos = [], 0
    for n in lens:
        lists.append([f"D{int(v):06d}" for v in flat[pos:pos + n]])
        pos += n
    (slow,) = pack.pack_sets([gen_comp_value(v) for v in lists])
    assert np.array_equal(fast.item_level_off, slow.item_level_off)
    assert np.array_equal(fast.level_tok_off, slow.level_tok_off)
    # ids differ (factorize order) but the sets must be isomorphic: same all-pairs result
    a, _ = c_oracle.all_pairs(fast.rows(0, 120), fast.rows(120, 300), 0.1)
    b, _ = c_oracle.all_pairs(slow.rows(0, 120), slow.rows(120, 300), 0.1)
    assert np.array_equal(a, b)


def test_pack_strings_alphabet_and_rows():
    pl, pr = pack.pack_strings([["abc", ""], ["größe 12"]], [["cab"]])
    assert pl.n_alphabet == pr.n_alphabet == len(set("abcgröße 12"))
    assert list(pl.level_lengths()) == [3, 0, 8] and pl.max_len == 8 and pl.max_levels == 2
    assert list(pl.level_chr_off) == [0, 8, 8] and len(pl.chr) == 16
    alphabet = sorted(set("abcgröße 12"))
    assert "".join(alphabet[c] for c in pl.level_string_codes(2)) == "größe 12"
    sub = pl.rows(1, 2)
    assert sub.n_items == 1 and list(sub.level_lengths()) == [8] and list(sub.level_chr_off) == [0]
    assert "".join(alphabet[c] for c in sub.level_string_codes(0)) == "größe 12"


def test_pack_limits_are_loud():
    with pytest.raises(pack.PackError):       # more code points than one byte can name
        pack.pack_strings([["".join(chr(0x4E00 + i) for i in range(300))]])
    with pytest.raises(pack.PackError):       # a level beyond the 16-bit size field
        pack.pack_sets([[[f"t{i}" for i in range(70000)]]])
    (p,) = pack.pack_strings([["x" * 600], ["y"]])   # longer than the kernel's 8 words: last class
    assert int(p.class_end[-1]) == 1 and p.n_items == 2


def test_reciprocal_fma_division_equals_true_division_for_small_counts():
    """csrc/jaccard.cu:div_counts computes |A & B| / |A | B| as r = RN(1/u); q0 = RN(i*r);
    q = fma(fma(-q0, u, i), r, q0) for u <= 255.  Emulated here with exact rational arithmetic
    (one rounding per fused operation): it must equal Python's correctly rounded i / u, which is
    what score_functions.py:13 returns, for every 0 <= i <= u <= 255."""
    from fractions import Fraction

    for u in range(1, 256):
        r = 1.0 / u
        for i in range(u + 1):
            a, b = float(i), float(u)
            q0 = a * r
            e = float(Fraction(a) - Fraction(q0) * Fraction(b))
            q = float(Fraction(q0) + Fraction(e) * Fraction(r))
            assert q == a / b, (i, u)


def test_fast_ratio_arithmetic_is_exhaustively_exact(tmp_path):
    """tests/csrc/fast_ratio_check.c enumerates every (dist, lensum) the fuzzy kernel's
    division-free QRatio arithmetic can see (and every count pair of the Jaccard kernel's) and
    compares reciprocal + two FMAs with the IEEE division bit for bit."""
    import subprocess

    from conftest import ROOT

    exe = tmp_path / "fast_ratio_check"
    subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-o", str(exe),
                    str(ROOT / "tests" / "csrc" / "fast_ratio_check.c"), "-lm"], check=True)
    n, bad_div, bad_ratio = map(int, subprocess.run([str(exe), "1026"], check=True, capture_output=True,
                                                    text=True).stdout.split())
    assert n > 500_000 and bad_div == 0 and bad_ratio == 0
