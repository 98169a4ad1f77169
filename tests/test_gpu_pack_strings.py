"""GPU suite: device-side string packing (csrc/pack_strings.cu through the C ABI) against the host
packer — ``pack.pack_strings(pack.fuzzy_level_strings(...))``, every array bit for bit — and the
fuzzy kernels run on device-packed cohorts against the oracle."""
import numpy as np
import pytest

from conftest import assert_same_triples
from oracle import c_oracle
from napkon_string_matching import synthetic as syn
from napkon_string_matching.gpu import device_pack as dp
from napkon_string_matching.gpu import pack
from napkon_string_matching.text.process import default_process

pytestmark = pytest.mark.gpu


def assert_same_strings(cohort, want: pack.PackedStrings):
    got = dp.strings_to_host(cohort)
    assert got.n_items == want.n_items and got.n_levels == want.n_levels
    assert (got.n_alphabet, got.max_levels, got.max_len) == (want.n_alphabet, want.max_levels, want.max_len)
    assert np.array_equal(got.perm, want.perm) and np.array_equal(got.classes(), want.classes())
    for f in ("item_level_off", "level_chr_off", "level_len", "chr"):
        a, b = getattr(got, f), getattr(want, f)
        assert a.shape == b.shape, (f, a.shape, b.shape)
        assert np.array_equal(a, b), (f, np.argwhere(a != b)[:5].tolist())
    want.arrays()   # builds level_hist
    assert np.array_equal(got.level_hist, want.level_hist)
    assert np.array_equal(cohort.weights, np.diff(np.concatenate([[0], np.cumsum(want.level_lengths())])[
        want.item_level_off.astype(np.int64)]).astype(np.float64) + 1.0)


def noisy(rng, words, n, levels=(1, 1)):
    """Level strings as join_sorted leaves them: mixed case, punctuation, blanks at the ends."""
    junk = [" ", "  ", "-", "?", "! ", "(", ")", "/", "_", "\t", "3", "Ä", "ß", "é", "°", ""]
    out = []
    for _ in range(n):
        lv = []
        for _ in range(int(rng.integers(levels[0], levels[1] + 1))):
            parts = [str(rng.choice(junk))]
            for _ in range(int(rng.integers(0, 9))):
                w = str(rng.choice(words))
                parts += [w.upper() if rng.random() < 0.2 else w, str(rng.choice(junk))]
            lv.append("".join(parts))
        out.append(lv)
    return out


def host_packs(*sides):
    return pack.pack_strings(*[[[default_process(x) for x in lv] for lv in s] for s in sides])


@pytest.mark.parametrize("levels", [(1, 1), (0, 4)])
def test_noisy_strings_equal_host_packer(engine, levels):
    rng = np.random.default_rng(21 + levels[1])
    words = syn.vocabulary(3000)
    left, right = noisy(rng, words, 1500, levels), noisy(rng, words, 1100, levels)
    left[7], right[9] = [" ?! "] * max(1, levels[0]), [""] * max(1, levels[0])     # nothing survives the trim
    left[11] = ["x" * 70 + " " + "y" * 300 + "?"] * max(1, levels[0])               # several word classes
    right[3] = ["Zz " * 200] * max(1, levels[0])                                      # beyond 512 characters
    cohorts = dp.DeviceStringPacker(engine).pack([left, right])
    for cohort, want in zip(cohorts, host_packs(left, right)):
        assert_same_strings(cohort, want)


def test_alphabet_beyond_255_code_points(engine):
    """More than 255 code points in all, fewer in common: every one-sided code point shares a code."""
    rng = np.random.default_rng(5)
    base = [chr(c) for c in range(0x4E00, 0x4E00 + 180)]
    only_l = [chr(c) for c in range(0x5E00, 0x5E00 + 150)]
    only_r = [chr(c) for c in range(0x6E00, 0x6E00 + 150)]
    mk = lambda own, n: [["".join(rng.choice(base + own, size=int(rng.integers(1, 40))))] for _ in range(n)]
    left, right = mk(only_l, 400), mk(only_r, 300)
    cohorts = dp.DeviceStringPacker(engine).pack([left, right])
    for cohort, want in zip(cohorts, host_packs(left, right)):
        assert_same_strings(cohort, want)


def test_position_dependent_lower_case_is_refused(engine):
    with pytest.raises(dp.PackUnsupported):
        dp.DeviceStringPacker(engine).pack([[["ΟΔΟΣ"]], [["οδος"]]])
    with pytest.raises(dp.PackUnsupported):
        dp.DeviceStringPacker(engine).pack([[["İstanbul"]], [["istanbul"]]])


def test_kernels_on_device_packed_strings_match_the_oracle(engine):
    rng = np.random.default_rng(8)
    words = syn.vocabulary(800)
    left, right = noisy(rng, words, 700, (1, 1)), noisy(rng, words, 650, (1, 1))
    dl, dr = dp.DeviceStringPacker(engine).pack([left, right])
    hl, hr = host_packs(left, right)
    got = engine.all_pairs(dl, dr, 0.6, flat=True)
    want, _ = c_oracle.all_pairs(hl, hr, 0.6, flat=True)
    assert len(got) > 50
    assert_same_triples((got["left"], got["right"], got["score"]), (want["left"], want["right"], want["score"]))
