"""GPU suite: device-side token packing (csrc/pack.cu through the C ABI) against the numpy packer
— every array bit for bit — and the comparison kernels run on device-packed cohorts against the
oracle."""
import numpy as np
import pytest

from conftest import assert_same_triples, load_golden
from oracle import c_oracle
from napkon_string_matching import synthetic as syn
from napkon_string_matching.gpu import device_pack as dp
from napkon_string_matching.gpu import lib as nsmlib
from napkon_string_matching.gpu import pack
from napkon_string_matching.text.tokenize import gen_comp_value

pytestmark = pytest.mark.gpu

FIELDS = dp._SETS_FIELDS


def assert_same_pack(cohort, want: pack.PackedSets):
    got = dp.to_host(cohort)
    assert got.n_items == want.n_items and got.n_levels == want.n_levels
    assert got.max_levels == want.max_levels and got.n_slots == want.n_slots
    assert got.exact_bits == want.exact_bits and got.nested == want.nested
    for f in FIELDS:
        a, b = getattr(got, f), getattr(want, f)
        assert a.shape == b.shape, (f, a.shape, b.shape)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)[:5].tolist()
            raise AssertionError(f"{f} differs at {bad}: got {a[tuple(bad[0])]}, want {b[tuple(bad[0])]}")
    # the per-item work estimate the row-block partitioner uses
    csum = np.concatenate([[0], np.cumsum(want.level_sizes())])
    off = want.item_level_off.astype(np.int64)
    assert np.array_equal(cohort.weights, (csum[off[1:]] - csum[off[:-1]]).astype(np.float64) + 1.0)


def host_pack_levels(raw: dp.RawSets, n_vocab: int, rank) -> pack.PackedSets:
    """The numpy packer on explicit levels of integer codes."""
    codes = raw.ids.astype(np.int64)
    if rank is not None:
        codes = np.asarray(rank)[codes]
    new_off, codes = pack._sort_unique_levels(raw.grp_id_off.astype(np.int64), codes)
    return pack.finish_sets(raw.item_grp_off, new_off, codes, n_vocab)


def test_term_parts_equal_host_packer(engine):
    raws = {k: syn.term_level_sets(n, seed) for k, (n, seed) in {"l": (5000, 11), "r": (3777, 12)}.items()}
    rank = pack.frequency_rank([f for _, f in raws.values()], 20000)
    cohorts = engine.device_packer.pack([dp.raw_from_parts(pl, f) for pl, f in raws.values()], 20000)
    assert np.array_equal(engine.device_packer.last_rank, rank)
    for cohort, (pl, f) in zip(cohorts, raws.values()):
        assert_same_pack(cohort, pack.pack_part_id_sets(pl, f, 20000, rank))


@pytest.mark.parametrize("n_vocab,max_len,rank", [(30000, 10, None), (30000, 10, "frequency"),
                                                   (100, 12, None), (128, 40, "frequency"),
                                                   (129, 40, None)])
def test_id_lists_equal_host_packer(engine, n_vocab, max_len, rank):
    lens, flat = syn.token_id_level_sets(4001, 5, n_ids=n_vocab, max_len=max_len)
    host_rank = pack.frequency_rank([flat], n_vocab) if rank else None
    (cohort,) = engine.device_packer.pack([dp.raw_from_id_lists(lens, flat)], n_vocab, rank=rank)
    assert_same_pack(cohort, pack.pack_suffix_id_sets(lens, flat, n_vocab, host_rank))


def _random_levels(rng, n_items, n_vocab, max_k, max_size, nested):
    items = []
    for _ in range(n_items):
        k = int(rng.integers(0, max_k + 1))
        levels, cur = [], []
        for _ in range(k):
            new = rng.integers(0, n_vocab, size=int(rng.integers(0, max_size + 1))).tolist()
            cur = cur + new if nested else new
            lv = list(cur)
            rng.shuffle(lv)
            levels.append(lv + lv[: int(rng.integers(0, 3))])  # duplicates inside a level
        items.append(levels)
    return items


@pytest.mark.parametrize("nested", [True, False])
@pytest.mark.parametrize("n_vocab,max_k,max_size", [(50, 5, 6), (5000, 12, 9), (300, 3, 120), (70, 24, 3)])
def test_explicit_levels_equal_host_packer(engine, nested, n_vocab, max_k, max_size):
    rng = np.random.default_rng(n_vocab + max_k)
    raw = dp.raw_from_levels(_random_levels(rng, 700, n_vocab, max_k, max_size, nested))
    rank = pack.frequency_rank([raw.ids], n_vocab)
    (cohort,) = engine.device_packer.pack([raw], n_vocab, rank=rank)
    want = host_pack_levels(raw, n_vocab, rank)
    assert_same_pack(cohort, want)
    if nested:
        assert cohort.struct.nested == 1


def test_edge_shapes(engine):
    packer = engine.device_packer
    # no items; items without parts; one item holding exactly the per-item limit
    for lens in ([], [0, 0, 0], [3, 0, 2], [nsmlib.PACK_MAX_ITEM_IDS], [1] * 300):
        lens = np.asarray(lens, dtype=np.int64)
        flat = (np.arange(int(lens.sum())) * 7919 % 997).astype(np.uint32)
        (cohort,) = packer.pack([dp.raw_from_id_lists(lens, flat)], 997, rank=None)
        assert_same_pack(cohort, pack.pack_suffix_id_sets(lens, flat, 997))
    with pytest.raises(pack.PackError):
        lens = np.asarray([nsmlib.PACK_MAX_ITEM_IDS + 1])
        packer.pack([dp.raw_from_id_lists(lens, np.zeros(int(lens[0]), np.uint32))], 10, rank=None)
    with pytest.raises(pack.PackError):
        packer.pack([dp.raw_from_id_lists(np.asarray([2]), np.asarray([1, 99], np.uint32))], 10, rank=None)


def test_long_parts_cross_chunk_boundaries(engine):
    rng = np.random.default_rng(3)
    part_lens = rng.integers(0, 90, size=(400, 5))
    flat = rng.integers(0, 700, size=int(part_lens.sum())).astype(np.uint32)
    rank = pack.frequency_rank([flat], 700)
    (cohort,) = engine.device_packer.pack([dp.raw_from_parts(part_lens, flat)], 700)
    assert_same_pack(cohort, pack.pack_part_id_sets(part_lens, flat, 700, rank))


def test_kernels_on_device_packed_cohorts_vs_oracle(engine):
    lens_l, flat_l = syn.token_id_level_sets(900, 1)
    lens_r, flat_r = syn.token_id_level_sets(1100, 2)
    dl, dr = engine.device_packer.pack([dp.raw_from_id_lists(lens_l, flat_l),
                                        dp.raw_from_id_lists(lens_r, flat_r)], 30000, rank=None)
    out = engine.all_pairs(dl, dr, 0.1)
    pl, pr = pack.pack_suffix_id_sets(lens_l, flat_l, 30000), pack.pack_suffix_id_sets(lens_r, flat_r, 30000)
    want, _ = c_oracle.all_pairs(pl, pr, 0.1)
    assert_same_triples((out["left"], out["right"], out["score"]),
                        (want["left"], want["right"], want["score"]))


@pytest.mark.parametrize("name,column", [("cfg1_400_term_jaccard", "Term"),
                                         ("cfg2_300_tokenids_jaccard", "TokenIds"),
                                         ("variable_80_jaccard", "Variable")])
def test_device_packed_path_matches_reference_golden(engine, name, column):
    """pairing.upload_levels (host: tokens -> codes; GPU: everything else) + the Jaccard kernel
    against the reference-generated goldens."""
    from napkon_string_matching.gpu import pairing

    meta, inputs, arrays = load_golden(name)
    L = [gen_comp_value(v) if v is not None else None for v in inputs["left"][column]]
    R = [gen_comp_value(v) if v is not None else None for v in inputs["right"][column]]
    lkeep = np.array([i for i, v in enumerate(L) if v is not None])
    rkeep = np.array([i for i, v in enumerate(R) if v is not None])
    dl, dr, _ = pairing.upload_levels(engine, [L[i] for i in lkeep], [R[i] for i in rkeep],
                                      "intersection_vs_union")
    assert dl.sizes is not None  # packed on the device, not by the numpy packer
    out = engine.all_pairs(dl, dr, meta["kwargs"]["score_threshold"])
    assert_same_triples((lkeep[out["left"]], rkeep[out["right"]], out["score"]),
                        (arrays["left_pos"], arrays["right_pos"], arrays["score"]))


def test_full_size_term_pack(engine):
    """200k Term items: device pack equals the numpy packer (the 1M-item cohorts of cfg5 go
    through the same code; the host packer needs ~10 s per million items)."""
    pl, f = syn.term_level_sets(200_000, 9)
    rank = pack.frequency_rank([f], 20000)
    (cohort,) = engine.device_packer.pack([dp.raw_from_parts(pl, f)], 20000)
    assert_same_pack(cohort, pack.pack_part_id_sets(pl, f, 20000, rank))


def test_oversized_items_fall_back_to_the_numpy_packer(engine):
    """An item beyond the device packer's 1024 ids per item sends the comparison through the
    numpy packer (pairing.upload_levels); the result is the oracle's either way."""
    from napkon_string_matching.gpu import pairing

    rng = np.random.default_rng(5)
    big = [[f"w{int(x)}" for x in rng.integers(0, 5000, size=1500)]]          # one level, 1500 ids
    L = [big] + [[[f"w{int(x)}" for x in rng.integers(0, 5000, size=6)]] for _ in range(40)]
    R = [[[f"w{int(x)}" for x in rng.integers(0, 5000, size=700)]] for _ in range(30)]
    dl, dr, _ = pairing.upload_levels(engine, L, R, "intersection_vs_union")
    assert dl.sizes is None                       # numpy-packed
    out = engine.all_pairs(dl, dr, 0.0)
    pl, pr = pack.pack_sets(L, R)
    want, _ = c_oracle.all_pairs(pl, pr, 0.0)
    assert_same_triples((out["left"], out["right"], out["score"]),
                        (want["left"], want["right"], want["score"]))
