"""CPU suite: the storage order of token-set cohorts (pack.chunked_level_order) and the CSR gather
that applies it to raw codes (device_pack.RawSets.reordered)."""
import numpy as np
import pytest

from napkon_string_matching.gpu import pack
from napkon_string_matching.gpu import device_pack as dp


@pytest.mark.parametrize("n,chunk", [(0, 512), (1, 512), (700, 512), (5000, 128), (4096, 128), (12345, 512), (100000, 128)])
def test_chunked_level_order_is_a_permutation_with_homogeneous_chunks(n, chunk):
    rng = np.random.default_rng(n + chunk)
    k = rng.integers(0, 11, size=n)
    perm = pack.chunked_level_order(k, chunk)
    assert sorted(perm.tolist()) == list(range(n))
    stored = k[perm]
    full = n // chunk
    # a chunk holds at most two adjacent level counts (it is a run of the sorted order) ...
    for c in range(full):
        vals = np.unique(stored[c * chunk:(c + 1) * chunk])
        assert vals.max() - vals.min() <= max(1, int(np.ceil(11 * chunk / max(n, 1))))
    # ... and any eighth of the rows sees nearly the cohort's mean level count
    if full >= 256:
        eighth = full // 8 * chunk
        means = [stored[i:i + eighth].mean() for i in range(0, full * chunk - eighth + 1, eighth)]
        assert max(means) - min(means) < 1.0


def test_raw_sets_reordered_gathers_items_groups_and_ids():
    rng = np.random.default_rng(3)
    items = [[[int(x) for x in rng.integers(0, 50, size=int(rng.integers(0, 5)))]
              for _ in range(int(rng.integers(0, 6)))] for _ in range(300)]
    raw = dp.raw_from_levels(items)
    got, perm = raw.ordered_by_levels(128)
    assert got.mode == raw.mode and got.n_items == raw.n_items and got.n_groups == raw.n_groups
    for pos, src in enumerate(perm):
        g0, g1 = int(got.item_grp_off[pos]), int(got.item_grp_off[pos + 1])
        levels = [got.ids[int(got.grp_id_off[g]):int(got.grp_id_off[g + 1])].tolist() for g in range(g0, g1)]
        assert levels == items[int(src)]
    k = np.diff(got.item_grp_off.astype(np.int64))
    assert np.all(np.diff(k[:256].reshape(2, 128), axis=1) >= 0)   # each chunk is a run of the sorted order
