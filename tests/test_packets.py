"""The compact record format (include/nsm.h:nsm_packet_t).  CPU: the host decoder on hand-built
packets.  GPU: NSM_OUT_PACKETS against NSM_OUT_PAIRS and the oracle, the probe-sized pipeline of
the engine, and the product's multi-GPU path on real GPUs (skipped below two devices)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import PKG, ROOT, assert_same_triples
from napkon_string_matching import synthetic as syn
from napkon_string_matching.gpu import lib as nsmlib
from napkon_string_matching.gpu import pack


def test_decode_packets_on_hand_built_packets():
    rng = np.random.default_rng(5)
    n = 37
    pk = np.zeros(n, dtype=nsmlib.PACKET_DTYPE)
    want = []
    for i in range(n):
        cnt = int(rng.integers(0, nsmlib.PACKET_RECORDS + 1)) if i % 5 else nsmlib.PACKET_RECORDS
        l0, r0 = int(rng.integers(0, 1 << 20)) * 512, int(rng.integers(0, 1 << 18)) * 128
        pk[i]["left0"], pk[i]["right0"], pk[i]["count"] = l0, r0, cnt
        pk[i]["score"][:] = np.nan          # unused slots hold garbage
        pk[i]["local"][:] = 0xffff
        for s in range(cnt):
            li, rc, sc = int(rng.integers(0, 512)), int(rng.integers(0, 128)), float(rng.random())
            pk[i]["local"][s] = (li << 7) | rc
            pk[i]["score"][s] = sc
            want.append((l0 + li, r0 + rc, sc))
    out = nsmlib.decode_packets(pk)
    assert out.dtype == nsmlib.PAIR_DTYPE and len(out) == len(want)
    assert [(int(a), int(b), float(c)) for a, b, c in zip(out["left"], out["right"], out["score"])] == want
    assert len(nsmlib.decode_packets(pk[:0])) == 0


def test_decode_cpackets_on_hand_built_packets():
    rng = np.random.default_rng(6)
    table = np.full(nsmlib.DICT_SLOTS, 0xffffffffffffffff, dtype=np.uint64)
    slots = rng.choice(nsmlib.DICT_SLOTS, size=500, replace=False)
    table[slots] = rng.random(500).view(np.uint64)
    n = 29
    pk = np.zeros(n, dtype=nsmlib.CPACKET_DTYPE)
    want = []
    for i in range(n):
        cnt = int(rng.integers(0, nsmlib.CPACKET_RECORDS + 1)) if i % 4 else nsmlib.CPACKET_RECORDS
        l0, r0 = int(rng.integers(0, 1 << 20)) * 512, int(rng.integers(0, 1 << 18)) * 128
        pk[i]["left0"], pk[i]["right0"], pk[i]["count"] = l0, r0, cnt
        pk[i]["rec"][:] = 0xffffffff
        for s_ in range(cnt):
            li, rc, code = int(rng.integers(0, 512)), int(rng.integers(0, 128)), int(rng.choice(slots))
            pk[i]["rec"][s_] = (code << 16) | (li << 7) | rc
            want.append((l0 + li, r0 + rc, float(table[code:code + 1].view(np.float64)[0])))
    out = nsmlib.decode_cpackets(pk, table)
    assert [(int(a), int(b), float(c)) for a, b, c in zip(out["left"], out["right"], out["score"])] == want


@pytest.mark.parametrize("mode", [nsmlib.OUT_PAIRS, nsmlib.OUT_PACKETS, nsmlib.OUT_CODED])
def test_library_host_decoders_equal_the_numpy_decoders(mode):
    """nsm_decode_packets / nsm_decode_cpackets (plain CPU loops inside the library; no GPU needed)
    against the numpy decoders, with and without the stored-position -> item-index maps, on full,
    partial and empty packets; Records.decode goes through them."""
    from napkon_string_matching.gpu.engine import Records

    rng = np.random.default_rng(40 + mode)
    n = 513
    table = rng.random(nsmlib.DICT_SLOTS).view(np.uint64)
    if mode == nsmlib.OUT_PAIRS:
        part = np.zeros(n, dtype=nsmlib.PAIR_DTYPE)
        part["left"], part["right"], part["score"] = rng.integers(0, 5000, n), rng.integers(0, 4000, n), rng.random(n)
        want = part.copy()
    else:
        dt, cap = (nsmlib.PACKET_DTYPE, nsmlib.PACKET_RECORDS) if mode == nsmlib.OUT_PACKETS else \
            (nsmlib.CPACKET_DTYPE, nsmlib.CPACKET_RECORDS)
        part = np.zeros(n, dtype=dt)
        part["left0"], part["right0"] = rng.integers(0, 9, n) * 512, rng.integers(0, 31, n) * 128
        part["count"] = np.where(rng.random(n) < 0.7, cap, rng.integers(0, cap + 1, n))
        if mode == nsmlib.OUT_PACKETS:
            part["score"], part["local"] = rng.random((n, cap)), rng.integers(0, 1 << 16, (n, cap))
            want = nsmlib.decode_packets(part)
        else:
            part["rec"] = rng.integers(0, 1 << 32, (n, cap), dtype=np.uint64).astype(np.uint32)
            want = nsmlib.decode_cpackets(part, table)
    out = np.empty(len(part) * nsmlib.ENTRY_RECORDS[mode] + 7, dtype=nsmlib.PAIR_DTYPE)
    m = nsmlib.decode_into(out, mode, part, table)
    assert m == len(want) and np.array_equal(out[:m], want)
    lperm, rperm = rng.permutation(5200).astype(np.uint32), rng.permutation(4100).astype(np.uint32)
    mapped = want.copy()
    mapped["left"], mapped["right"] = lperm[want["left"]], rperm[want["right"]]
    m = nsmlib.decode_into(out, mode, part, table, lperm, rperm)
    assert m == len(want) and np.array_equal(out[:m], mapped)
    assert nsmlib.decode_into(out, mode, part[:0], table) == 0
    # Records: several parts of different formats in one result
    pairs = np.zeros(5, dtype=nsmlib.PAIR_DTYPE)
    pairs["left"], pairs["right"], pairs["score"] = np.arange(5), np.arange(5) + 10, np.arange(5) / 8
    rec = Records([(mode, part), (nsmlib.OUT_PAIRS, pairs)], len(want) + 5, lperm, rperm, table)
    got = rec.decode()
    tail = pairs.copy()
    tail["left"], tail["right"] = lperm[pairs["left"]], rperm[pairs["right"]]
    assert np.array_equal(got, np.concatenate([mapped, tail]))
    assert np.array_equal(Records([(nsmlib.OUT_PAIRS, pairs)], 5).decode(copy=False), pairs)


def test_library_sort_pairs_is_the_row_major_order():
    rng = np.random.default_rng(77)
    n, n_left, n_right = 20000, 300, 5000
    key = rng.choice(n_left * n_right, size=n, replace=False)
    rec = np.zeros(n, dtype=nsmlib.PAIR_DTYPE)
    rec["left"], rec["right"], rec["score"] = key // n_right, key % n_right, rng.random(n)
    got = nsmlib.sort_pairs(rec, n_left)
    assert np.array_equal(got, rec[np.lexsort((rec["right"], rec["left"]))])
    assert len(nsmlib.sort_pairs(rec[:0], n_left)) == 0
    assert np.array_equal(nsmlib.sort_pairs(rec[:1], n_left), rec[:1])
    with pytest.raises(nsmlib.NsmError):
        nsmlib.sort_pairs(rec, int(rec["left"].max()))      # a left index beyond n_left


def _tokenid_packs(nl, nr):
    lens, flat = syn.token_id_level_sets(nl, syn.SEED_LEFT)
    pl = pack.pack_suffix_id_sets(lens, flat, 30000)
    lens, flat = syn.token_id_level_sets(nr, syn.SEED_RIGHT)
    return pl, pack.pack_suffix_id_sets(lens, flat, 30000)


@pytest.mark.gpu
def test_packets_equal_pairs_and_oracle(engine):
    from oracle import c_oracle

    pl, pr = _tokenid_packs(1500, 1100)          # ragged: 3 left chunks (the last partial) x 9 right blocks
    dl, dr = engine.upload(pl), engine.upload(pr)
    want, _ = c_oracle.all_pairs(pl, pr, 0.1)
    old = engine.compact
    try:
        engine.compact = False
        plain = engine.all_pairs(dl, dr, 0.1)
        assert engine.last_info["packets"] == 0
        engine.compact = "packets"
        packed = engine.all_pairs(dl, dr, 0.1)
        info = engine.last_info
        assert info["packets"] > 0 and info["count"] == len(packed)
        # full packets except at most one per warp and unit
        assert info["packets"] <= len(packed) // nsmlib.PACKET_RECORDS + 4 * 3 * 9
        engine.compact = "coded"
        coded = engine.all_pairs(dl, dr, 0.1)
        info = engine.last_info
        assert info["packets"] > 0 and info["count"] == len(coded)
        assert info["packets"] <= (len(coded) - info["uncoded"]) // nsmlib.CPACKET_RECORDS + 4 * 3 * 9
        assert info["uncoded"] < 0.05 * len(coded)
        assert info["d2h_bytes"] < 5.5 * len(coded) + nsmlib.DICT_SLOTS * 8
        # a row block and an overflowing arena (exact re-run) in packet mode
        engine._buffers.pop("out0", None), engine._buffers.pop("out1", None)   # arenas only grow
        block = engine.all_pairs(dl, dr, 0.1, rows=(300, 1301), capacity=64)
        assert engine.last_info["reruns"] == 1
    finally:
        engine.compact = old
    for got in (plain, packed, coded):
        assert_same_triples((got["left"], got["right"], got["score"]),
                            (want["left"], want["right"], want["score"]))
    sel = (want["left"] >= 300) & (want["left"] < 1301)
    assert_same_triples((block["left"], block["right"], block["score"]),
                        (want["left"][sel], want["right"][sel], want["score"][sel]))


@pytest.mark.gpu
def test_probe_sized_pipeline_has_no_reruns(engine, monkeypatch):
    """A job large enough for the probe path: the probe block sizes the arenas of the rest (no
    overflow re-run), dense results switch to packets, and the union is the plain result."""
    pl, pr = _tokenid_packs(9000, 3000)
    dl, dr = engine.upload(pl), engine.upload(pr)
    monkeypatch.setattr(engine, "PIPELINE_MIN_PAIRS", 1 << 20)
    monkeypatch.setattr(engine, "PIPELINE_BLOCK_BYTES", 4 << 20)   # several blocks
    old = engine.compact
    monkeypatch.setattr(engine, "direct_host", False)   # the device-arena pipeline (Engine._run_jobs)
    try:
        engine.compact = "auto"
        piped = engine.all_pairs(dl, dr, 0.1)
        info = dict(engine.last_info)
        engine.compact = False
        monkeypatch.setattr(engine, "PIPELINE_MIN_PAIRS", 1 << 62)
        plain = engine.all_pairs(dl, dr, 0.1)
    finally:
        engine.compact = old
    assert info["reruns"] == 0 and info["blocks"] >= 3 and info["packets"] > 0
    assert info["d2h_bytes"] < 6 * len(plain) + nsmlib.DICT_SLOTS * 8   # ~4.5 bytes per kept pair on the wire
    assert_same_triples((piped["left"], piped["right"], piped["score"]),
                        (plain["left"], plain["right"], plain["score"]))
    # sparse results stay in the 16-byte format
    monkeypatch.setattr(engine, "PIPELINE_MIN_PAIRS", 1 << 20)
    engine.compact = "auto"
    sparse = engine.all_pairs(dl, dr, 0.6)
    assert engine.last_info["packets"] == 0 and engine.last_info["reruns"] == 0
    engine.compact = old
    want = plain[plain["score"] >= 0.6]
    assert_same_triples((sparse["left"], sparse["right"], sparse["score"]),
                        (want["left"], want["right"], want["score"]))


NCCL_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {pkg!r}); sys.path.insert(0, {root!r})
    import numpy as np
    import torch, torch.distributed as dist
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import distributed, pack
    from napkon_string_matching.gpu.engine import Engine, Job
    from oracle import c_oracle

    rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    eng = Engine()
    packs = []
    for n, seed in ((5000, 11), (3000, 12), (2500, 13)):
        lens, flat = syn.token_id_level_sets(n, seed)
        packs.append(pack.pack_suffix_id_sets(lens, flat, 30000))
    dev = [eng.upload(p) for p in packs]
    pairs = [(0, 1), (0, 2), (1, 2)]
    eng.compact = os.environ.get("NSM_TEST_COMPACT") or "auto"
    outs, counts = distributed.sharded_run_jobs(eng, [Job(dev[a], dev[b], 0.1) for a, b in pairs])
    assert len(counts) == dist.get_world_size()
    key = lambda a: np.lexsort((a["right"], a["left"]))
    for j, (a, b) in enumerate(pairs):
        want, _ = c_oracle.all_pairs(packs[a], packs[b], 0.1)
        assert sum(c[j] for c in counts) == len(want), (j, counts, len(want))
        lo, hi = distributed.partition_rows(dev[a].weights, dist.get_world_size())[dist.get_rank()]
        sel = (want["left"] >= lo) & (want["left"] < hi)
        mine, w = outs[j], want[sel]
        assert len(mine) == counts[dist.get_rank()][j] == len(w)
        mine, w = mine[key(mine)], w[key(w)]
        assert np.array_equal(mine["left"], w["left"]) and np.array_equal(mine["right"], w["right"])
        assert np.array_equal(mine["score"].view(np.uint64), w["score"].view(np.uint64))
    # the drop-in's single-comparison path: rank 0 ends up with the complete result
    got = distributed.sharded_all_pairs(lambda b, e: eng.all_pairs(dev[0], dev[1], 0.1, rows=(b, e)), dev[0].weights)
    if dist.get_rank() == 0:
        want, _ = c_oracle.all_pairs(packs[0], packs[1], 0.1)
        got, want = got[key(got)], want[key(want)]
        assert np.array_equal(got["left"], want["left"]) and np.array_equal(got["score"].view(np.uint64), want["score"].view(np.uint64))
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write("rank%sok" % os.environ["RANK"] + chr(10))
""")


@pytest.mark.gpu
@pytest.mark.parametrize("compact", ["auto", "coded", "packets"])
def test_sharded_run_jobs_on_real_gpus(tmp_path, compact):
    import torch

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n_dev < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(NCCL_WORKER.format(pkg=str(PKG), root=str(ROOT)))
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29581", str(script)],
        capture_output=True, text=True, timeout=600,
        env={**os.environ, "OMP_NUM_THREADS": "4", "NSM_TEST_COMPACT": compact})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == world, res.stdout


@pytest.mark.gpu
def test_records_stored_straight_into_the_host_arena(engine, monkeypatch):
    """Engine._run_jobs_direct: a probe of four chunks, then the rest of the rows with the pinned host
    arena as the kernel's output buffer.  Dense results arrive as coded packets, sparse ones as
    16-byte pairs; a rest that keeps far more than its probe announced is counted exactly and run
    again; several jobs share one call; a row range and a result beyond max_pairs_per_block (rest
    split into parts) give the same records."""
    from napkon_string_matching.gpu.engine import Job

    assert engine.direct_host
    pl, pr = _tokenid_packs(9000, 3000)
    dl, dr = engine.upload(pl), engine.upload(pr)
    monkeypatch.setattr(engine, "PIPELINE_MIN_PAIRS", 1 << 62)
    plain = engine.all_pairs(dl, dr, 0.1)
    assert engine.last_info["blocks"] == 1
    monkeypatch.setattr(engine, "PIPELINE_MIN_PAIRS", 1 << 20)
    old = engine.compact, engine.max_pairs_per_block
    try:
        engine.compact = "auto"
        direct = engine.all_pairs(dl, dr, 0.1)
        info = dict(engine.last_info)
        assert info["blocks"] == 2 and info["reruns"] == 0 and info["packets"] > 0
        assert info["count"] == len(plain) and info["d2h_bytes"] < 6 * len(plain) + nsmlib.DICT_SLOTS * 8
        assert_same_triples((direct["left"], direct["right"], direct["score"]),
                            (plain["left"], plain["right"], plain["score"]))
        # sparse: the rest goes out as 16-byte pairs (only the probe is in packets)
        sparse = engine.all_pairs(dl, dr, 0.6)
        want = plain[plain["score"] >= 0.6]
        assert engine.last_info["reruns"] == 0 and engine.last_info["d2h_bytes"] < 16 * len(want) + (4 << 20)
        assert_same_triples((sparse["left"], sparse["right"], sparse["score"]),
                            (want["left"], want["right"], want["score"]))
        # a row range, several jobs in one call, and a result split by max_pairs_per_block
        part = engine.all_pairs(dl, dr, 0.1, rows=(700, 8123))
        sel = (plain["left"] >= 700) & (plain["left"] < 8123)
        assert_same_triples((part["left"], part["right"], part["score"]),
                            (plain["left"][sel], plain["right"][sel], plain["score"][sel]))
        engine.max_pairs_per_block = len(plain) // 3
        outs = engine.run_jobs([Job(dl, dr, 0.1), Job(dl, dr, 0.6), Job(dl, dr, 0.1, rows=(0, 600))])
        assert engine.last_infos[0]["blocks"] >= 4
        for got, w in zip(outs, (plain, want, plain[plain["left"] < 600])):
            assert_same_triples((got["left"], got["right"], got["score"]), (w["left"], w["right"], w["score"]))
    finally:
        engine.compact, engine.max_pairs_per_block = old
    # probe rows that keep nothing: the rest overflows its (minimal) region, is counted and run again
    lens, flat = syn.token_id_level_sets(9000, syn.SEED_LEFT)
    lens = lens.copy(); flat = flat.copy()
    head = int(lens[:2048].sum())
    flat[:head] = 29999 - (flat[:head] % 64)        # ids the right side (almost) never holds
    odd = pack.pack_suffix_id_sets(lens, flat, 30000)
    do = engine.upload(odd)
    monkeypatch.setattr(engine, "PIPELINE_MIN_PAIRS", 1 << 62)
    want = engine.all_pairs(do, dr, 0.1)
    monkeypatch.setattr(engine, "PIPELINE_MIN_PAIRS", 1 << 20)
    got = engine.all_pairs(do, dr, 0.1)
    assert engine.last_info["reruns"] >= 1 and engine.last_info["count"] == len(want)
    assert_same_triples((got["left"], got["right"], got["score"]), (want["left"], want["right"], want["score"]))
