#!/usr/bin/env python
"""
bench.py — the reference's headline metric on B200: item pair-scores/sec (scored + thresholded).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one pass of the comparison hot path over one batch of synthetic cohorts:
the workload `tokenids50k` (BASELINE.json configs[1]) scores HAP x POP, HAP x SUEP and POP x SUEP,
50 000 items per cohort, `intersection_vs_union` on the `TokenIds` column, score_threshold 0.1.
A pair-score is one `score_func` evaluation that compare_terms asks for (max(K_left, K_right)
per item pair), either computed exactly or proven to belong to a pair below the threshold.

N > 1 (launched by torchrun, one rank per GPU): every rank scores its own left row block of the
same size against the replicated right cohort (weak scaling, no data-path collective); the only
collective is the NCCL all-gather of the per-rank kept-pair counts.

`--impl reference` times the CPU restatement of the reference's own Python pair loop
(oracle/reference_port.py, all host cores through multiprocessing) on a bounded sample of the
same workload; the reference itself is Python and /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "napkon-string-matching_b200"))
sys.path.insert(0, str(ROOT))

# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
# committed `ncu --set full` captures (profiles/r1_v9_jaccard_tokenids50k.txt,
# profiles/r1_v9_jaccard_term200k.txt, profiles/r1_qratio_v2_fuzzy20k.txt); other workloads: null
# dram__bytes_read.sum + dram__bytes_write.sum per kernel launch, from the committed ncu --set full
# captures (profiles/r1_v13_*.txt)
NCU_DRAM_BYTES_PER_LAUNCH = {"tokenids50k": 2.1932e9, "term200k": 4.45e7, "fuzzy20k": 2.3e6}

METRIC = "item pair-scores/sec (scored+thresholded)"
UNIT = "pair-scores/s"

WORKLOADS = {
    # name: (kind, items per cohort, threshold)
    "tokenids50k": dict(kind="tokenids", n=50_000, thr=0.1,
                        desc="HAPxPOPxSUEP, 50k items/cohort, intersection_vs_union on TokenIds, thr 0.1"),
    "tokenids5k": dict(kind="tokenids", n=5_000, thr=0.1, desc="reduced tokenids (debug)"),
    "term200k": dict(kind="term", n=200_000, n_right=200_000, thr=0.5,
                     desc="cfg5 shape reduced: 200k x 200k Term items (K 2-4), intersection_vs_union, thr 0.5"),
    "term1m": dict(kind="term", n=1_000_000, n_right=1_000_000, thr=0.5,
                   desc="cfg5: 1M x 1M Term items (K 2-4), intersection_vs_union, thr 0.5"),
    "defs1m": dict(kind="term", n=1_000_000, n_right=20_000, thr=0.5, defs=True,
                   desc="cfg4: 1M cohort items x 20k GECCO/KDS-style definitions, Tokens Jaccard, thr 0.5"),
    "term20k": dict(kind="term", n=20_000, n_right=20_000, thr=0.5, desc="reduced term (debug)"),
    "variable20k": dict(kind="variable", n=20_000, thr=0.9,
                        desc="the `variables` step: 20k x 20k items compared on the Variable column "
                             "(a str: one level per character suffix, K = 12-16), thr 0.9"),
    "fuzzy20k": dict(kind="fuzzy", n=20_000, thr=0.7,
                     desc="fuzzy_match flat strings 20k x 20k, avg 60 chars, thr 0.7"),
    "fuzzy200k": dict(kind="fuzzy", n=200_000, thr=0.7,
                      desc="cfg3: fuzzy_match flat strings 200k x 200k, avg 60 chars, thr 0.7"),
    "fuzzyterm10k": dict(kind="fuzzyterm", n=10_000, thr=0.5,
                         desc="the reference's shipped config: fuzzy_match on Term (K 2-4 levels), "
                              "10k x 10k items, cache_threshold 0.5"),
}


# ------------------------------------------------------------------------------------------
# workload construction (host, outside every timed region)
# ------------------------------------------------------------------------------------------
BUILD_INFO: dict = {}   # host_pack_s: seconds the numpy packer took for the workload's cohorts


def build_tokenids(n: int, rank: int):
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import pack

    seeds = {"hap": syn.SEED_LEFT, "pop": syn.SEED_RIGHT, "suep": syn.SEED_THIRD}
    packs, raw = {}, {}
    for name, seed in seeds.items():
        lens, flat = syn.token_id_level_sets(n, seed + 1000 * rank)
        raw[name] = (lens, flat)
        t0 = time.perf_counter()
        packs[name] = pack.pack_suffix_id_sets(lens, flat, 30000)
        BUILD_INFO["host_pack_s"] = BUILD_INFO.get("host_pack_s", 0.0) + time.perf_counter() - t0
    pairs = [("hap", "pop"), ("hap", "suep"), ("pop", "suep")]
    return packs, raw, pairs


def build_term(wl: dict, rank: int):
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import pack

    raw = {"left": syn.term_level_sets(wl["n"], syn.SEED_LEFT + 1000 * rank)}
    if wl.get("defs"):
        raw["right"] = syn.definition_level_sets(wl["n_right"], syn.SEED_DEFS + 1000 * rank)
    else:
        raw["right"] = syn.term_level_sets(wl["n_right"], syn.SEED_RIGHT + 1000 * rank)
    t0 = time.perf_counter()
    rank_map = pack.frequency_rank([f for _, f in raw.values()], 20000)
    packs = {k: pack.pack_part_id_sets(pl, f, 20000, rank_map) for k, (pl, f) in raw.items()}
    BUILD_INFO["host_pack_s"] = time.perf_counter() - t0
    return packs, raw, [("left", "right")]


def build_fuzzy(n: int, rank: int):
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import pack
    from napkon_string_matching.text.process import default_process

    vocab = syn.vocabulary()
    sl = [[default_process(s)] for s in syn.question_strings(n, syn.SEED_LEFT + 1000 * rank, vocab)]
    sr = [[default_process(s)] for s in syn.question_strings(n, syn.SEED_RIGHT + 1000 * rank, vocab)]
    pl, pr = pack.pack_strings(sl, sr)
    return {"left": pl, "right": pr}, {"left": sl, "right": sr}, [("left", "right")]


def build_fuzzyterm(n: int, rank: int):
    """Term items as strings: per level the joined, processed token string QRatio sees."""
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import pack

    vocab = syn.vocabulary()
    raw, levels = {}, {}
    for name, seed in (("left", syn.SEED_LEFT), ("right", syn.SEED_RIGHT)):
        part_lens, flat = syn.term_level_sets(n, seed + 1000 * rank)
        raw[name] = (part_lens, flat)
        starts = np.concatenate([[0], np.cumsum(part_lens.sum(axis=1))])
        items = []
        for i in range(n):
            pos, parts = int(starts[i]), []
            for q in part_lens[i]:
                if q:
                    parts.append([vocab[int(v)] for v in flat[pos:pos + q]])
                    pos += int(q)
            items.append([sorted({w for part in parts[-j:] for w in part}, key=lambda w: (w.casefold(), w))
                          for j in range(1, len(parts) + 1)])
        levels[name] = items
    pl, pr = pack.pack_strings(pack.fuzzy_level_strings(levels["left"]),
                               pack.fuzzy_level_strings(levels["right"]))
    return {"left": pl, "right": pr}, levels, [("left", "right")]


def build_variable(n: int, rank: int):
    """Variable names as the reference compares them: gen_comp_value of a str (Q2)."""
    from napkon_string_matching.gpu import pack
    from napkon_string_matching.text.tokenize import gen_comp_value

    levels = {}
    for name, seed in (("left", 1), ("right", 2)):
        rng = np.random.default_rng(seed + 1000 * rank)
        names = [f"{'gec_' if rng.random() < 0.2 else ''}{name[0]}{int(rng.integers(0, 4))}_v{int(rng.integers(0, 3 * n)):07d}"
                 for _ in range(n)]
        levels[name] = [gen_comp_value(v) for v in names]
    pl, pr = pack.pack_sets(levels["left"], levels["right"])
    return {"left": pl, "right": pr}, levels, [("left", "right")]


def build_workload(wl: dict, rank: int):
    if wl["kind"] == "variable":
        return build_variable(wl["n"], rank)
    if wl["kind"] == "fuzzyterm":
        return build_fuzzyterm(wl["n"], rank)
    if wl["kind"] == "tokenids":
        return build_tokenids(wl["n"], rank)
    if wl["kind"] == "term":
        return build_term(wl, rank)
    return build_fuzzy(wl["n"], rank)


def schedule_counts(left, right):
    """(evals, algorithmic int32 ops) of one comparison, in closed form from the packs.

    evals = sum over item pairs of max(K_l, K_r) (compare_terms, comparable_data.py:261).
    ops   = SURVEY.md §8(d): a + b per Jaccard evaluation (|A| + |B| of the two level sets),
            8 * ceil(m/64) * n per fuzzy evaluation (m = shorter string)."""
    kl, kr = left.levels_per_item(), right.levels_per_item()
    kmax = int(max(kl.max(initial=0), kr.max(initial=0)))
    hist_r = np.bincount(kr, minlength=kmax + 1).astype(np.float64)
    hist_l = np.bincount(kl, minlength=kmax + 1).astype(np.float64)
    ks = np.arange(kmax + 1)
    evals = float((np.maximum.outer(ks, ks) * np.outer(hist_l, hist_r))[1:, 1:].sum())
    if not hasattr(left, "tok") and kmax > 1:
        # levelled strings: 8 * ceil(min/64) * max per evaluation, estimated from a 512 x 512 sample
        # of items along compare_terms' schedule
        rng = np.random.default_rng(0)
        li = rng.integers(0, left.n_items, size=min(512, left.n_items))
        ri = rng.integers(0, right.n_items, size=min(512, right.n_items))
        ll, lr = left.level_lengths(), right.level_lengths()
        lo, ro = left.item_level_off.astype(np.int64), right.item_level_off.astype(np.int64)
        ops = n_ev = 0.0
        for t in range(1, kmax + 1):
            a = ll[lo[li] + np.minimum(t, kl[li] - 1)].astype(np.float64)
            b = lr[ro[ri] + np.minimum(t, kr[ri] - 1)].astype(np.float64)
            active = t <= np.maximum.outer(kl[li], kr[ri])
            mn, mx = np.minimum.outer(a, b), np.maximum.outer(a, b)
            ops += float((8.0 * np.ceil(mn / 64.0) * mx * active).sum())
            n_ev += float(active.sum())
        return evals, ops / max(n_ev, 1.0) * evals
    if not hasattr(left, "tok"):
        # flat strings: one evaluation per pair; sum over pairs of 8 * ceil(min/64) * max through
        # the two length histograms (an outer product over the items themselves is N^2 work:
        # 12 minutes of host time for 200k x 200k)
        ll, lr = left.level_lengths(), right.level_lengths()
        top = int(max(ll.max(initial=0), lr.max(initial=0))) + 1
        hl = np.bincount(ll, minlength=top).astype(np.float64)
        hr = np.bincount(lr, minlength=top).astype(np.float64)
        lens = np.arange(top, dtype=np.float64)
        per_pair = 8.0 * np.ceil(np.minimum.outer(lens, lens) / 64.0) * np.maximum.outer(lens, lens)
        ops = float(hl @ per_pair @ hr)
        return float(len(ll)) * float(len(lr)), ops

    def cum_sizes(p, k_items):
        """S[i, T] = sum_{t=1..T} size(level min(t, K_i - 1)) for T = 0..kmax."""
        sizes = p.level_sizes()
        off = p.item_level_off.astype(np.int64)
        out = np.zeros((p.n_items, kmax + 1))
        for t in range(1, kmax + 1):
            idx = off[:-1] + np.minimum(t, np.maximum(k_items - 1, 0))
            out[:, t] = out[:, t - 1] + np.where(k_items > 0, sizes[np.minimum(idx, len(sizes) - 1)], 0)
        return out

    sl, sr = cum_sizes(left, kl), cum_sizes(right, kr)
    ops = 0.0
    for k in range(1, kmax + 1):  # right items with K_r == k
        if hist_r[k]:
            ops += hist_r[k] * float(sl[np.arange(left.n_items), np.maximum(kl, k)].sum())
        if hist_l[k]:
            ops += hist_l[k] * float(sr[np.arange(right.n_items), np.maximum(kr, k)].sum())
    return evals, ops


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or \
               [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        num = lambda s: float(s) if s.replace(".", "", 1).isdigit() else None
        sm = [num(r[1]) for r in rows if num(r[1]) is not None]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": num(rows[0][2]),
                "reasons": reasons, "samples": len(rows),
                "power_w_max": max((num(r[3]) or 0.0) for r in rows)}


# ------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's Python pair loop)
# ------------------------------------------------------------------------------------------
def _levels_from_ids(lens, flat, begin, end):
    """String level lists of items begin:end, exactly what gen_comp_value yields for TokenIds."""
    starts = np.concatenate([[0], np.cumsum(lens)])
    out = []
    for i in range(begin, end):
        ids = [f"D{int(v):06d}" for v in flat[starts[i]:starts[i + 1]]]
        out.append([sorted(set(ids[-j:])) for j in range(1, len(ids) + 1)])
    return out


def _levels_from_parts(part_lens, flat, begin, end):
    """String level lists of Term-shaped items begin:end (what gen_comp_value yields)."""
    starts = np.concatenate([[0], np.cumsum(part_lens.sum(axis=1))])
    out = []
    for i in range(begin, end):
        pos, parts = int(starts[i]), []
        for n in part_lens[i]:
            if n:
                parts.append([f"w{int(v)}" for v in flat[pos:pos + n]])
                pos += int(n)
        out.append([sorted({w for part in parts[-j:] for w in part}) for j in range(1, len(parts) + 1)])
    return out


def _cpu_block(args):
    from oracle import reference_port as port

    left, right, func, thr = args
    t0 = time.perf_counter()
    kept = port.all_pairs(left, right, func, thr)
    evals = sum(max(len(a), len(b)) for a in left for b in right)
    return len(kept), evals, time.perf_counter() - t0


def cpu_reference(workload: dict, raw, pairs, seconds: float, procs: int, packs=None):
    """Times the CPU restatement of the reference's pair loop on left row blocks of the first
    comparison until about `seconds` of wall time are used.  Returns (pair-scores/s, sample).

    intersection_vs_union: oracle/reference_port.all_pairs — the reference's own Python loop
    (compare_terms + set arithmetic per pair), `procs` processes.
    fuzzy_match: the reference calls rapidfuzz's C++ QRatio from that Python loop; rapidfuzz is not
    installable here, so the C restatement (oracle/nsm_oracle.c, textbook LCS) stands in for it,
    `procs` OpenMP threads — a pure-Python LCS would understate the reference by ~1000x."""
    import multiprocessing as mp

    a, b = pairs[0]
    if workload["kind"] in ("fuzzy", "fuzzyterm"):
        from oracle import c_oracle

        os.environ["OMP_NUM_THREADS"] = str(procs)
        left, right = packs[a], packs[b].rows(0, min(2000, packs[b].n_items))
        rows_per_block, done_pairs, n_blocks = 64 * procs, 0, 0
        avail = max(1, left.n_items // rows_per_block)
        t_start = time.perf_counter()
        while time.perf_counter() - t_start < seconds:
            i = n_blocks % avail
            blk = left.rows(i * rows_per_block, min(left.n_items, (i + 1) * rows_per_block))
            c_oracle.all_pairs(blk, right, workload["thr"], flat=workload["kind"] == "fuzzy")
            done_pairs += int(np.maximum.outer(blk.levels_per_item(), right.levels_per_item()).sum())
            n_blocks += 1
        wall = time.perf_counter() - t_start
        sample = (f"{n_blocks} blocks of {rows_per_block} x {right.n_items} strings of {a} x {b} "
                  f"({done_pairs} pair-scores) through oracle/nsm_oracle.c (C, LCS by dynamic "
                  f"programming), {procs} OpenMP threads, {wall:.1f} s wall")
        return done_pairs / wall, sample

    n_right, rows_per_block = 1000, 40
    func, thr = "intersection_vs_union", workload["thr"]
    if workload["kind"] == "variable":   # raw holds the level lists themselves
        n_right = min(n_right, len(raw[b]))
        right = raw[b][:n_right]
        avail = max(1, len(raw[a]) // rows_per_block)
        block = lambda i: raw[a][(i % avail) * rows_per_block:(i % avail + 1) * rows_per_block]
    else:
        levels = _levels_from_ids if workload["kind"] == "tokenids" else _levels_from_parts
        n_right = min(n_right, len(raw[b][0]))
        right = levels(*raw[b], 0, n_right)
        avail = max(1, len(raw[a][0]) // rows_per_block)
        block = lambda i: levels(*raw[a], (i % avail) * rows_per_block, (i % avail + 1) * rows_per_block)
    done_evals = done_pairs = n_blocks = 0
    t_start = time.perf_counter()
    with mp.get_context("fork").Pool(procs) as pool:
        nxt = 0
        while time.perf_counter() - t_start < seconds:
            jobs = [(block(nxt + j), right, func, thr) for j in range(procs)]
            nxt += procs
            for kept, evals, _ in pool.map(_cpu_block, jobs):
                done_evals += evals
                done_pairs += rows_per_block * n_right
            n_blocks += procs
    wall = time.perf_counter() - t_start
    sample = (f"{n_blocks * rows_per_block} x {n_right} items of {a} x {b} "
              f"({done_pairs} item pairs, {done_evals} pair-scores) through "
              f"oracle/reference_port.all_pairs (Python), {procs} processes, {wall:.1f} s wall")
    return done_evals / wall, sample


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    procs = os.cpu_count() or 1
    small = {**wl, "n": min(wl["n"], 20000), "n_right": min(wl.get("n_right", wl["n"]), 20000)}
    packs, raw, pairs = build_workload(small, 0)
    per_step = max(5.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample = cpu_reference(wl, raw, pairs, per_step, procs, packs)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32+f64" if not wl["kind"].startswith("fuzzy") else "u64+f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    from napkon_string_matching.gpu.engine import Engine

    torch.cuda.set_device(local_rank)
    bound_cpus = None
    if world > 1:
        # one rank per GPU: keep the rank (and so its pinned record arena) on the GPU's NUMA node
        from napkon_string_matching.gpu import affinity

        bound_cpus = affinity.bind_to_gpu(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = WORKLOADS[args.workload]
    packs, raw, pairs = build_workload(wl, rank)
    host_pack_s = BUILD_INFO.get("host_pack_s", 0.0)
    flat = wl["kind"] == "fuzzy"
    thr = wl["thr"]

    eng = Engine()
    pinned = {k: eng.pin(p) for k, p in packs.items()}
    counts = [schedule_counts(packs[a], packs[b]) for a, b in pairs]
    evals_step = sum(c[0] for c in counts)
    ops_step = sum(c[1] for c in counts)
    item_pairs_step = sum(packs[a].n_items * packs[b].n_items for a, b in pairs)
    in_bytes = sum(p.nbytes() for p in packs.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    dev = {k: eng.upload(p, pinned[k]) for k, p in packs.items()}
    torch.cuda.synchronize()

    from napkon_string_matching.gpu.engine import Job

    def step_resident():
        eng.run_jobs([Job(dev[a], dev[b], thr, flat=flat) for a, b in pairs], to_host=False)
        return sum(i["count"] for i in eng.last_infos)

    # Token-set workloads go end to end from the host's token codes: H2D of the raw code CSR,
    # device-side packing (csrc/pack.cu; the frequency ranking included where the workload ranks),
    # the comparison kernels, D2H of the kept records.  String workloads start from host packs.
    raw_sets, pack_info, e2e_from = None, None, "host packs (numpy)"
    if wl["kind"] in ("tokenids", "term"):
        from napkon_string_matching.gpu import device_pack as dp

        make = dp.raw_from_id_lists if wl["kind"] == "tokenids" else dp.raw_from_parts
        names = list(packs)
        raw_sets = [make(*raw[k]).pin() for k in names]
        n_vocab, rank_mode = (30000, None) if wl["kind"] == "tokenids" else (20000, "frequency")
        e2e_from = "token codes (packed on the device)"
        in_bytes_e2e = sum(r.nbytes() for r in raw_sets) + (4 * n_vocab if rank_mode else 0)
        best = float("inf")
        for _ in range(3):
            torch.cuda.synchronize()
            t_pack = time.perf_counter()
            eng.device_packer.pack(raw_sets, n_vocab, rank=rank_mode)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t_pack)
        pack_info = {"device_ms": best * 1e3, "host_numpy_ms": host_pack_s * 1e3,
                     "what": "all cohorts of the step, token codes -> packed arrays in HBM"}
    else:
        in_bytes_e2e = in_bytes

    def step_e2e():
        if raw_sets is not None:
            d = dict(zip(names, eng.device_packer.pack(raw_sets, n_vocab, rank=rank_mode)))
        else:
            d = {k: eng.upload(p, pinned[k]) for k, p in packs.items()}
        outs = eng.run_jobs([Job(d[a], d[b], thr, flat=flat) for a, b in pairs], to_host=True,
                            copy=False)
        return sum(len(o) for o in outs), sum(i["d2h_bytes"] for i in eng.last_infos)

    # ---- kernel-resident timing (value) -------------------------------------------------
    for _ in range(max(3, args.warmup)):
        kept = step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = eng.launches
    eng.time_kernels, eng.kernel_ms, eng.kernel_launches_timed = True, 0.0, 0
    # L2 (126 MB) between timed steps: a step that writes more than 2x L2 of records flushes it by
    # itself; otherwise a 256 MB buffer is overwritten before every step, outside the timed events
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    self_flushing = 16 * kept > 2 * (126 << 20)

    def timed(step_fn):
        """K steps, each bracketed by CUDA events on the launching stream; returns (ms, last)."""
        pairs_ev, last = [], None
        for _ in range(args.steps):
            if not self_flushing:
                flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            last = step_fn()
            b.record(stream)
            pairs_ev.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in pairs_ev), last

    t0 = time.time()
    ms, kept = timed(step_resident)
    barrier()
    t1 = time.time()
    eng.time_kernels = False
    launches = eng.launches - launches0
    kernel_ms, kernel_launches = eng.kernel_ms, eng.kernel_launches_timed
    stats = {k: sum(i["stats"][k] for i in eng.last_infos) for k in eng.last_infos[0]["stats"]}
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    # ---- end-to-end timing (host packs -> host records) ---------------------------------
    for _ in range(max(3, args.warmup)):
        step_e2e()
    barrier()
    ms_e2e, (kept_e2e, d2h) = timed(step_e2e)
    barrier()
    assert kept_e2e == kept, (kept_e2e, kept)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([kept], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(gathered, cnt)  # the path's one collective: kept-pair counts
        kept_all = [int(g.item()) for g in gathered]
    else:
        kept_all = [kept]
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except OSError:
            pass
        int_peaks = {name: eng.microbench(kind) for kind, name in
                     enumerate(("lop3", "iadd3", "popc", "lcs_step_u64"))}
        peak_ops = max(int_peaks["lop3"], int_peaks["iadd3"])
        sec_step = ms * 1e-3 / args.steps
        value = evals_step * world / sec_step
        # the roofline is per kernel launch: algorithmic ops of one step / kernel time of one step
        # (CUDA events around every launch on the launching stream)
        kernel_sec_step = kernel_ms * 1e-3 / args.steps
        achieved = ops_step / kernel_sec_step
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = in_bytes + 16 * kept
        sec_step_hbm = kernel_sec_step
        if args.quick:   # kernel experiments: no CPU baseline leg (not a bench line to report)
            cpu_val, cpu_sample = None, "skipped (--quick)"
        elif world > 1:  # the CPU baseline is timed at N = 1 only
            cpu_val, cpu_sample = None, "timed at N=1 only"
        else:
            cpu_val, cpu_sample = cpu_reference(wl, raw, pairs, 12.0, 1, packs)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+f64" if not wl["kind"].startswith("fuzzy") else "u64+f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"],
                       "item_pairs_per_step_per_gpu": item_pairs_step,
                       "pair_scores_per_step_per_gpu": evals_step,
                       "kept_pairs_per_step": kept_all,
                       "l2": ("flushed by the step itself: each step writes %.0f MB of records, %.0fx "
                              "the 126 MB L2" % (16e-6 * kept, 16 * kept / (126 << 20))) if self_flushing
                       else "a 256 MB buffer is overwritten before every timed step (outside the events)"},
            "item_pairs_per_s": item_pairs_step * world / sec_step,
            "roofline": {"bound": "int32_alu", "achieved": achieved / 1e12, "peak": peak_ops / 1e12,
                         "unit": "Tiop/s", "frac": achieved / peak_ops,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(args.workload),
                         "peak_source": "nsm_microbench measured in this run (LOP3/IADD3 stream)",
                         "alg_ops_per_step": ops_step, "launches_per_step": kernel_launches / args.steps,
                         "kernel_ms_per_launch": kernel_ms / max(1, kernel_launches),
                         "kernel_share_of_step": kernel_ms / ms},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / sec_step_hbm / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / sec_step_hbm / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "int_peaks_tiops": {k: v / 1e12 for k, v in int_peaks.items()},
            "kernel_stats_per_step": stats,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": cpu_sample},
            "e2e": {"value": evals_step * world / (ms_e2e * 1e-3 / args.steps), "unit": UNIT,
                    "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": in_bytes_e2e,
                    "d2h_bytes_per_step": d2h, "from": e2e_from},
            "pack": pack_info,
            "gpu_launches": launches,
            "numa": {"rank0_bound_to_cpus": len(bound_cpus) if bound_cpus else None},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tokenids50k", choices=sorted(WORKLOADS))
    ap.add_argument("--quick", action="store_true",
                    help="skip the CPU baseline leg (kernel experiments and ncu captures only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
