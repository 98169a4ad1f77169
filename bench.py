#!/usr/bin/env python
"""
bench.py — the reference's headline metric on B200: item pair-scores/sec (scored + thresholded).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                    [--secondary none|NAME,NAME] [--quick]

One "step" is one pass of the comparison hot path over one batch of synthetic cohorts:
the workload `tokenids50k` (BASELINE.json configs[1]) scores HAP x POP, HAP x SUEP and POP x SUEP,
50 000 items per cohort, `intersection_vs_union` on the `TokenIds` column, score_threshold 0.1.
A pair-score is one `score_func` evaluation that compare_terms asks for (max(K_left, K_right)
per item pair), either computed exactly or proven to belong to a pair below the threshold
(`kernel_stats_per_step` says how many were computed exactly).

N > 1 (launched by torchrun, one rank per GPU) is STRONG scaling of that one job through the
product's multi-GPU path (gpu/distributed.py:sharded_run_jobs): every rank holds the same cohorts,
scores its own left row block (balanced by level sizes) of every comparison against the replicated
right cohort, and the ranks all-gather their kept-pair counts (NCCL) inside every step.  Records
stay sharded: each rank's go to its own pinned host arena.

With the default workload the line also carries `secondary`: BASELINE configs[4] (`term1m`,
1M x 1M Term items) and configs[2] (`fuzzy200k`), sharded the same way, a few steps each.

`--impl reference` times the UNMODIFIED reference's own `gen_comparable` (oracle/_ref, a copy of
/root/reference made by oracle/make_ref.py; nltk / rapidfuzz shimmed) on all host cores, on a
bounded sample of the same workload; without oracle/_ref it falls back to the oracle port.

The cohorts are stored the way the product path stores them (pack.chunked_level_order: items of one
level count in chunks of the kernel's unit).  Experiment switches (not used by the driver):
NSM_BENCH_ITEM_ORDER=drawn|sorted (other storage orders), NSM_BENCH_TRACE_TIMED=1 (+
NSM_BENCH_TRACE_DIR) records the engine's host-side marks of the timed end-to-end steps per rank.
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
    "mesh50k": dict(kind="mesh", n=50_000, n_right=300_000, thr=0.85,
                    desc="token enrichment (SURVEY 8 f1, MeshProvider.get_matches for a whole cohort): 50k item "
                         "terms x 300k MeSH-style synonyms, flat fuzzy_match, tokens.score_threshold 0.85"),
    "mesh5k": dict(kind="mesh", n=5_000, n_right=30_000, thr=0.85, desc="reduced mesh (debug)"),
    "fuzzyterm10k": dict(kind="fuzzyterm", n=10_000, thr=0.5,
                         desc="the reference's shipped config: fuzzy_match on Term (K 2-4 levels), "
                              "10k x 10k items, cache_threshold 0.5"),
}


# ------------------------------------------------------------------------------------------
# workload construction (host, outside every timed region)
# ------------------------------------------------------------------------------------------
BUILD_INFO: dict = {}   # host_pack_s: seconds the numpy packer took for the workload's cohorts
REF_POOL = 4000         # items per side the CPU reference arm may draw its sample from


def level_ordered(lens: np.ndarray, flat: np.ndarray):
    """A token-set cohort in the storage order the product path gives it (gpu/pairing.py:upload_levels,
    pack.chunked_level_order): items of one level count next to each other in chunks of the kernel's
    unit, the chunks interleaved.  ``lens``: ids per item (n,) or per part (n, Q); an item's level
    count is its number of non-empty parts.  NSM_BENCH_ITEM_ORDER=drawn keeps the drawn order."""
    if os.environ.get("NSM_BENCH_ITEM_ORDER") == "drawn":
        return lens, flat
    from napkon_string_matching.gpu import pack

    per_item = lens if lens.ndim == 1 else lens.sum(axis=1)
    k = lens if lens.ndim == 1 else (lens > 0).sum(axis=1)
    perm = pack.chunked_level_order(k, 512)
    if os.environ.get("NSM_BENCH_ITEM_ORDER") == "sorted":   # experiment: no interleaving of the chunks
        perm = np.argsort(k, kind="stable")
    start = np.cumsum(per_item) - per_item
    size = per_item[perm]
    new_start = np.cumsum(size) - size
    idx = np.repeat(start[perm] - new_start, size) + np.arange(int(size.sum()), dtype=np.int64)
    return lens[perm], flat[idx]


def build_tokenids(n: int, rank: int):
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import pack

    seeds = {"hap": syn.SEED_LEFT, "pop": syn.SEED_RIGHT, "suep": syn.SEED_THIRD}
    packs, raw = {}, {}
    for name, seed in seeds.items():
        lens, flat = level_ordered(*syn.token_id_level_sets(n, seed + 1000 * rank))
        raw[name] = (lens, flat)
        t0 = time.perf_counter()
        packs[name] = pack.pack_suffix_id_sets(lens, flat, 30000)
        BUILD_INFO["host_pack_s"] = BUILD_INFO.get("host_pack_s", 0.0) + time.perf_counter() - t0
    pairs = [("hap", "pop"), ("hap", "suep"), ("pop", "suep")]
    return packs, raw, pairs


def build_term(wl: dict, rank: int):
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import pack

    raw = {"left": level_ordered(*syn.term_level_sets(wl["n"], syn.SEED_LEFT + 1000 * rank))}
    if wl.get("defs"):
        raw["right"] = level_ordered(*syn.definition_level_sets(wl["n_right"], syn.SEED_DEFS + 1000 * rank))
    else:
        raw["right"] = level_ordered(*syn.term_level_sets(wl["n_right"], syn.SEED_RIGHT + 1000 * rank))
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


def build_mesh(wl: dict, rank: int):
    """Left: what get_matches scores for an item, " ".join(Term) (header, question, parameter);
    right: synonym terms of 1-4 words."""
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import pack
    from napkon_string_matching.text.process import default_process

    vocab = syn.vocabulary()
    rng = np.random.default_rng(syn.SEED_LEFT + 1000 * rank)
    body = syn._Drawer(rng, vocab, syn.zipf_probs(len(vocab)))
    head = syn._Drawer(rng, vocab[:400], syn.zipf_probs(400))
    terms = []
    for _ in range(wl["n"]):
        u = rng.random()
        parts = [head.text(2) for _ in range(0 if u < 0.4 else (1 if u < 0.7 else 2))]
        parts += [body.text(int(rng.integers(3, 10))), body.text(int(rng.integers(1, 6)))]
        terms.append([default_process(" ".join(parts))])
    rng = np.random.default_rng(syn.SEED_DEFS + 1000 * rank)
    words = syn._Drawer(rng, vocab, syn.zipf_probs(len(vocab)))
    synonyms = [[default_process(words.text(int(rng.integers(1, 5))))] for _ in range(wl["n_right"])]
    pl, pr = pack.pack_strings(terms, synonyms)
    return {"left": pl, "right": pr}, {"left": terms, "right": synonyms}, [("left", "right")]


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
            if i < REF_POOL:   # the Term value itself, for the reference arm
                raw.setdefault(name + "__terms", []).append([" ".join(part) for part in parts])
        levels[name] = items
    pl, pr = pack.pack_strings(pack.fuzzy_level_strings(levels["left"]),
                               pack.fuzzy_level_strings(levels["right"]))
    return {"left": pl, "right": pr}, {**levels, **{k: v for k, v in raw.items() if k.endswith("__terms")}}, \
        [("left", "right")]


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
        levels[name + "__names"] = names[:REF_POOL]
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
    if wl["kind"] == "mesh":
        return build_mesh(wl, rank)
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
# CPU arms: the reference itself (oracle/_ref through oracle/ref_arm.py) and the oracle port
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


def cpu_port(workload: dict, raw, pairs, seconds: float, procs: int, packs=None):
    """The CPU restatement (oracle/) of the pair loop on left row blocks of the first comparison
    for about `seconds` of wall time.  Returns (pair-scores/s, sample).  intersection_vs_union:
    oracle/reference_port.all_pairs (Python, `procs` processes); fuzzy_match: oracle/nsm_oracle.c
    (C, LCS by dynamic programming, `procs` OpenMP threads)."""
    import multiprocessing as mp

    a, b = pairs[0]
    if workload["kind"] in ("fuzzy", "fuzzyterm", "mesh"):
        from oracle import c_oracle

        os.environ["OMP_NUM_THREADS"] = str(procs)
        left, right = packs[a], packs[b].rows(0, min(2000, packs[b].n_items))
        rows_per_block, done_pairs, n_blocks = 64 * procs, 0, 0
        avail = max(1, left.n_items // rows_per_block)
        t_start = time.perf_counter()
        while time.perf_counter() - t_start < seconds:
            i = n_blocks % avail
            blk = left.rows(i * rows_per_block, min(left.n_items, (i + 1) * rows_per_block))
            c_oracle.all_pairs(blk, right, workload["thr"], flat=workload["kind"] != "fuzzyterm")
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


def reference_job(workload: dict, raw, pairs) -> dict:
    """The first comparison of the workload as frames the reference's own code takes
    (oracle/ref_arm.py): a pool of left rows (cut into blocks there) and 1000 right rows, the
    first rows of the same seeded cohorts the GPU arm scores."""
    a, b = pairs[0]
    kind, thr = workload["kind"], workload["thr"]
    n_right = 1000

    def frame(name, n, column, values, terms=None):
        return {"Identifier": [f"{name}#{i:07d}" for i in range(n)], "Sheet": ["s"] * n,
                "Variable": [f"{name}_v{i:07d}" for i in range(n)],
                "Term": terms if terms is not None else [["q"]] * n, **({column: values} if column != "Term" else {})}

    if kind == "tokenids":
        def ids(name, n):
            lens, flat = raw[name]
            n = min(n, len(lens))
            starts = np.concatenate([[0], np.cumsum(lens[:n])])
            return [[f"D{int(v):06d}" for v in flat[starts[i]:starts[i + 1]]] for i in range(n)], lens[:n]
        (lv, kl), (rv, kr) = ids(a, REF_POOL), ids(b, n_right)
        left, right = frame(a, len(lv), "TokenIds", lv), frame(b, len(rv), "TokenIds", rv)
        column, mode, func = "TokenIds", "gen_comparable", "intersection_vs_union"
    elif kind == "term":
        def terms(name, n):
            part_lens, flat = raw[name]
            n = min(n, len(part_lens))
            starts = np.concatenate([[0], np.cumsum(part_lens.sum(axis=1))])
            out = []
            for i in range(n):
                pos, parts = int(starts[i]), []
                for q in part_lens[i]:
                    if q:
                        parts.append(" ".join(f"w{int(v)}" for v in flat[pos:pos + q]))
                        pos += int(q)
                out.append(parts)
            return out, (part_lens[:n] > 0).sum(axis=1)
        (lv, kl), (rv, kr) = terms(a, REF_POOL), terms(b, n_right)
        left, right = frame(a, len(lv), "Term", None, lv), frame(b, len(rv), "Term", None, rv)
        column, mode, func = "Term", "gen_comparable", "intersection_vs_union"
    elif kind == "variable":
        lv, rv = raw[a + "__names"][:REF_POOL], raw[b + "__names"][:n_right]
        kl, kr = np.array([len(v) for v in lv]), np.array([len(v) for v in rv])
        left, right = frame(a, len(lv), "Variable", lv), frame(b, len(rv), "Variable", rv)
        column, mode, func = "Variable", "gen_comparable", "intersection_vs_union"
    elif kind == "fuzzyterm":
        lv, rv = raw[a + "__terms"][:REF_POOL], raw[b + "__terms"][:n_right]
        kl, kr = np.array([len(v) for v in lv]), np.array([len(v) for v in rv])
        left, right = frame(a, len(lv), "Term", None, lv), frame(b, len(rv), "Term", None, rv)
        column, mode, func = "Term", "gen_comparable", "fuzzy_match"
    else:   # flat strings: the reference's flat use of fuzzy_match is np.vectorize (mesh.py:209)
        lv, rv = [v[0] for v in raw[a][:REF_POOL]], [v[0] for v in raw[b][:n_right]]
        kl, kr = np.ones(len(lv), dtype=np.int64), np.ones(len(rv), dtype=np.int64)
        left, right = frame(a, len(lv), "Question", lv), frame(b, len(rv), "Question", rv)
        column, mode, func = "Question", "vectorize", "fuzzy_match"
    evals_per_row = np.maximum.outer(np.asarray(kl), np.asarray(kr)).sum(axis=1)
    rows_per_block = 50 if mode == "gen_comparable" else 10
    return {"mode": mode, "left": left, "right": right, "rows_per_block": rows_per_block,
            "evals_per_row": [float(v) for v in evals_per_row],
            "kwargs": dict(score_func=func, compare_column=column, score_threshold=thr,
                           left_name=a, right_name=b),
            "what": ("ComparableData.gen_comparable (comparable_data.py:133-246)" if mode == "gen_comparable"
                     else "np.vectorize(fuzzy_match) (terminology/mesh.py:209)")}


def cpu_reference(workload: dict, raw, pairs, seconds: float, procs: int, packs=None) -> dict:
    """`cpu_baseline` object: the reference's own code when oracle/_ref travelled with the
    snapshot (kind "reference"), else the oracle port (kind "port")."""
    from oracle import ref_arm

    if ref_arm.available():
        job = reference_job(workload, raw, pairs)
        res = ref_arm.run(job, procs, seconds)
        fuzzy = workload["kind"] in ("fuzzy", "fuzzyterm", "mesh")
        sample = (f"{res['blocks']} blocks of {job['rows_per_block']} x {len(job['right']['Identifier'])} items of "
                  f"{pairs[0][0]} x {pairs[0][1]} ({res['pairs']} item pairs, {res['evals']:.0f} pair-scores) through the "
                  f"unmodified reference's {job['what']} from oracle/_ref (manifest {ref_arm.manifest_digest()}), "
                  f"{procs} process(es), {res['wall_s']:.1f} s inside the reference's calls; nltk"
                  + (", rapidfuzz (QRatio restated, LCS in C)" if fuzzy else "") + " shimmed (absent from the image)")
        return {"value": res["evals_per_s"], "unit": UNIT, "cores": procs, "kind": "reference",
                "sample": sample, "item_pairs_per_s": res["pairs_per_s"]}
    value, sample = cpu_port(workload, raw, pairs, seconds, procs, packs)
    return {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": sample + " (oracle/_ref missing: run oracle/make_ref.py in the build container)"}


def workload_config(name: str, wl: dict, packs, pairs, world: int) -> dict:
    """`config` of the JSON line; identical keys and values in both arms."""
    counts = [schedule_counts(packs[a], packs[b]) for a, b in pairs]
    return {"workload": name, "desc": wl["desc"],
            "comparisons_per_step": len(pairs),
            "item_pairs_per_step": int(sum(packs[a].n_items * packs[b].n_items for a, b in pairs)),
            "pair_scores_per_step": float(sum(c[0] for c in counts)),
            "threshold": wl["thr"],
            "partitioning": ("one GPU" if world == 1 else
                             f"left row blocks over {world} GPUs, right cohort replicated (strong scaling)")}


def dtype_of(wl: dict) -> str:
    return "u64+f64" if wl["kind"] in ("fuzzy", "fuzzyterm", "mesh") else "int32+f64"


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    procs = os.cpu_count() or 1
    packs, raw, pairs = build_workload(wl, 0)
    per_step = max(5.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    vals, base, walls = [], None, []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        base = cpu_reference(wl, raw, pairs, per_step, procs, packs)
        if i >= args.warmup:
            vals.append(base["value"])
            walls.append(time.perf_counter() - t0)
    value = float(np.mean(vals))
    base["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": dtype_of(wl), "data": "synthetic",
        "config": workload_config(args.workload, wl, packs, pairs, world),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "each step is a bounded sample of the workload (see cpu_baseline.sample); "
                "ms_per_step is the wall time of one sample, child start-up included",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class Prepared:
    """One workload on this rank: host packs, pinned inputs, device cohorts, the step functions."""

    def __init__(self, name: str, eng, world: int):
        import torch

        from napkon_string_matching.gpu.engine import Job

        self.name, self.wl, self.eng, self.world = name, WORKLOADS[name], eng, world
        wl = self.wl
        BUILD_INFO.clear()
        self.packs, self.raw, self.pairs = build_workload(wl, 0)   # the same problem on every rank
        self.host_pack_s = BUILD_INFO.get("host_pack_s", 0.0)
        self.flat, self.thr = wl["kind"] in ("fuzzy", "mesh"), wl["thr"]
        self.config = workload_config(name, wl, self.packs, self.pairs, world)
        counts = [schedule_counts(self.packs[a], self.packs[b]) for a, b in self.pairs]
        self.evals_step = sum(c[0] for c in counts)
        self.ops_step = sum(c[1] for c in counts)
        self.in_bytes = sum(p.nbytes() for p in self.packs.values())
        self.pinned = {k: eng.pin(p) for k, p in self.packs.items()}
        self.dev = {k: eng.upload(p, self.pinned[k]) for k, p in self.packs.items()}
        torch.cuda.synchronize()
        self.Job = Job
        # Token-set workloads go end to end from the host's token codes: H2D of the raw code CSR,
        # device-side packing (csrc/pack.cu; the frequency ranking included where the workload
        # ranks), the comparison kernels, D2H of the kept records.  Strings start from host packs.
        self.raw_sets, self.pack_info, self.e2e_from = None, None, "host packs (numpy)"
        self.in_bytes_e2e = self.in_bytes
        if wl["kind"] in ("tokenids", "term"):
            from napkon_string_matching.gpu import device_pack as dp

            make = dp.raw_from_id_lists if wl["kind"] == "tokenids" else dp.raw_from_parts
            self.names = list(self.packs)
            self.raw_sets = [make(*self.raw[k]).pin() for k in self.names]
            self.n_vocab, self.rank_mode = (30000, None) if wl["kind"] == "tokenids" else (20000, "frequency")
            self.e2e_from = "token codes (packed on the device)"
            self.in_bytes_e2e = sum(r.nbytes() for r in self.raw_sets) + (4 * self.n_vocab if self.rank_mode else 0)
            best = float("inf")
            for _ in range(3):
                torch.cuda.synchronize()
                t_pack = time.perf_counter()
                eng.device_packer.pack(self.raw_sets, self.n_vocab, rank=self.rank_mode)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t_pack)
            self.pack_info = {"device_ms": best * 1e3, "host_numpy_ms": self.host_pack_s * 1e3,
                              "what": "all cohorts of the step, token codes -> packed arrays in HBM"}

    def jobs(self, dev):
        return [self.Job(dev[a], dev[b], self.thr, flat=self.flat) for a, b in self.pairs]

    def step_resident(self):
        """Inputs resident in HBM, records left on the device; the count all-gather included."""
        from napkon_string_matching.gpu import distributed

        _, counts = distributed.sharded_run_jobs(self.eng, self.jobs(self.dev), to_host=False)
        return counts

    def step_e2e(self):
        """The call a user makes: pinned host inputs -> this rank's records in pinned host memory."""
        from napkon_string_matching.gpu import distributed

        if self.raw_sets is not None:
            d = dict(zip(self.names, self.eng.device_packer.pack(self.raw_sets, self.n_vocab, rank=self.rank_mode)))
        else:
            d = {k: self.eng.upload(p, self.pinned[k]) for k, p in self.packs.items()}
        if self.eng.trace is not None:
            self.eng._mark("inputs uploaded / packed (launched)")
        outs, counts = distributed.sharded_run_jobs(self.eng, self.jobs(d), to_host=True, decode=False)
        if self.eng.trace is not None:
            self.eng._mark("counts gathered")
        self.last_records = outs
        infos = self.eng.last_infos
        return counts, sum(i["d2h_bytes"] for i in infos), sum(i["packets"] for i in infos), \
            sum(i["reruns"] for i in infos), sum(i["uncoded"] for i in infos)


def measure(prep: Prepared, steps: int, warmup: int, rank: int, world: int, int_peaks: dict, hbm_peak: float,
            separate_resident: bool, sampler=None) -> dict:
    """Times `steps` steps of one workload.  separate_resident: the device-resident steps (`value`)
    are timed by themselves, then the end-to-end steps; otherwise `value` comes from the CUDA
    events around the kernel launches inside the end-to-end steps (secondary workloads)."""
    import torch
    import torch.distributed as dist

    eng = prep.eng
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(step_fn, flush: bool):
        """K steps, each bracketed by CUDA events on the launching stream; returns (ms, last)."""
        evs, last = [], None
        for _ in range(steps):
            if flush:
                flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            last = step_fn()
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        timed.steps_ms = [round(a.elapsed_time(b), 3) for a, b in evs]
        return sum(timed.steps_ms), last

    def max_over_ranks(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t]

    clocks = None
    ms = kernel_ms = kernel_launches = launches = None
    stats = None
    if separate_resident:
        for _ in range(warmup):
            counts = prep.step_resident()
        kept_mine = counts[rank]
        # L2 (126 MB) between timed steps: a step that writes more than 2x L2 of records flushes it
        # by itself; otherwise a 256 MB buffer is overwritten before every step, outside the events
        self_flushing = 16 * sum(kept_mine) > 2 * (126 << 20)
        if sampler is not None:
            sampler.start()
            time.sleep(0.3)
        barrier()
        launches0 = eng.launches
        eng.time_kernels, eng.kernel_ms, eng.kernel_launches_timed = True, 0.0, 0
        t0 = time.time()
        ms, counts = timed(prep.step_resident, not self_flushing)
        barrier()
        t1 = time.time()
        eng.time_kernels = False
        launches = eng.launches - launches0
        kernel_ms, kernel_launches = eng.kernel_ms, eng.kernel_launches_timed
        stats = {k: sum(i["stats"][k] for i in eng.last_infos) for k in eng.last_infos[0]["stats"]}
        if sampler is not None:
            clocks = sampler.stop(t0, t1)
    # ---- end to end ---------------------------------------------------------------------
    for _ in range(warmup):
        counts_e2e, d2h, packets, reruns, uncoded = prep.step_e2e()
    self_flushing = d2h > 2 * (126 << 20)
    if not separate_resident and sampler is not None:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = eng.launches
    eng.time_kernels, eng.kernel_ms, eng.kernel_launches_timed = True, 0.0, 0
    t0 = time.time()
    eng.trace = [(time.perf_counter(), "timed region begins")] if os.environ.get("NSM_BENCH_TRACE_TIMED") else None
    import gc

    gc.collect()
    gc.disable()   # no collector pause inside a timed step (seen as one 300 ms step in eight at N = 8)
    ms_e2e, (counts_e2e, d2h, packets, reruns, uncoded) = timed(prep.step_e2e, not self_flushing)
    barrier()
    gc.enable()
    timed_trace = None
    if eng.trace:
        timed_trace = [(round((t - eng.trace[0][0]) * 1e3, 2), label) for t, label in eng.trace]
        out_dir = os.environ.get("NSM_BENCH_TRACE_DIR")
        if out_dir:   # every rank's marks of the timed region, to see which rank a slow step waited for
            pathlib.Path(out_dir, f"trace_{prep.config['workload']}_rank{rank}.json").write_text(
                json.dumps({"steps_ms": list(timed.steps_ms), "marks": timed_trace}))
        timed_trace = timed_trace[:40]
    t1 = time.time()
    eng.time_kernels = False
    e2e_kernel_ms, e2e_kernel_launches, e2e_launches = eng.kernel_ms, eng.kernel_launches_timed, eng.launches - launches0
    # host-side timeline of one more (untimed) end-to-end step: where the step's wall time goes
    eng.trace = []
    t_step = time.perf_counter()
    prep.step_e2e()
    torch.cuda.synchronize()
    timeline = [(round((t - t_step) * 1e3, 2), label) for t, label in eng.trace]
    timeline.append((round((time.perf_counter() - t_step) * 1e3, 2), "step returned"))
    eng.trace = None
    if not separate_resident:
        ms, kernel_ms, kernel_launches, launches = e2e_kernel_ms, e2e_kernel_ms, e2e_kernel_launches, e2e_launches
        stats = {k: sum(i["stats"][k] for i in eng.last_infos) for k in eng.last_infos[0]["stats"]}
        counts = counts_e2e
        if sampler is not None:
            clocks = sampler.stop(t0, t1)
    else:
        assert counts_e2e == counts, (counts_e2e, counts)

    e2e_steps_mine = list(timed.steps_ms)
    per_rank = [0.0] * world
    per_rank[rank] = ms_e2e / steps
    e2e_ms_per_rank = [round(v, 3) for v in sum_over_ranks(*per_rank)]
    ms, ms_e2e = max_over_ranks(ms, ms_e2e)
    names = list(stats)
    sums = sum_over_ranks(kernel_ms, launches, kernel_launches, d2h, packets, reruns, uncoded,
                          *[stats[k] for k in names])
    kernel_ms_sum, launches_sum, kernel_launches_sum, d2h_sum, packets_sum, reruns_sum, uncoded_sum = sums[:7]
    stats = dict(zip(names, (int(v) for v in sums[7:])))
    kept_jobs = [sum(c[j] for c in counts) for j in range(len(prep.pairs))]
    kept = sum(kept_jobs)

    sec_step = ms * 1e-3 / steps
    value = prep.evals_step / sec_step
    peak_ops = max(int_peaks["lop3"], int_peaks["iadd3"])
    # the roofline is per kernel launch and per GPU: algorithmic ops of the step / the kernel time
    # all GPUs spent on it (CUDA events around every launch on the launching stream)
    kernel_sec_step = kernel_ms_sum * 1e-3 / steps
    achieved = prep.ops_step / kernel_sec_step
    alg_bytes = prep.in_bytes * world + 16 * kept
    out = {
        "value": value, "ms_per_step": ms / steps,
        "config": {**prep.config, "kept_pairs_per_step": kept_jobs,
                   "kept_pairs_per_rank": [sum(c) for c in counts],
                   "value_timing": ("device-resident steps timed by themselves (records stay in HBM)"
                                    if separate_resident else
                                    "sum of the CUDA-event times of the kernel launches inside the end-to-end steps"),
                   "l2": ("flushed by the step itself: each step writes %.0f MB of records, %.0fx the 126 MB L2"
                          % (16e-6 * kept / world, 16 * kept / world / (126 << 20))) if self_flushing
                   else "a 256 MB buffer is overwritten before every timed step (outside the events)"},
        "item_pairs_per_s": prep.config["item_pairs_per_step"] / sec_step,
        "roofline": {"bound": "int32_alu", "achieved": achieved / 1e12, "peak": peak_ops / 1e12,
                     "unit": "Tiop/s", "frac": achieved / peak_ops,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(prep.name) if world == 1 else None,
                     "peak_source": "nsm_microbench measured in this run (LOP3/IADD3 stream), per GPU",
                     "alg_ops_per_step": prep.ops_step,
                     "launches_per_step": kernel_launches_sum / steps,
                     "kernel_ms_per_launch": kernel_ms_sum / max(1, kernel_launches_sum),
                     "kernel_share_of_step": kernel_ms_sum / world / ms if separate_resident else None,
                     "exactly_scored_share": stats["level_evals"] / max(1.0, prep.evals_step)},
        "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / kernel_sec_step / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": alg_bytes / kernel_sec_step / 1e9 / hbm_peak},
        "kernel_stats_per_step": stats,
        "e2e": {"value": prep.evals_step / (ms_e2e * 1e-3 / steps), "unit": UNIT,
                "ms_per_step": ms_e2e / steps, "h2d_bytes_per_step": prep.in_bytes_e2e * world,
                "d2h_bytes_per_step": int(d2h_sum), "from": prep.e2e_from,
                "to": "records in each rank's pinned host arena, in the C ABI's wire format",
                "record_format": ("packets (include/nsm.h NSM_OUT_CODED / NSM_OUT_PACKETS): %.2f bytes per kept "
                                  "pair (%d packets, %d pairs uncoded)"
                                  % (d2h_sum / max(1, kept), packets_sum, uncoded_sum)) if packets_sum else
                                 "nsm_pair_t: 16 bytes per kept pair",
                "kernel_ms_per_step": e2e_kernel_ms / steps, "overflow_reruns_per_step": reruns_sum / steps,
                "collective": "NCCL all-gather of the kept-pair counts, every step" if world > 1 else None,
                "ms_per_step_by_rank": e2e_ms_per_rank, "ms_steps_rank0": e2e_steps_mine,
                **({"timed_trace_rank0": timed_trace} if timed_trace else {}),
                "timeline_ms_rank0": timeline},
        "pack": prep.pack_info,
        "gpu_launches": int(launches_sum),
        "clocks": clocks,
    }
    return out


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
    eng = Engine()
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    int_peaks = {name: eng.microbench(kind) for kind, name in
                 enumerate(("lop3", "iadd3", "popc", "lcs_step_u64"))}

    wl = WORKLOADS[args.workload]
    prep = Prepared(args.workload, eng, world)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    warmup = max(3, args.warmup)
    res = measure(prep, args.steps, warmup, rank, world, int_peaks, hbm_peak, True, sampler)

    decode = None
    if rank == 0 and not args.quick:
        # host side of the packet format: ms to expand one comparison's records of the last step
        rec = max(prep.last_records, key=len)
        t0 = time.perf_counter()
        arr = rec.decode(copy=False)
        decode = {"records": int(len(arr)), "ms": (time.perf_counter() - t0) * 1e3,
                  "what": "numpy expansion of one comparison's records to (left, right, score) arrays; not in e2e"}
        del arr
    if args.quick:   # kernel experiments: no CPU baseline leg (not a bench line to report)
        cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "skipped (--quick)"}
    elif world > 1 or rank != 0:  # the CPU baseline is timed at N = 1 only
        cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "timed at N=1 only"}
    else:
        cpu = cpu_reference(wl, prep.raw, prep.pairs, 12.0, 1, prep.packs)
        port_value, port_sample = cpu_port(wl, prep.raw, prep.pairs, 5.0, 1, prep.packs)
        cpu["port"] = {"value": port_value, "sample": port_sample}

    secondary = {}
    names = [] if args.secondary == "none" else \
        (["term1m", "fuzzy200k"] if args.secondary == "default" and args.workload == "tokenids50k"
         else [n for n in args.secondary.split(",") if n in WORKLOADS and args.secondary != "default"])
    del prep
    torch.cuda.empty_cache()
    for name in names:
        eng._buffers.clear()
        torch.cuda.empty_cache()
        p2 = Prepared(name, eng, world)
        r2 = measure(p2, max(1, min(3, args.steps)), 1, rank, world, int_peaks, hbm_peak, False, None)
        r2["dtype"] = dtype_of(WORKLOADS[name])
        secondary[name] = r2
        del p2

    if rank == 0:
        line = {
            "metric": METRIC, "value": res.pop("value"), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": res.pop("ms_per_step"),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype_of(wl), "data": "synthetic",
            **res,
            "roofline_hbm": {**res["roofline_hbm"], "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "int_peaks_tiops": {k: v / 1e12 for k, v in int_peaks.items()},
            "cpu_baseline": cpu,
            "decode": decode,
            "secondary": secondary,
            "numa": {"rank0_bound_to_cpus": len(bound_cpus) if bound_cpus else None},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tokenids50k", choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default="default",
                    help="'default' (term1m,fuzzy200k beside the default workload), 'none', or a comma list")
    ap.add_argument("--quick", action="store_true",
                    help="skip the CPU baseline leg (kernel experiments and ncu captures only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        if args.quick and args.secondary == "default":
            args.secondary = "none"
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
