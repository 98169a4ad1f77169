#!/usr/bin/env python
"""Cold, one-shot timing through the drop-in API (VERDICT r1 "what's weak" 6): what a user of
`Questionnaire.compare()` waits for when the call runs ONCE — fresh engine, no arenas, no pinned
memory, no cache file — with the wall time of every stage.

    python tools/cold_compare.py [--n50k 50000] [--skip50k]

cfg1: 2 000 x 2 000 Term items, intersection_vs_union, thr 0.1, full compare() incl. the cache
JSON.  cfg2 shape: 50k x 50k TokenIds, thr 0.1 through gen_comparable (1.4e8 kept pairs; the
reference's cache JSON of that many records is not written) and once more at thr 0.5 (the shipped
cache_threshold).  Prints one JSON line per case."""
from __future__ import annotations

import argparse
import json
import pathlib
import sys
import tempfile
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "napkon-string-matching_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n50k", type=int, default=50000)
    ap.add_argument("--skip50k", action="store_true")
    args = ap.parse_args()

    import pandas as pd
    import torch

    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import engine as engine_mod
    from napkon_string_matching.gpu.stages import collect
    from napkon_string_matching.types.mapping import Mapping
    from napkon_string_matching.types.questionnaire import Questionnaire

    def run(name, left, right, **kw):
        engine_mod._default_engine = None          # a fresh engine: cold arenas, nothing pinned
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with collect() as stages:
            res = kw.pop("call")(left, right, **kw)
        wall = time.perf_counter() - t0
        eng = engine_mod.default_engine()
        info = eng.last_info
        print(json.dumps({"case": name, "wall_s": round(wall, 3), "kept": len(res),
                          "stages_s": {k: round(v, 3) for k, v in stages.items()},
                          "unaccounted_s": round(wall - sum(stages.values()), 3),
                          "engine": {k: info.get(k) for k in ("blocks", "reruns", "packets", "uncoded", "d2h_bytes")}}),
              flush=True)

    empty = Mapping()
    vocab = syn.vocabulary()
    torch.zeros(1, device="cuda")   # CUDA context creation is not part of any case
    hap = Questionnaire(syn.questionnaire_frame(2000, syn.SEED_LEFT, vocab, "hap"))
    pop = Questionnaire(syn.questionnaire_frame(2000, syn.SEED_RIGHT, vocab, "pop"))
    with tempfile.TemporaryDirectory() as cache:
        run("cfg1 2k x 2k Term, compare()", hap, pop,
            call=lambda l, r, **kw: l.compare(r, **kw), existing_mappings_whitelist=empty,
            existing_mappings_blacklist=empty, compare_column="Term", score_func="intersection_vs_union",
            score_threshold=0.1, left_name="hap", right_name="pop", cache_dir=cache)
    if args.skip50k:
        return
    n = args.n50k

    def cohort(name, seed):
        ids = syn.token_id_lists(n, seed)
        return Questionnaire(pd.DataFrame({
            "Identifier": [f"{name}#{i:07d}" for i in range(n)], "Sheet": "s",
            "Variable": [f"{name}_v{i:07d}" for i in range(n)], "Term": [["q"]] * n, "TokenIds": ids}))

    hap, pop = cohort("hap", syn.SEED_LEFT), cohort("pop", syn.SEED_RIGHT)
    for thr in (0.5, 0.1):
        run(f"{n} x {n} TokenIds, gen_comparable, thr {thr}", hap, pop,
            call=lambda l, r, **kw: l.gen_comparable(r, **kw), existing_mappings_whitelist=empty,
            existing_mappings_blacklist=empty, compare_column="TokenIds", score_func="intersection_vs_union",
            score_threshold=thr, left_name="hap", right_name="pop")


if __name__ == "__main__":
    main()
