"""
TEST INFRASTRUCTURE — times the UNMODIFIED reference's own compare path on host CPU cores
(``bench.py --impl reference`` and the ``cpu_baseline`` leg).

What runs is ``oracle/_ref/napkon_string_matching`` (oracle/make_ref.py: byte-for-byte copies of
/root/reference, git-ignored, shipped with the snapshot), imported in CHILD processes because the
product package has the same import name.  ``oracle/shims`` stands in for ``nltk``,
``rapidfuzz`` and ``psycopg2``, which are absent from the image (the reference cannot be imported
without them).  Timed per block of left rows:

* ``gen_comparable``  the reference's whole path — ``ComparableData.gen_comparable``
  (types/comparable_data.py:133-246: gen_comp_value per item, cross merge, black-list loop,
  one ``compare_terms`` per pair, threshold) on ``Questionnaire`` frames;
* ``vectorize``       the reference's flat use of a score function —
  ``np.vectorize(fuzzy_match)(synonyms["Term"], term)`` exactly as
  ``MeshProvider.get_matches`` calls it (terminology/mesh.py:209) — for the flat string workloads.

The reference's loop is single-threaded; the all-cores figure shards blocks of left rows over
``procs`` child processes in this harness, reference functions untouched.

A job (pickled to the children):
    {"mode": "gen_comparable" | "vectorize", "left": frame (dict of lists), "right": frame,
     "kwargs": gen_comparable kwargs, "rows_per_block": int, "evals_per_row": [per left row]}
"""
from __future__ import annotations

import json
import os
import pathlib
import pickle
import subprocess
import sys
import tempfile
import time

HERE = pathlib.Path(__file__).resolve().parent
REF = HERE / "_ref"
SHIMS = HERE / "shims"


def available() -> bool:
    return (REF / "napkon_string_matching" / "types" / "comparable_data.py").exists()


def manifest_digest() -> str:
    try:
        import hashlib

        return hashlib.sha256((REF / "MANIFEST.json").read_bytes()).hexdigest()[:16]
    except OSError:
        return ""


# ------------------------------------------------------------------------------------ child
def _child(workdir: str, index: int, procs: int, seconds: float) -> None:
    import logging

    logging.disable(logging.CRITICAL)
    import tqdm as _tqdm_mod

    _orig = _tqdm_mod.tqdm
    _tqdm_mod.tqdm = lambda it=None, *a, **k: _orig(it, *a, **{**k, "disable": True})

    import numpy as np
    import pandas as pd

    from napkon_string_matching.compare import score_functions as sf
    from napkon_string_matching.types.mapping import Mapping
    from napkon_string_matching.types.questionnaire import Questionnaire

    assert str(REF) in sf.__file__, sf.__file__   # the copy of the reference, not the product
    job = pickle.load(open(os.path.join(workdir, "job.pkl"), "rb"))
    left_all = pd.DataFrame(job["left"])
    right_df = pd.DataFrame(job["right"])
    rows = job["rows_per_block"]
    n_blocks = max(1, len(left_all) // rows)
    evals_per_row = np.asarray(job["evals_per_row"], dtype=np.float64)
    empty = Mapping()
    done_pairs = done_evals = kept = blocks = 0
    in_call = 0.0
    t_start = time.perf_counter()
    nxt = index
    while time.perf_counter() - t_start < seconds:
        b = nxt % n_blocks
        nxt += procs
        block = left_all.iloc[b * rows:(b + 1) * rows].reset_index(drop=True)
        t0 = time.perf_counter()
        if job["mode"] == "gen_comparable":
            res = Questionnaire(block).gen_comparable(
                Questionnaire(right_df), existing_mappings_whitelist=empty,
                existing_mappings_blacklist=empty, **job["kwargs"])
            kept += len(res)
        else:   # terminology/mesh.py:209
            column, thr = job["kwargs"]["compare_column"], job["kwargs"]["score_threshold"]
            for term in block[column]:
                scores = np.vectorize(sf.fuzzy_match)(right_df[column], term)
                kept += int((scores >= thr).sum())
        in_call += time.perf_counter() - t0
        done_pairs += len(block) * len(right_df)
        done_evals += float(evals_per_row[b * rows:(b + 1) * rows].sum())
        blocks += 1
    print(json.dumps({"pairs": done_pairs, "evals": done_evals, "kept": kept, "blocks": blocks,
                      "in_call_s": in_call, "wall_s": time.perf_counter() - t_start}), flush=True)


# ------------------------------------------------------------------------------------ parent
def run(job: dict, procs: int, seconds: float) -> dict:
    """Runs ``procs`` children for about ``seconds`` each, concurrently.  Returns
    ``{"evals_per_s", "pairs_per_s", "evals", "pairs", "kept", "blocks", "wall_s", "procs"}``;
    throughput = work of all children / the longest child's time inside the reference's calls."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where "
                           "/root/reference exists (the build container)")
    with tempfile.TemporaryDirectory() as work:
        pickle.dump(job, open(os.path.join(work, "job.pkl"), "wb"))
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join([str(SHIMS), str(REF)])
        env["OMP_NUM_THREADS"] = "1"
        children = [subprocess.Popen(
            [sys.executable, str(pathlib.Path(__file__).resolve()), "--child", work, str(i),
             str(procs), str(seconds)], env=env, cwd=work, stdout=subprocess.PIPE, text=True)
            for i in range(procs)]
        outs = []
        for c in children:
            stdout, _ = c.communicate()
            if c.returncode:
                raise RuntimeError(f"reference child failed (rc {c.returncode})")
            outs.append(json.loads(stdout.strip().splitlines()[-1]))
    longest = max(o["in_call_s"] for o in outs)
    total = {k: sum(o[k] for o in outs) for k in ("pairs", "evals", "kept", "blocks")}
    return {**total, "wall_s": longest, "procs": procs,
            "evals_per_s": total["evals"] / longest, "pairs_per_s": total["pairs"] / longest}


if __name__ == "__main__":
    if len(sys.argv) >= 6 and sys.argv[1] == "--child":
        _child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
