"""
Generates the committed golden vectors by running the UNMODIFIED reference
(/root/reference, read-only) in the build container.  Not run on the GPU box and not run by
pytest: the outputs (tests/golden/*.npz, *.json) are committed.

    python tests/golden/make_golden.py            # all fixtures
    python tests/golden/make_golden.py --full     # additionally the 2k x 2k config-1 run (slow,
                                                  # ~3 GB RAM); only its digest is stored

Stage A (this process, product package on sys.path) draws the seeded synthetic cohorts.
Stage B (child process, /root/reference + oracle/shims on sys.path) feeds them to the
reference's own ``gen_comparable`` / ``compare`` / ``compare_terms`` / score functions and dumps
what it returns.  ``nltk`` and ``rapidfuzz`` are absent from the image, so they are shimmed
(oracle/shims); Jaccard results are 100 % the reference's arithmetic, fuzzy results go through
the shim's restated QRatio and are marked ``"pinned": false``.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import pathlib
import pickle
import subprocess
import sys
import tempfile

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parents[1]
PKG = ROOT / "napkon-string-matching_b200"
REFERENCE = pathlib.Path("/root/reference")


# ---------------------------------------------------------------------------------- stage B
def stage_b(workdir: str) -> None:
    import logging

    import numpy as np
    import pandas as pd

    logging.disable(logging.CRITICAL)
    import tqdm as _tqdm_mod  # silence progress bars of the reference

    _orig = _tqdm_mod.tqdm
    _tqdm_mod.tqdm = lambda it=None, *a, **k: _orig(it, *a, **{**k, "disable": True})

    from napkon_string_matching.compare import score_functions as sf
    from napkon_string_matching.types.comparable_data import ComparableData
    from napkon_string_matching.types.gecco_definition import GeccoDefinition
    from napkon_string_matching.types.mapping import Mapping
    from napkon_string_matching.types.questionnaire import Questionnaire

    assert sf.__file__.startswith(str(REFERENCE)), sf.__file__
    jobs = pickle.load(open(os.path.join(workdir, "jobs.pkl"), "rb"))
    results = {}
    for name, job in jobs.items():
        kind = job["kind"]
        if kind == "score_func":
            fn = getattr(sf, job["func"])
            out = []
            for a, b in job["pairs"]:
                try:
                    out.append(float(fn(a, b)).hex())
                except Exception as e:  # noqa: BLE001
                    out.append("raises:" + type(e).__name__)
            results[name] = out
        elif kind == "compare_terms":
            fn = getattr(sf, job["func"])
            out = []
            for a, b in job["pairs"]:
                try:
                    v = ComparableData.compare_terms(a, b, fn)
                    out.append(float(v).hex())
                except Exception as e:  # noqa: BLE001
                    out.append("raises:" + type(e).__name__)
            results[name] = out
        elif kind == "get_matches":
            from napkon_string_matching.terminology.mesh import MeshProvider

            provider = MeshProvider(None)
            provider._synonyms = pd.DataFrame(job["synonyms"])
            provider._headings = pd.DataFrame(job["synonyms"]).drop_duplicates("Id")
            results[name] = [provider.get_matches(term, job["score_threshold"]) for term in job["terms"]]
        elif kind == "gen_comp_value":
            results[name] = [ComparableData.gen_comp_value(v) for v in job["values"]]
        elif kind in ("gen_comparable", "compare"):
            cls = {"questionnaire": Questionnaire, "gecco": GeccoDefinition}
            left = cls[job.get("left_cls", "questionnaire")](pd.DataFrame(job["left"]))
            right = cls[job.get("right_cls", "questionnaire")](pd.DataFrame(job["right"]))
            wl, bl = Mapping(job.get("whitelist")), Mapping(job.get("blacklist"))
            try:
                if kind == "gen_comparable":
                    res = left.gen_comparable(
                        right, existing_mappings_whitelist=wl, existing_mappings_blacklist=bl,
                        **job["kwargs"])
                else:
                    with tempfile.TemporaryDirectory() as cache:
                        res = left.compare(
                            right, existing_mappings_whitelist=wl,
                            existing_mappings_blacklist=bl, cache_dir=cache, **job["kwargs"])
                df = res.dataframe()
                results[name] = {
                    "columns": list(df.columns),
                    "index": np.asarray(df.index, dtype=np.int64),
                    "records": df.to_dict(orient="list"),
                    "left_name": res.left_name,
                    "right_name": res.right_name,
                }
            except Exception as e:  # noqa: BLE001
                results[name] = {"raises": type(e).__name__}
        else:
            raise ValueError(kind)
    pickle.dump(results, open(os.path.join(workdir, "results.pkl"), "wb"))


# ---------------------------------------------------------------------------------- stage A
def _digest(l_ids, r_ids, scores) -> str:
    rows = sorted(zip(l_ids, r_ids, [float(s).hex() for s in scores]))
    h = hashlib.sha256()
    for row in rows:
        h.update(("|".join(row) + "\n").encode())
    return h.hexdigest()


def stage_a(full: bool, only=None) -> None:
    import numpy as np

    sys.path.insert(0, str(PKG))
    from napkon_string_matching import synthetic as syn

    vocab = syn.vocabulary()
    jobs = {}

    # -- hand cases for the scalar functions (Q1, Q4, Q5/Q6) ------------------------------
    jac_pairs = [
        (["a", "b"], ["a", "b"]), (["a"], ["b"]), (["a", "b", "c"], ["b"]), (["Haus"], ["haus"]),
        (["a", "a", "b"], ["b", "b"]), ("x y z", "z y"), ("x y", ["x", "y", "w"]), ([], []),
        ([], ["a"]), (["a", "b", "c", "d", "e", "f", "g"], ["a", "c", "e", "x", "y"]),
    ]
    jobs["jaccard_hand"] = {"kind": "score_func", "func": "intersection_vs_union",
                            "pairs": jac_pairs}
    fuzzy_pairs = [
        ("this is a test", "this is a test!"), ("Dialyse", "Hatte Sie Dialyse oder sonstiges?"),
        ("Sonstiges", "Hatte Sie Dialyse oder sonstiges?"), ("", "abc"), ("!!!", "abc"), ("", ""),
        (["beta", "Alpha", "gamma"], ["alpha", "Beta"]), ("Größe in cm", "groesse (cm)"),
        ("a_b", "a b"), ("abc", "abc"), ("abc", "xyz"), ("kitten", "sitting"),
        ("x" * 70 + "abc", "abc" + "x" * 65), ("aaaa", "aa"),
    ]
    jobs["fuzzy_hand"] = {"kind": "score_func", "func": "fuzzy_match", "pairs": fuzzy_pairs}
    ct_pairs = [
        ([["x"], ["a"]], [["a"]]),
        ([["x"], ["a"], ["a", "b"]], [["y"], ["a"]]),
        ([["a"]], [["a"]]),
        ([["q"], ["a", "q"]], [["q"], ["a", "q"]]),
        ([["p"], ["p", "q"], ["p", "q", "r"], ["p", "q", "r", "s"]],
         [["p"], ["p", "q"], ["p", "q", "r"], ["p", "q", "r", "s"]]),
        ([["p"], ["p", "q"], ["p", "q", "r"], ["p", "q", "r", "s"]], [["s"], ["r", "s"]]),
        ([], []), ([], [["a"]]), ([["a"]], []),
        ([[], ["a"]], [[], ["b"]]), ([["a"], []], [["b"], []]),
        ([["a", "b", "c"]], [["z"], ["a"], ["a", "b"], ["a", "b", "c"], ["a", "b", "c", "d"]]),
    ]
    jobs["compare_terms_jaccard"] = {"kind": "compare_terms", "func": "intersection_vs_union",
                                     "pairs": ct_pairs}
    jobs["compare_terms_fuzzy"] = {"kind": "compare_terms", "func": "fuzzy_match",
                                   "pairs": [p for p in ct_pairs if p[0] and p[1]]}
    jobs["gen_comp_value"] = {"kind": "gen_comp_value", "values": [
        ["Kopf Teil", "Frage eins zwei", "Wert"], ["nur eine Frage"], "gec_abc", "hap_v0000012",
        ["D000001", "D000002", "D000001"], [["Alpha Beta", "Gamma"], "Delta"],
        ["Wie ist der Wert", "und die Summe"],
    ]}

    # -- terminology lookups: MeshProvider.get_matches on a synthetic synonym table
    rng = np.random.default_rng(77)
    syn_rows = [("D900001", "Dialyse"), ("D900001", "Renale Dialyse"), ("D900002", "Sonstiges"),
                ("D900003", "Entlassung"), ("D900003", "Entlassung aus dem Krankenhaus")]
    for i in range(395):
        words = [vocab[int(w)] for w in rng.integers(0, 3000, size=int(rng.integers(1, 4)))]
        syn_rows.append((f"D{int(rng.integers(0, 150)):06d}", " ".join(words)))
    synonyms = {"Id": [r[0] for r in syn_rows], "Term": [r[1] for r in syn_rows]}
    terms = [["Dialyse", "nach", "Entlassung"], "Dialyse nach Entlassung",
             "Hatte Sie Dialyse oder sonstiges?".split(), ["Dialyse"], ["zzz", "qqq"],
             [vocab[5], vocab[17]], [syn_rows[40][1]], []]
    jobs["get_matches"] = {"kind": "get_matches", "synonyms": synonyms, "terms": terms,
                           "score_threshold": 0.3}

    # -- gen_comparable runs -------------------------------------------------------------
    def frame(n, seed, name):
        return syn.questionnaire_frame(n, seed, vocab, name).to_dict(orient="list")

    hap, pop = frame(400, syn.SEED_LEFT, "hap"), frame(400, syn.SEED_RIGHT, "pop")
    jobs["cfg1_400_term_jaccard"] = {
        "kind": "gen_comparable", "left": hap, "right": pop,
        "kwargs": dict(score_func="intersection_vs_union", compare_column="Term",
                       score_threshold=0.1, left_name="hap", right_name="pop")}
    jobs["cfg1_400_term_jaccard_cat"] = {
        "kind": "gen_comparable", "left": hap, "right": pop,
        "kwargs": dict(score_func="intersection_vs_union", compare_column="Term",
                       score_threshold=0.1, left_name="hap", right_name="pop",
                       filter_categories=True)}
    # TokenIds (cfg2 shape) on 300 x 300 with a few missing values (dropna, Q7)
    l300, r300 = frame(300, syn.SEED_LEFT, "hap"), frame(300, syn.SEED_THIRD, "suep")
    tl, tr = syn.token_id_lists(300, syn.SEED_LEFT), syn.token_id_lists(300, syn.SEED_THIRD)
    for i in (3, 77, 150):
        tl[i] = None
    for i in (0, 299):
        tr[i] = None
    l300["TokenIds"], r300["TokenIds"] = tl, tr
    jobs["cfg2_300_tokenids_jaccard"] = {
        "kind": "gen_comparable", "left": l300, "right": r300,
        "kwargs": dict(score_func="intersection_vs_union", compare_column="TokenIds",
                       score_threshold=0.1, left_name="hap", right_name="suep")}
    # white-list + black-list (Q7, Q8) on 120 x 120
    l120, r120 = frame(120, syn.SEED_LEFT, "hap"), frame(120, syn.SEED_RIGHT, "pop")
    lid, rid = l120["Identifier"], r120["Identifier"]
    whitelist = {"w1": {"hap": [lid[5]], "pop": [rid[9], rid[10]]},
                 "w2": {"hap": [lid[17], lid[18]], "pop": [rid[40]]},
                 "w3": {"hap": ["hap#nowhere"], "pop": [rid[41]]}}
    blacklist = {"b1": {"hap": [lid[0], lid[1]], "pop": [rid[0], rid[2]]},
                 "b2": {"hap": [lid[30]], "suep": ["x"]},
                 "b3": {"hap": [lid[31]], "pop": list(rid[50:60])}}
    jobs["wl_bl_120_term_jaccard"] = {
        "kind": "gen_comparable", "left": l120, "right": r120, "whitelist": whitelist,
        "blacklist": blacklist,
        "kwargs": dict(score_func="intersection_vs_union", compare_column="Term",
                       score_threshold=0.05, left_name="hap", right_name="pop")}
    # a white-list entry lacking one group key: KeyError swallowed, removal skipped (Q7)
    jobs["wl_keyerror_120"] = {
        "kind": "gen_comparable", "left": l120, "right": r120,
        "whitelist": {**whitelist, "w4": {"hap": [lid[3]]}}, "blacklist": {},
        "kwargs": dict(score_func="intersection_vs_union", compare_column="Term",
                       score_threshold=0.3, left_name="hap", right_name="pop")}
    # Variable column (str -> per-character suffix levels, Q2), the `variables` step
    jobs["variable_80_jaccard"] = {
        "kind": "gen_comparable", "left": frame(80, syn.SEED_LEFT, "hap"),
        "right": frame(80, syn.SEED_RIGHT, "pop"),
        "kwargs": dict(score_func="intersection_vs_union", compare_column="Variable",
                       score_threshold=0.9, left_name="hap", right_name="pop")}
    # gecco (left) vs questionnaire: Variable := Identifier on the gecco side
    gec = syn.definitions_frame(60, syn.SEED_DEFS, vocab).to_dict(orient="list")
    jobs["gecco_60_vs_hap_150"] = {
        "kind": "gen_comparable", "left": gec, "left_cls": "gecco",
        "right": frame(150, syn.SEED_LEFT, "hap"),
        "kwargs": dict(score_func="intersection_vs_union", compare_column="Term",
                       score_threshold=0.1, left_name="gecco", right_name="hap")}
    # compare(): cache_threshold, re-filter at score_threshold, sort (Q10, Q11)
    jobs["compare_200_term_jaccard"] = {
        "kind": "compare", "left": frame(200, syn.SEED_LEFT, "hap"),
        "right": frame(200, syn.SEED_RIGHT, "pop"),
        "kwargs": dict(score_func="intersection_vs_union", compare_column="Term",
                       score_threshold=0.3, cache_threshold=0.2, left_name="hap",
                       right_name="pop", calculate_tokens=False, filter_column="Variable")}
    # fuzzy_match through the shimmed QRatio: NOT a pin of rapidfuzz, but pins compare_terms /
    # join_sorted / thresholding around it
    jobs["fuzzy_150_term"] = {
        "kind": "gen_comparable", "left": frame(150, syn.SEED_LEFT, "hap"),
        "right": frame(150, syn.SEED_RIGHT, "pop"),
        "kwargs": dict(score_func="fuzzy_match", compare_column="Term", score_threshold=0.5,
                       left_name="hap", right_name="pop")}
    jobs["fuzzy_60_question_str"] = {
        "kind": "gen_comparable", "left": frame(60, syn.SEED_LEFT, "hap"),
        "right": frame(60, syn.SEED_RIGHT, "pop"),
        "kwargs": dict(score_func="fuzzy_match", compare_column="Question", score_threshold=0.3,
                       left_name="hap", right_name="pop")}
    if full:
        jobs["cfg1_2000_term_jaccard"] = {
            "kind": "gen_comparable", "left": frame(2000, syn.SEED_LEFT, "hap"),
            "right": frame(2000, syn.SEED_RIGHT, "pop"),
            "kwargs": dict(score_func="intersection_vs_union", compare_column="Term",
                           score_threshold=0.1, left_name="hap", right_name="pop")}

    if only:
        jobs = {k: v for k, v in jobs.items() if k in only}
    with tempfile.TemporaryDirectory() as work:
        pickle.dump(jobs, open(os.path.join(work, "jobs.pkl"), "wb"))
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join([str(ROOT / "oracle" / "shims"), str(REFERENCE)])
        subprocess.run([sys.executable, __file__, "--stage-b", work], check=True, env=env,
                       cwd=work)
        results = pickle.load(open(os.path.join(work, "results.pkl"), "rb"))

    if only:
        if "get_matches" in jobs:
            gm = jobs["get_matches"]
            (HERE / "get_matches.json").write_text(json.dumps(
                {"pinned": False, "synonyms": gm["synonyms"], "score_threshold": gm["score_threshold"],
                 "cases": [{"term": t, "result": [[i, s, float(v).hex()] for i, s, v in r]}
                           for t, r in zip(gm["terms"], results["get_matches"])]},
                indent=1, ensure_ascii=False), encoding="utf-8")
        return
    scalars = {}
    for name in ("jaccard_hand", "fuzzy_hand", "compare_terms_jaccard", "compare_terms_fuzzy"):
        scalars[name] = {"pinned": "fuzzy" not in name,
                         "cases": [{"left": a, "right": b, "result": r}
                                   for (a, b), r in zip(jobs[name]["pairs"], results[name])]}
    scalars["gen_comp_value"] = {"pinned": True, "cases": [
        {"value": v, "result": r}
        for v, r in zip(jobs["gen_comp_value"]["values"], results["gen_comp_value"])]}
    (HERE / "scalar_cases.json").write_text(json.dumps(scalars, indent=1, ensure_ascii=False),
                                            encoding="utf-8")
    gm = jobs["get_matches"]
    (HERE / "get_matches.json").write_text(json.dumps(
        {"pinned": False, "synonyms": gm["synonyms"], "score_threshold": gm["score_threshold"],
         "cases": [{"term": t, "result": [[i, s, float(v).hex()] for i, s, v in r]}
                   for t, r in zip(gm["terms"], results["get_matches"])]},
        indent=1, ensure_ascii=False), encoding="utf-8")

    index = json.loads((HERE / "index.json").read_text()) if (HERE / "index.json").exists() else {}
    for name, job in jobs.items():
        if job["kind"] not in ("gen_comparable", "compare"):
            continue
        res = results[name]
        meta = {"kind": job["kind"], "kwargs": job["kwargs"], "pinned": "fuzzy" not in name,
                "n_left": len(job["left"]["Identifier"]),
                "n_right": len(job["right"]["Identifier"])}
        if "raises" in res:
            meta["raises"] = res["raises"]
            index[name] = meta
            continue
        rec, ln, rn = res["records"], res["left_name"], res["right_name"]
        meta.update(columns=res["columns"], left_prefix=ln, right_prefix=rn,
                    n_rows=len(res["index"]),
                    digest=_digest(rec[ln + "Identifier"], rec[rn + "Identifier"],
                                   rec["MatchScore"]))
        index[name] = meta
        if name.startswith("cfg1_2000"):
            continue  # digest only
        lpos = {v: i for i, v in enumerate(job["left"]["Identifier"])}
        rpos = {v: i for i, v in enumerate(job["right"]["Identifier"])}
        np.savez_compressed(
            HERE / f"{name}.npz",
            left_pos=np.array([lpos[v] for v in rec[ln + "Identifier"]], dtype=np.uint32),
            right_pos=np.array([rpos[v] for v in rec[rn + "Identifier"]], dtype=np.uint32),
            score=np.array(rec["MatchScore"], dtype=np.float64),
            frame_index=res["index"],
            left_argument=np.array(rec[ln + "Argument"], dtype=object).astype(str),
            left_variable=np.array(rec[ln + "Variable"], dtype=object).astype(str),
        )
        # the inputs too, so the GPU box can rebuild them without the generator drifting
        with open(HERE / f"{name}.inputs.json", "w", encoding="utf-8") as f:
            json.dump({"left": job["left"], "right": job["right"],
                       "left_cls": job.get("left_cls", "questionnaire"),
                       "whitelist": job.get("whitelist"), "blacklist": job.get("blacklist")},
                      f, ensure_ascii=False)
    (HERE / "index.json").write_text(json.dumps(index, indent=1, sort_keys=True))
    for k, v in sorted(index.items()):
        print(k, v.get("n_rows"), v.get("raises"), v.get("digest", "")[:12])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage-b", default=None)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default=None, help="comma separated job names (others keep their files)")
    a = ap.parse_args()
    if a.stage_b:
        stage_b(a.stage_b)
    else:
        stage_a(a.full, a.only.split(",") if a.only else None)
