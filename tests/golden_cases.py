"""Shared body of the gen_comparable / compare golden checks: the CPU suite runs it with the
oracle-backed engine double (host logic only), the GPU suite with the CUDA engine."""
import json
import tempfile

import numpy as np
import pandas as pd

from conftest import GOLDEN, load_golden
from napkon_string_matching.types.gecco_definition import GeccoDefinition
from napkon_string_matching.types.mapping import Mapping
from napkon_string_matching.types.questionnaire import Questionnaire

INDEX = json.loads((GOLDEN / "index.json").read_text())
FRAME_CASES = sorted(k for k, v in INDEX.items() if v["kind"] in ("gen_comparable", "compare"))
JACCARD_FRAME_CASES = [k for k in FRAME_CASES if "fuzzy" not in k]
FUZZY_FRAME_CASES = [k for k in FRAME_CASES if "fuzzy" in k]


def run_case(name):
    meta, inputs, arrays = load_golden(name)
    left_cls = GeccoDefinition if inputs.get("left_cls") == "gecco" else Questionnaire
    left = left_cls(pd.DataFrame(inputs["left"]))
    right = Questionnaire(pd.DataFrame(inputs["right"]))
    wl, bl = Mapping(inputs.get("whitelist")), Mapping(inputs.get("blacklist"))
    if meta["kind"] == "gen_comparable":
        res = left.gen_comparable(right, existing_mappings_whitelist=wl,
                                  existing_mappings_blacklist=bl, **meta["kwargs"])
    else:
        with tempfile.TemporaryDirectory() as cache:
            res = left.compare(right, existing_mappings_whitelist=wl,
                               existing_mappings_blacklist=bl, cache_dir=cache, **meta["kwargs"])
    return meta, inputs, arrays, res


def check_case(name):
    meta, inputs, arrays, res = run_case(name)
    df = res.dataframe()
    lp, rp = meta["left_prefix"], meta["right_prefix"]
    assert res.left_name == lp and res.right_name == rp
    assert list(df.columns) == meta["columns"]
    assert len(df) == meta["n_rows"]
    # same rows, identified by their position in the cross product (the reference's frame index)
    got = df.sort_index(kind="stable")
    order = np.argsort(arrays["frame_index"], kind="stable")
    assert np.array_equal(got.index.to_numpy(), arrays["frame_index"][order])
    assert np.array_equal(got["MatchScore"].to_numpy().view(np.uint64),
                          arrays["score"][order].view(np.uint64))
    lid = np.asarray(inputs["left"]["Identifier"], dtype=object)
    rid = np.asarray(inputs["right"]["Identifier"], dtype=object)
    assert list(got[lp + "Identifier"]) == list(lid[arrays["left_pos"][order]])
    assert list(got[rp + "Identifier"]) == list(rid[arrays["right_pos"][order]])
    assert list(got[lp + "Argument"].astype(str)) == list(arrays["left_argument"][order])
    assert list(got[lp + "Variable"].astype(str)) == list(arrays["left_variable"][order])
    if meta["kind"] == "compare":  # sorted by score, descending (ties in any order)
        scores = df["MatchScore"].to_numpy()
        assert np.all(scores[:-1] >= scores[1:])
        assert scores.min() >= meta["kwargs"]["score_threshold"]
    return res
