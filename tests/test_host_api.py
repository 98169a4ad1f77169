"""CPU suite: the host side of the drop-in API (types, filters, exceptions, frames, sharding),
with pair scoring delegated to the oracle through the engine test double."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pandas as pd
import pytest

import golden_cases
from conftest import PKG, ROOT
from napkon_string_matching.gpu import distributed, pairing
from napkon_string_matching.types.comparable import Comparable, ComparisonResults
from napkon_string_matching.types.comparable_data import (ComparableData, flatten_list,
                                                          flatten_mapping, remove_existing_mappings)
from napkon_string_matching.types.mapping import Mapping
from napkon_string_matching.types.questionnaire import Questionnaire


@pytest.mark.parametrize("name", golden_cases.FRAME_CASES)
def test_gen_comparable_host_logic_matches_reference(oracle_engine, name):
    golden_cases.check_case(name)


def _q(identifiers, terms, **cols):
    return Questionnaire(pd.DataFrame({"Identifier": identifiers, "Term": terms,
                                       "Variable": [f"v{i}" for i in range(len(terms))], **cols}))


def test_unknown_score_func_raises_attribute_error(oracle_engine):
    q = _q(["a"], [["x y", "z"]])
    with pytest.raises(AttributeError):
        q.gen_comparable(q, Mapping(), Mapping(), score_func="no_such_func", compare_column="Term",
                         left_name="hap", right_name="pop")


def test_reference_exceptions_are_reproduced(oracle_engine):
    # both deepest levels tokenise to nothing (stop words only): ZeroDivisionError
    l = _q(["a"], [["und", "oder aber"]])
    r = _q(["b"], [["der", "die das"]])
    with pytest.raises(ZeroDivisionError):
        l.gen_comparable(r, Mapping(), Mapping(), score_func="intersection_vs_union",
                         compare_column="Term", left_name="hap", right_name="pop")
    # fuzzy_match never divides by zero on those
    res = l.gen_comparable(r, Mapping(), Mapping(), score_func="fuzzy_match", compare_column="Term",
                           left_name="hap", right_name="pop", score_threshold=0.0)
    assert len(res) == 1
    # an item with an empty list value meets a normal one: IndexError
    l = _q(["a"], [[]])
    with pytest.raises(IndexError):
        l.gen_comparable(r, Mapping(), Mapping(), score_func="intersection_vs_union",
                         compare_column="Term", left_name="hap", right_name="pop")
    # ... unless the pair is black-listed, in which case the reference never scores it
    bl = Mapping({"m": {"hap": ["a"], "pop": ["b"]}})
    res = l.gen_comparable(r, Mapping(), bl, score_func="intersection_vs_union",
                           compare_column="Term", left_name="hap", right_name="pop")
    assert len(res) == 0


def test_first_raising_pair_is_row_major():
    L = [[["a"]], [], [["a"], []]]
    R = [[["a"]], [["b"], []], []]
    assert pairing.first_raising_pair(L, R, True) == (IndexError, 0, 2)
    assert pairing.first_raising_pair(L[1:], R[:2], True) == (IndexError, 0, 0)
    assert pairing.first_raising_pair(L[2:], R[:2], True) == (ZeroDivisionError, 0, 1)
    assert pairing.first_raising_pair(L[2:], R[:2], False) is None
    assert pairing.first_raising_pair(L[:1], R[:2], True) is None
    assert pairing.first_raising_pair(L, R, True, skip=lambda l, r: l == 0) == (IndexError, 1, 0)


def test_compare_terms_and_gen_comp_value_scalar_path():
    jac = lambda a, b: len(set(a) & set(b)) / len(set(a) | set(b))
    assert ComparableData.compare_terms([["x"], ["a"]], [["a"]], jac) == 0.75
    assert ComparableData.compare_terms([["x"], ["a"], ["a", "b"]], [["y"], ["a"]], jac) == 0.6875
    assert ComparableData.compare_terms([], [], jac) == 0
    with pytest.raises(IndexError):
        ComparableData.compare_terms([], [["a"]], jac)
    assert ComparableData.gen_comp_value("gec_abc")[:3] == [["c"], ["b", "c"], ["a", "b", "c"]]
    assert flatten_list([["a", "b"], "c"]) == ["a", "b", "c"]


def test_mapping_semantics():
    m = Mapping({"1": {"hap": ["h1", "h2"], "pop": ["p1"]}, "2": {"hap": ["h3"], "suep": ["s1"]}})
    assert flatten_mapping("hap", "pop", m) == [("h1", "p1"), ("h2", "p1")]
    assert m.get_all_mapping_for_groups("hap", "suep") == [(["h3"], ["s1"])]
    with pytest.raises(KeyError):
        m.filter_by_group("pop")
    assert m.get_filtered(["2"]).dict() == {"2": {"hap": ["h3"], "suep": ["s1"]}}
    other = Mapping({"1": {"pop": ["p9"]}, "3": {"hap": ["h4"]}})
    m.update(other)
    assert m.dict()["1"]["pop"] == ["p1", "p9"] and "3" in m.dict()
    assert Mapping(m.dict()).dict() == m.dict()


def test_whitelist_removal_needs_members_on_both_sides():
    l = _q(["h1", "h2", "h3"], [["a"], ["b"], ["c"]])
    r = _q(["p1", "p2"], [["a"], ["b"]])
    wl = Mapping({"1": {"hap": ["h1"], "pop": ["p1"]}, "2": {"hap": ["h2"], "pop": ["nowhere"]}})
    remove_existing_mappings(l, r, "hap", "pop", wl)
    assert list(l.identifier) == ["h2", "h3"] and list(r.identifier) == ["p2"]
    l = _q(["h1"], [["a"]])
    remove_existing_mappings(l, r, "hap", "pop", Mapping({"1": {"hap": ["h1"]}}))  # KeyError inside
    assert list(l.identifier) == ["h1"]


def test_comparable_attribute_routing_and_json_roundtrip(tmp_path):
    frame = pd.DataFrame({"HapIdentifier": ["a", "b"], "HapVariable": ["va", "vb"],
                          "PopIdentifier": ["c", "d"], "PopVariable": ["vc", "vd"],
                          "MatchScore": [0.2, 0.9]})
    comp = Comparable(frame, left_name="Hap", right_name="Pop")
    assert list(comp.match_identifier) == ["a", "b"] and list(comp.identifier) == ["c", "d"]
    assert list(comp.match_score) == [0.2, 0.9]
    kept = comp[comp.match_score >= 0.5]
    assert isinstance(kept, Comparable) and len(kept) == 1 and kept.left_name == "Hap"
    comp.sort_by_score()
    assert list(comp.match_score) == [0.9, 0.2]
    comp.write_json(tmp_path / "c.json")
    back = Comparable.read_json(tmp_path / "c.json")
    assert back.left_name == "Hap" and list(back.match_score) == [0.9, 0.2]
    with pytest.raises(AttributeError):
        Comparable({"data": []})
    res = ComparisonResults()
    res["hap vs pop"] = comp
    res.write_excel(str(tmp_path / "out" / "result.xlsx"))
    assert any((tmp_path / "out").iterdir())


def test_category_masks_follow_the_row0_type_rule():
    cm = pairing.category_masks([["a", "b"], [], ["c"]], [["b"], [], ["x"]])
    assert cm["mode"] == "list_list"
    keep = lambda l, r: (int(cm["left"][l]) & int(cm["right"][r])) != 0 or \
        (int(cm["left"][l]) == 0 and int(cm["right"][r]) == 0)
    assert keep(0, 0) and keep(1, 1) and not keep(0, 1) and not keep(2, 2) and not keep(1, 0)
    cm = pairing.category_masks(["a", None, "b"], [["a", "c"], ["b"], []])
    assert cm["mode"] == "member" and int(cm["left"][1]) == 0
    cm = pairing.category_masks(["a", None], ["a", None])
    assert cm["mode"] == "equal" and int(cm["left"][1]) & int(cm["right"][1])
    with pytest.raises(TypeError):
        pairing.category_masks([["a"]], ["a"])
    many = pairing.category_masks([[f"c{i}" for i in range(70)]], [["c1"]])
    assert "host" in many


def test_matcher_steps_with_injected_data(oracle_engine, tmp_path):
    from napkon_string_matching import matching
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.matcher import Matcher
    from napkon_string_matching.types.gecco_definition import GeccoDefinition

    vocab = syn.vocabulary(1500)
    qs = {n: Questionnaire(syn.questionnaire_frame(40, s, vocab, n))
          for n, s in (("pop", 2), ("hap", 1), ("suep", 3))}
    gecco = GeccoDefinition(syn.definitions_frame(15, 4, vocab))
    config = {"matching": {"score_threshold": 0.1, "cache_threshold": 0.05, "compare_column": "Term",
                           "score_func": "intersection_vs_union", "variable_score_threshold": 0.9,
                           "filter_categories": False, "calculate_tokens": False},
              "steps": ["variables", "gecco", "questionnaires"],
              "output_dir": str(tmp_path / "out"), "cache_dir": str(tmp_path / "cache")}
    m = Matcher(None, config, gecco=gecco, questionnaires=qs)
    matching.match(config, matcher=m)
    names = list(m.results.results)
    assert names == ["var_hap vs pop", "var_hap vs suep", "var_pop vs suep", "gecco vs pop",
                     "gecco vs hap", "gecco vs suep", "hap vs pop", "hap vs suep", "pop vs suep"]
    assert m.results["hap vs pop"].left_name == "Hap"
    assert (m.results["hap vs pop"].match_score >= 0.1).all()
    assert any((tmp_path / "out").iterdir()) and any((tmp_path / "cache").iterdir())


def test_partition_rows_balances_weight():
    w = np.array([5, 1, 1, 1, 1, 1, 5, 5], dtype=float)
    parts = distributed.partition_rows(w, 3)
    assert parts[0][0] == 0 and parts[-1][1] == len(w)
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    sums = [w[b:e].sum() for b, e in parts]
    assert max(sums) <= 10
    assert distributed.partition_rows(np.ones(3), 5)[-1][1] == 3


GLOO_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {pkg!r}); sys.path.insert(0, {root!r})
    import numpy as np
    import torch.distributed as dist
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.gpu import distributed, pack
    from napkon_string_matching.gpu.lib import PAIR_DTYPE
    from oracle import c_oracle

    dist.init_process_group("gloo")
    lens, flat = syn.token_id_level_sets(300, 5, n_ids=400)
    p = pack.pack_suffix_id_sets(lens, flat, 400)
    score = lambda b, e: c_oracle.all_pairs(p, p, 0.2, l_begin=b, l_end=e)[0].astype(PAIR_DTYPE)
    weights = p.level_sizes().sum() * np.ones(p.n_items)
    want = score(0, p.n_items)
    key = lambda a: np.lexsort((a["right"], a["left"]))
    got = distributed.sharded_all_pairs(score, weights, gather="all")   # identical on every rank
    assert sum(distributed.last_counts) == len(want), distributed.last_counts
    assert len(distributed.last_counts) == 2 and min(distributed.last_counts) > 0
    assert np.array_equal(got[key(got)], want[key(want)])
    mine = distributed.sharded_all_pairs(score, weights, gather=False)
    assert len(mine) == distributed.last_counts[dist.get_rank()]
    # the default: rank 0 holds the complete result, the others the pairs of their own rows
    assert distributed.default_gather() == "rank0"
    got = distributed.sharded_all_pairs(score, weights)
    if dist.get_rank() == 0:
        assert np.array_equal(got[key(got)], want[key(want)])
    else:
        assert np.array_equal(got, mine)
    # several comparisons through one engine call per rank + ONE count all-gather
    from dataclasses import dataclass
    @dataclass
    class FakeCohort:
        n_items: int
        weights: np.ndarray
    @dataclass
    class FakeJob:
        left: FakeCohort
        right: FakeCohort
        threshold: float
        rows: tuple = None
    class FakeEngine:
        def run_jobs(self, jobs, **kw):
            outs = [c_oracle.all_pairs(p, p, j.threshold, l_begin=j.rows[0], l_end=j.rows[1])[0] for j in jobs]
            self.last_infos = [dict(count=len(o)) for o in outs]
            return outs
    side = FakeCohort(p.n_items, weights)
    outs, counts = distributed.sharded_run_jobs(FakeEngine(), [FakeJob(side, side, 0.2), FakeJob(side, side, 0.5, (10, 250))])
    assert len(counts) == 2 and len(counts[0]) == 2
    assert counts[0][0] + counts[1][0] == len(want)
    assert counts[0][1] + counts[1][1] == len(c_oracle.all_pairs(p, p, 0.5, l_begin=10, l_end=250)[0])
    assert [len(o) for o in outs] == counts[dist.get_rank()]
    # a block that fails on one rank fails the call on every rank (nobody hangs in the gather)
    def broken(b, e):
        if dist.get_rank() == 1:
            raise ZeroDivisionError("division by zero")
        return score(b, e)
    try:
        distributed.sharded_all_pairs(broken, weights)
        raise SystemExit("no exception")
    except ZeroDivisionError:
        assert dist.get_rank() == 1
    except RuntimeError as exc:
        assert dist.get_rank() == 0 and "rank 1" in str(exc)
    dist.destroy_process_group()
    sys.stdout.write("rank%sok" % os.environ["RANK"] + chr(10))
""")


def test_sharded_all_pairs_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(pkg=str(PKG), root=str(ROOT)))
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
        capture_output=True, text=True, timeout=300, env={**os.environ, "OMP_NUM_THREADS": "2"})
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2, res.stdout


def test_terminology_get_matches_host_logic(oracle_engine):
    import terminology_cases

    terminology_cases.check_all()
    terminology_cases.check_add_tokens()


def test_cache_json_fast_writer_is_byte_identical_to_json_dumps():
    """Comparable.to_json(orient="records", indent=4) — the reference's cache file format — is
    assembled column-wise with the C encoder; the text must equal the generic indented dump."""
    import json

    import pandas as pd

    from napkon_string_matching.types.comparable import Comparable

    def generic(c, indent):
        payload = {"left_name": c.left_name, "right_name": c.right_name,
                   "data": c.data.to_dict(orient="records")}
        return json.dumps(payload, indent=indent)

    rng = np.random.default_rng(0)
    nasty = ['plain', 'quote " and \\\\ backslash', 'umlaut äöüß €', 'tab\\tnewline\\n', 'nul \\x00 ctl \\x1f',
             '', '", "', '[not a list]', '{"not": "a dict"}', 'emoji \\U0001F600']
    frame = pd.DataFrame({
        "HapIdentifier": [nasty[i % len(nasty)] for i in range(57)],
        "HapVariable": [None if i % 7 == 0 else f"v{i}" for i in range(57)],
        "PopSheet": [float("nan") if i % 5 == 0 else f"s{i}" for i in range(57)],
        "Count": rng.integers(-5, 5, size=57),
        "Flag": rng.random(57) < 0.5,
        "MatchScore": np.where(rng.random(57) < 0.1, np.nan, rng.random(57)),
    })
    for indent in (4, 1, 2):
        for fr in (frame, frame.iloc[:1], frame.iloc[:0]):
            c = Comparable(data=fr, left_name="Hap", right_name='P"op')
            assert c.to_json(orient="records", indent=indent) == generic(c, indent)
    # cells that are lists go through the generic path and still round-trip
    c = Comparable(data=pd.DataFrame({"A": [[1, 2], [3]], "MatchScore": [0.5, 0.25]}), left_name="L", right_name="R")
    assert c.to_json(orient="records", indent=4) == generic(c, 4)
    # read_json(write_json(x)) == x
    c = Comparable(data=frame[["HapIdentifier", "MatchScore"]].dropna(), left_name="Hap", right_name="Pop")
    import tempfile, pathlib
    with tempfile.TemporaryDirectory() as d:
        c.write_json(pathlib.Path(d) / "c.json")
        back = Comparable.read_json(pathlib.Path(d) / "c.json")
    assert back.dataframe().reset_index(drop=True).equals(c.dataframe().reset_index(drop=True))
