"""Shared body of the terminology look-up checks: CPU suite with the engine
double, GPU suite with the CUDA engine.  Golden: the reference's MeshProvider.get_matches run on
a synthetic synonym table (tests/golden/make_golden.py; fuzzy arithmetic through the shim)."""
import json

import pandas as pd

from conftest import GOLDEN
from napkon_string_matching.terminology.provider import MeshProvider, TerminologyProvider

GOLD = json.loads((GOLDEN / "get_matches.json").read_text(encoding="utf-8"))


def provider():
    frame = pd.DataFrame(GOLD["synonyms"])
    return MeshProvider(None, synonyms=frame, headings=frame.drop_duplicates("Id"))


def check_matches(got, want, query):
    """Same (Id, score) multiset, scores bit-equal, best first, one row per Id, and the term is
    a synonym of that Id (the reference's tie order comes from an unstable sort)."""
    assert sorted((i, float(s).hex()) for i, _, s in got) == sorted((i, h) for i, _, h in want)
    scores = [s for _, _, s in got]
    assert scores == sorted(scores, reverse=True)
    assert len({i for i, _, _ in got}) == len(got)
    synonyms = set(zip(GOLD["synonyms"]["Id"], GOLD["synonyms"]["Term"]))
    assert all((i, t) in synonyms for i, t, _ in got)
    # ... and it is a synonym that reaches that score (two synonyms of one Id may tie)
    from oracle import reference_port as port

    for _, t, s in got[:25]:
        assert port.fuzzy_match(t, query) == s


def check_all():
    prov = provider()
    thr = GOLD["score_threshold"]
    terms = [c["term"] for c in GOLD["cases"]]
    many = prov.get_matches_many(terms, thr)
    for case, got in zip(GOLD["cases"], many):
        check_matches(got, case["result"], " ".join(case["term"]))
    # the one-term entry point and the combining provider agree with the batch
    single = prov.get_matches(terms[0], thr)
    assert single == many[0]
    combined = TerminologyProvider(None, providers=[prov])
    assert combined.get_matches(terms[0], thr) == many[0]
    assert combined.get_matches(["zzzzzzzz"], 0.99) is None
    assert combined.initialized


def check_add_tokens():
    from napkon_string_matching.prepare.match_preparator import MatchPreparator
    from napkon_string_matching.types.questionnaire import Questionnaire

    prep = MatchPreparator({"terminology": {"mesh": None}})
    prep.terminology_provider.providers[0] = provider()
    data = Questionnaire([
        {"Sheet": "s", "Header": None, "Question": "Dialyse", "Parameter": "Hatte Sie Dialyse oder sonstiges?"},
        {"Sheet": "s", "Header": None, "Question": "9", "Parameter": "8"},
    ])
    data.add_terms()
    prep.add_tokens(data, 0.25, verbose=False, timeout=None)
    assert any("Dialyse" in t for t in data.tokens[0]) and "D900001" in data.token_ids[0]
    assert data.tokens[1] is None and data.token_ids[1] is None and data.token_match[1] is None
    ids, terms, scores = zip(*data.token_match[0])
    assert ids == data.token_ids[0] and terms == data.tokens[0] and min(scores) >= 0.25
