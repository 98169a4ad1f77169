"""
The two pair score functions, with the reference's names and signatures
(/root/reference/napkon_string_matching/compare/score_functions.py:6-27) so that
``getattr(score_functions, config["score_func"])`` keeps working.

Each call packs its two operands and runs them as a one-by-one batch through the same CUDA
library the all-pairs path uses (gpu/engine.py); there is no CPU implementation here.  The
``*_many`` variants score one operand against many in a single launch (what
``MeshProvider.get_matches`` does with ``np.vectorize(fuzzy_match)``, terminology/mesh.py:209).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from napkon_string_matching.text.process import join_sorted, prepare_fuzzy_operand  # noqa: F401

#: score function name -> packed operand kind understood by the engine
KINDS = {"intersection_vs_union": "sets", "fuzzy_match": "strings"}


def _token_list(value) -> list:
    return value if isinstance(value, list) else value.split()


def _one_by_many(kind: str, left, rights: Sequence) -> np.ndarray:
    from napkon_string_matching.gpu import pack
    from napkon_string_matching.gpu.engine import default_engine

    engine = default_engine()
    if kind == "sets":
        pl, pr = pack.pack_sets([[_token_list(left)]], [[_token_list(r)] for r in rights])
    else:
        pl, pr = pack.pack_strings([[prepare_fuzzy_operand(left)]],
                                   [[prepare_fuzzy_operand(r)] for r in rights])
    out = engine.all_pairs(engine.upload(pl), engine.upload(pr), -1.0, flat=True)
    if kind == "sets" and len(out) != len(rights):
        # an empty set met an empty set: len(set()) / len(set()) in the reference
        raise ZeroDivisionError("division by zero")
    scores = np.empty(len(rights), dtype=np.float64)
    scores[out["right"]] = out["score"]
    return scores


def intersection_vs_union(left: List[str] | str, right: List[str] | str) -> float:
    """Ratio between the intersection and union of the tokens in `left` and `right`."""
    return float(_one_by_many("sets", left, [right])[0])


def fuzzy_match(left: str | List[str], right: str | List[str]) -> float:
    """QRatio/100 (normalised Indel similarity after default processing) of the two terms."""
    return float(_one_by_many("strings", left, [right])[0])


def intersection_vs_union_many(left, rights: Sequence) -> np.ndarray:
    return _one_by_many("sets", left, list(rights))


def fuzzy_match_many(left, rights: Sequence) -> np.ndarray:
    return _one_by_many("strings", left, list(rights))
