"""
TEST INFRASTRUCTURE — CPU restatement of the reference's cross-cohort comparison path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  The product (``napkon-string-matching_b200/``) never does; it fails loudly when
its CUDA library is missing.

Pure Python on purpose: every function mirrors the reference statement by statement (file:line
cited on each), operating on plain lists instead of pandas rows, so that small cases can be
checked by eye.  ``oracle/nsm_oracle.c`` is the same algorithm over the packed arrays for sizes
where Python is too slow.

Pinning:
* ``intersection_vs_union`` / ``compare_terms`` / ``gen_comparable`` pair loop: PINNED — checked
  against the unmodified reference run in the build container (tests/golden/make_golden.py
  imports /root/reference with ``nltk``/``rapidfuzz`` shims and writes tests/golden/*.npz).
* ``fuzzy_match``: **parity unpinned** — its arithmetic lives in rapidfuzz (pinned ``~=2.1.4``,
  requirements.txt:8), which is not vendored and not installable offline.  ``qratio`` below
  restates rapidfuzz 2.1.x's published algorithm and is anchored on the known answers from its
  documentation (tests/test_oracle.py).
"""
from __future__ import annotations

import re
from typing import Callable, Iterable, List, Sequence, Tuple

PREPARE_REMOVE_SYMBOLS = "!?,.()[]:;*"  # comparable_data.py:24


# --------------------------------------------------------------------------------------------
# compare/score_functions.py
# --------------------------------------------------------------------------------------------
def intersection_vs_union(left, right) -> float:
    """score_functions.py:6-13 — len(A∩B)/len(A∪B); ZeroDivisionError if both are empty."""
    set_left = set(left if isinstance(left, list) else left.split())
    set_right = set(right if isinstance(right, list) else right.split())
    return len(set_left.intersection(set_right)) / len(set_left.union(set_right))


def join_sorted(value: List[str]) -> str:
    """score_functions.py:16-17."""
    return " ".join(sorted(value, key=str.lower))


_NON_ALNUM = re.compile(r"(?u)\W")


def default_process(sentence: str) -> str:
    """rapidfuzz 2.1.x ``utils.default_process`` (Q6): non-alphanumerics -> blank, strip,
    lower.  ``_`` is alphanumeric under ``\\W``."""
    return _NON_ALNUM.sub(" ", sentence).strip().lower()


def lcs_length(a: Sequence, b: Sequence) -> int:
    """Textbook O(len(a)*len(b)) dynamic programme — deliberately NOT the bit-parallel
    recurrence the CUDA kernel uses, so the two can disagree."""
    if len(a) < len(b):
        a, b = b, a
    prev = [0] * (len(b) + 1)
    for ca in a:
        cur = [0]
        for j, cb in enumerate(b, 1):
            cur.append(prev[j - 1] + 1 if ca == cb else max(prev[j], cur[j - 1]))
        prev = cur
    return prev[-1]


def indel_ratio_from_counts(dist: int, lensum: int) -> float:
    """The one place the int -> double map of ``QRatio(...)/100`` lives (Q6, SURVEY §7 'hard
    parts'): ``norm_dist = dist/lensum`` (0 if lensum == 0); ``norm_sim = 1.0 - norm_dist``;
    ``ratio = norm_sim*100``; the reference then divides by 100 (score_functions.py:27)."""
    norm_dist = dist / lensum if lensum else 0.0
    norm_sim = 1.0 - norm_dist
    return (norm_sim * 100) / 100


def qratio_processed(a: str, b: str) -> float:
    """``QRatio/100`` of two already processed strings."""
    if not a or not b:
        return 0.0  # QRatio returns 0 when either processed string is empty
    lcs = lcs_length(a, b)
    return indel_ratio_from_counts(len(a) + len(b) - 2 * lcs, len(a) + len(b))


def fuzzy_match(left, right) -> float:
    """score_functions.py:20-27 with rapidfuzz's ``fuzz.QRatio`` restated (Q5, Q6)."""
    left_term = join_sorted(left) if isinstance(left, list) else left
    right_term = join_sorted(right) if isinstance(right, list) else right
    return qratio_processed(default_process(left_term), default_process(right_term))


SCORE_FUNCS = {"intersection_vs_union": intersection_vs_union, "fuzzy_match": fuzzy_match}


# --------------------------------------------------------------------------------------------
# types/comparable_data.py
# --------------------------------------------------------------------------------------------
def flatten_list(list_) -> List[str]:
    """comparable_data.py:567-574."""
    result = []
    for part in list_:
        if isinstance(part, list):
            result += part
        else:
            result.append(part)
    return result


def compare_terms(left, right, score_func: Callable) -> float:
    """comparable_data.py:248-265 (Q1).  Index starts at 1; both sides clamp at their deepest
    level; weight halves each step.  Returns int 0 when both are empty; IndexError when one is."""
    score = 0
    len_left, len_right = len(left), len(right)
    left_max, right_max = len_left - 1, len_right - 1
    factor = 1
    for i in range(1, max(len_left, len_right) + 1):
        score_ = score_func(left[min(i, left_max)], right[min(i, right_max)])
        factor /= 2
        score += score_ * factor
    return score


def tokenize(parts, word_tokenize: Callable[[str], List[str]], stop_words: Iterable[str]):
    """comparable_data.py:287-299 with the two nltk calls passed in."""
    tokens = word_tokenize(" ".join(flatten_list(parts)))
    stop = set(stop_words)
    tokens = {w for w in tokens if w.casefold() not in stop and w not in PREPARE_REMOVE_SYMBOLS}
    return sorted(tokens, key=lambda w: (w.casefold(), w))


def gen_comp_value(items, word_tokenize, stop_words):
    """comparable_data.py:283-285."""
    return [tokenize(items[-i:], word_tokenize, stop_words) for i in range(1, len(items) + 1)]


def categories_keep(cat_left, cat_right, first_left, first_right) -> bool:
    """comparable_data.py:464-476 — predicate picked from the types in row 0 (Q9)."""
    if isinstance(first_left, list):
        if isinstance(first_right, list):
            return (not set(cat_left).isdisjoint(set(cat_right))) or (not cat_left and not cat_right)
        # literally ``x in set(y)`` with x = left list: TypeError (unhashable) in the reference
        return cat_left in set(cat_right)
    if isinstance(first_right, list):
        return cat_left in set(cat_right)
    return cat_left == cat_right


def all_pairs(
    left_levels: Sequence[List[List[str]]],
    right_levels: Sequence[List[List[str]]],
    score_func: str,
    score_threshold: float,
    skip: Callable[[int, int], bool] | None = None,
) -> List[Tuple[int, int, float]]:
    """The hot loop, comparable_data.py:191 (cross product, row-major), :223-232 (one
    ``compare_terms`` per pair), :243 (keep ``>= score_threshold``).  ``skip(l, r)`` stands for
    the black-list / category filters applied before scoring (:195-218)."""
    func = SCORE_FUNCS[score_func]
    out = []
    for li, lv in enumerate(left_levels):
        for ri, rv in enumerate(right_levels):
            if skip is not None and skip(li, ri):
                continue
            score = compare_terms(lv, rv, func)
            if score >= score_threshold:
                out.append((li, ri, float(score)))
    return out
