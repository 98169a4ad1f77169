"""
String preparation for ``fuzzy_match``.

``join_sorted`` follows /root/reference/napkon_string_matching/compare/score_functions.py:16-17.
``default_process`` restates the default processor that rapidfuzz 2.1.x applies inside
``fuzz.QRatio`` (third-party, pinned ``rapidfuzz~=2.1.4`` in requirements.txt:8, not vendored):
every non-alphanumeric code point (regex ``\\W`` with the UNICODE flag, so ``_`` survives)
becomes a blank, the ends are trimmed and the string is lower-cased.
"""
from __future__ import annotations

import re
from typing import List

_NON_ALNUM = re.compile(r"(?u)\W")


def join_sorted(value: List[str]) -> str:
    return " ".join(sorted(value, key=str.lower))


def default_process(sentence: str) -> str:
    return _NON_ALNUM.sub(" ", sentence).strip().lower()


def prepare_fuzzy_operand(value) -> str:
    """What ``fuzzy_match`` hands to the distance: join (lists only), then process."""
    joined = join_sorted(value) if isinstance(value, list) else value
    return default_process(joined)
