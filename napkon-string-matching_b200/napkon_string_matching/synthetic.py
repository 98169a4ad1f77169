"""
Seeded synthetic cohorts of the shapes BASELINE.json names.

There is no reference code for this: the reference ships no data and no benchmark.  The
shapes follow what its loaders produce: questionnaire ``Term = [*header, question, parameter]``
(/root/reference/napkon_string_matching/types/questionnaire.py:59-68), GECCO
``Term = [category, parameter, choice]`` (types/gecco_definition.py:54-61), ``TokenIds`` /
``Tokens`` lists from MeSH enrichment (prepare/match_preparator.py:69-73).

Words contain only letters (a-z, ä ö ü ß), so every tokeniser splits them the same way, and
none is a German stop word.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import pandas as pd

from napkon_string_matching.text.tokenize import stop_words

SEED_LEFT, SEED_RIGHT, SEED_THIRD, SEED_DEFS = 1001, 1002, 1003, 1004
_LETTERS = np.array(list("abcdefghijklmnopqrstuvwxyzäöüß"))
# rough German letter frequencies so that words look (and collide) like words
_LETTER_P = np.array(
    [6.5, 1.9, 3.1, 5.1, 17.4, 1.7, 3.0, 4.8, 7.6, 0.3, 1.2, 3.4, 2.5, 9.8, 2.5, 0.8, 0.02,
     7.0, 7.3, 6.2, 4.4, 0.7, 1.9, 0.03, 0.04, 1.1, 0.5, 0.3, 0.6, 0.3]
)
_LETTER_P = _LETTER_P / _LETTER_P.sum()


def vocabulary(size: int = 20000, seed: int = 7) -> List[str]:
    """``size`` distinct pseudo-German words of 3-14 letters; about a third are Title-case."""
    rng = np.random.default_rng(seed)
    stops = stop_words("german")
    seen, words = set(), []
    while len(words) < size:
        n = int(rng.integers(3, 15))
        w = "".join(rng.choice(_LETTERS, size=n, p=_LETTER_P))
        if rng.random() < 0.35:
            w = w[0].upper() + w[1:]
        if w in seen or w.casefold() in stops:
            continue
        seen.add(w)
        words.append(w)
    return words


def zipf_probs(n: int, s: float = 1.0) -> np.ndarray:
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** s
    return p / p.sum()


class _Drawer:
    """Batched Zipf draws from a vocabulary."""

    def __init__(self, rng, words, probs):
        self.rng, self.words = rng, np.asarray(words, dtype=object)
        self.cdf = np.cumsum(probs)
        self.cdf[-1] = 1.0

    def idx(self, n: int) -> np.ndarray:
        return np.searchsorted(self.cdf, self.rng.random(n), side="right")

    def text(self, n_words: int) -> str:
        return " ".join(self.words[self.idx(n_words)])


def questionnaire_frame(n_items: int, seed: int, vocab: List[str] | None = None,
                        name: str = "hap") -> pd.DataFrame:
    """cfg1 item shape: Header absent 40 % / 1 part 30 % / 2 parts 30 % (two words each, from the
    400 most frequent words), Question 3-9 words, Parameter 1-5 words; Term K in {2,3,4}."""
    vocab = vocab if vocab is not None else vocabulary()
    rng = np.random.default_rng(seed)
    body = _Drawer(rng, vocab, zipf_probs(len(vocab)))
    head = _Drawer(rng, vocab[:400], zipf_probs(400))
    rows = []
    for i in range(n_items):
        u = rng.random()
        n_head = 0 if u < 0.4 else (1 if u < 0.7 else 2)
        header = [head.text(2) for _ in range(n_head)]
        question = body.text(int(rng.integers(3, 10)))
        parameter = body.text(int(rng.integers(1, 6)))
        sheet = f"sheet{int(rng.integers(0, 12)):02d}"
        rows.append(
            {
                "Identifier": f"{name}#{sheet}#{i:07d}",
                "Sheet": sheet,
                "Header": header if header else None,
                "Question": question,
                "Parameter": parameter,
                "Variable": f"{'gec_' if rng.random() < 0.2 else ''}{name}_v{i:07d}",
                "Category": [f"cat{int(c):02d}" for c in
                             np.unique(rng.integers(0, 24, size=int(rng.integers(0, 3))))],
                "Term": [*header, question, parameter],
            }
        )
    return pd.DataFrame(rows)


def definitions_frame(n_items: int, seed: int = SEED_DEFS, vocab: List[str] | None = None
                      ) -> pd.DataFrame:
    """cfg4 right-hand side: GECCO-style ``Term = [category, parameter, choice]`` (K <= 3)."""
    vocab = vocab if vocab is not None else vocabulary()
    rng = np.random.default_rng(seed)
    body = _Drawer(rng, vocab, zipf_probs(len(vocab)))
    cats = _Drawer(rng, vocab[:200], zipf_probs(200))
    rows = []
    for i in range(n_items):
        category = cats.text(int(rng.integers(1, 3)))
        parameter = body.text(int(rng.integers(1, 6)))
        choice = body.text(int(rng.integers(1, 4))) if rng.random() < 0.6 else None
        rows.append(
            {
                "Identifier": f"gecco#{i:06d}",
                "Category": [category],
                "Parameter": parameter,
                "Choices": choice,
                "Term": [t for t in (category, parameter, choice) if t],
            }
        )
    return pd.DataFrame(rows)


def token_id_lists(n_items: int, seed: int, n_ids: int = 30000, max_len: int = 10
                   ) -> List[List[str]]:
    """cfg2 ``TokenIds`` column: 1..max_len ids (uniform length), Zipf over ``D%06d`` ids."""
    rng = np.random.default_rng(seed)
    cdf = np.cumsum(zipf_probs(n_ids))
    cdf[-1] = 1.0
    lens = rng.integers(1, max_len + 1, size=n_items)
    flat = np.searchsorted(cdf, rng.random(int(lens.sum())), side="right")
    out, pos = [], 0
    for n in lens:
        out.append([f"D{int(v):06d}" for v in flat[pos : pos + n]])
        pos += n
    return out


def token_id_level_sets(n_items: int, seed: int, n_ids: int = 30000, max_len: int = 10):
    """Same draw as :func:`token_id_lists`, returned already as integer level sets in the
    packed CSR layout (see gpu/pack.py) without going through Python strings: for an id list
    ``v`` level j is ``set(v[-(j+1):])`` (Q2).  Used for the 50k..1M item benchmark shapes."""
    rng = np.random.default_rng(seed)
    cdf = np.cumsum(zipf_probs(n_ids))
    cdf[-1] = 1.0
    lens = rng.integers(1, max_len + 1, size=n_items)
    flat = np.searchsorted(cdf, rng.random(int(lens.sum())), side="right").astype(np.uint32)
    return lens.astype(np.int64), flat


def question_strings(n_items: int, seed: int, vocab: List[str] | None = None,
                     mean: float = 60.0, sd: float = 15.0, lo: int = 8, hi: int = 160
                     ) -> List[str]:
    """cfg3 strings: length ~ clipped normal(60, 15) in [8, 160], vocabulary words joined by
    blanks and cut to the drawn length (no trailing blank)."""
    vocab = vocab if vocab is not None else vocabulary()
    rng = np.random.default_rng(seed)
    body = _Drawer(rng, vocab, zipf_probs(len(vocab)))
    target = np.clip(np.rint(rng.normal(mean, sd, size=n_items)), lo, hi).astype(int)
    out = []
    for t in target:
        s = body.text(int(t) // 4 + 2)
        while len(s) < t:
            s = s + " " + body.text(4)
        s = s[:t].rstrip()
        out.append(s if s else "a")
    return out


def cohort_frames(n_items: int, names=("hap", "pop", "suep")) -> Dict[str, pd.DataFrame]:
    seeds = {"hap": SEED_LEFT, "pop": SEED_RIGHT, "suep": SEED_THIRD}
    vocab = vocabulary()
    return {n: questionnaire_frame(n_items, seeds[n], vocab, n) for n in names}


def term_level_sets(n_items: int, seed: int, n_vocab: int = 20000):
    """cfg1 / cfg5 item shape (``Term = [*header, question, parameter]``) drawn directly as
    integer word ids, without Python strings, for the 200k..1M item benchmark shapes: per item up
    to four parts (header absent 40 % / one 30 % / two 30 %, two words each from the 400 most
    frequent words; question 3-9 words; parameter 1-5 words), Zipf(1.0) over ``n_vocab`` words.
    Returns ``(part_lens[n_items, 4], flat_ids)``: the word counts of (header1, header2, question,
    parameter) — 0 for an absent header — and all word ids, item by item, part by part.
    Level j of an item is the id set of its last j+1 present parts (Q2)."""
    rng = np.random.default_rng(seed)
    u = rng.random(n_items)
    n_head = np.where(u < 0.4, 0, np.where(u < 0.7, 1, 2))
    part_lens = np.zeros((n_items, 4), dtype=np.int64)
    part_lens[:, 0] = np.where(n_head == 2, 2, 0)
    part_lens[:, 1] = np.where(n_head >= 1, 2, 0)
    part_lens[:, 2] = rng.integers(3, 10, size=n_items)
    part_lens[:, 3] = rng.integers(1, 6, size=n_items)
    total = int(part_lens.sum())
    body_cdf = np.cumsum(zipf_probs(n_vocab)); body_cdf[-1] = 1.0
    head_cdf = np.cumsum(zipf_probs(400)); head_cdf[-1] = 1.0
    is_head = np.repeat(np.tile(np.array([True, True, False, False]), n_items), part_lens.reshape(-1))
    r = rng.random(total)
    flat = np.where(is_head, np.searchsorted(head_cdf, r, side="right"),
                    np.searchsorted(body_cdf, r, side="right")).astype(np.uint32)
    return part_lens, flat


def definition_level_sets(n_items: int, seed: int = SEED_DEFS, n_vocab: int = 20000):
    """cfg4 right-hand side at scale (GECCO-style ``Term = [category, parameter, choice]``) as
    integer word ids; same return convention as :func:`term_level_sets` with three parts:
    category 1-2 words from the 200 most frequent, parameter 1-5 words, choice absent 40 % or
    1-3 words."""
    rng = np.random.default_rng(seed)
    part_lens = np.zeros((n_items, 3), dtype=np.int64)
    part_lens[:, 0] = rng.integers(1, 3, size=n_items)
    part_lens[:, 1] = rng.integers(1, 6, size=n_items)
    part_lens[:, 2] = np.where(rng.random(n_items) < 0.6, rng.integers(1, 4, size=n_items), 0)
    total = int(part_lens.sum())
    body_cdf = np.cumsum(zipf_probs(n_vocab)); body_cdf[-1] = 1.0
    cat_cdf = np.cumsum(zipf_probs(200)); cat_cdf[-1] = 1.0
    is_cat = np.repeat(np.tile(np.array([True, False, False]), n_items), part_lens.reshape(-1))
    r = rng.random(total)
    flat = np.where(is_cat, np.searchsorted(cat_cdf, r, side="right"),
                    np.searchsorted(body_cdf, r, side="right")).astype(np.uint32)
    return part_lens, flat
