"""
Token enrichment lookups: every synonym of a terminology (MeSH) scored against an item's term
with ``fuzzy_match``, kept at or above a threshold, best synonym per heading id first
(/root/reference/napkon_string_matching/terminology/mesh.py:192-220, provider.py:44-55).

The reference scores one term at a time with ``np.vectorize(fuzzy_match)`` over a deep copy of
the synonym frame, inside a fork pool (prepare/match_preparator.py:55-67).  Here all terms of a
cohort are scored against all synonyms in one all-pairs launch of the fuzzy kernel
(:meth:`TerminologyProvider.get_matches_many`): at the configured threshold (config.yml
``tokens.score_threshold: 0.85``) the kernel's distance bound proves nearly every term x synonym
pair below the threshold without computing its LCS.  The post-processing (sort by score,
``drop_duplicates("Id")``) is one vectorised pass over the kept pairs of all terms.

Reading MeSH from Postgres (mesh.py:122-168) is ETL and out of scope: the synonym / heading
frames are handed in (the reference's own tests inject them the same way, test_mesh.py:19-24).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import pandas as pd

TERMINOLOGY_COLUMN_TERM = "Term"
TERMINOLOGY_COLUMN_ID = "Id"
TERMINOLOGY_COLUMN_SCORE = "Score"
CONFIG_FIELD_MESH = "mesh"

Match = Tuple[str, str, float]


class MeshProvider:
    """Holds the ``Id`` / ``Term`` frames of one terminology."""

    def __init__(self, config=None, synonyms: pd.DataFrame | None = None,
                 headings: pd.DataFrame | None = None) -> None:
        self.config = config
        self._synonyms = synonyms
        self._headings = headings

    @property
    def initialized(self) -> bool:
        return self._synonyms is not None and self._headings is not None

    def initialize(self) -> None:
        if not self.initialized:
            raise NotImplementedError(
                "loading MeSH from the data base is outside the comparison path: set the "
                "`_synonyms` / `_headings` frames (columns Id, Term)")

    @property
    def headings(self) -> pd.DataFrame:
        return self._headings

    @property
    def synonyms(self) -> pd.DataFrame:
        return self._synonyms

    def get_matches(self, term: List[str], score_threshold: float = 0.1) -> List[Match]:
        return self.get_matches_many([term], score_threshold)[0]

    def get_matches_many(self, terms: Sequence[List[str]], score_threshold: float = 0.1
                         ) -> List[List[Match]]:
        """``[(Id, Term, Score), ...]`` per term: synonyms with ``fuzzy_match(synonym,
        " ".join(term)) >= score_threshold``, ordered by falling score, one per ``Id``."""
        from napkon_string_matching.gpu import pack
        from napkon_string_matching.gpu.engine import default_engine
        from napkon_string_matching.text.process import default_process

        syn = self.synonyms
        ids = syn[TERMINOLOGY_COLUMN_ID].to_numpy()
        syn_terms = syn[TERMINOLOGY_COLUMN_TERM].to_numpy()
        if len(terms) == 0:
            return []
        if len(syn) == 0:
            return [[] for _ in terms]
        engine = default_engine()
        # " ".join(term) is what the reference scores (a str term is joined per character)
        from napkon_string_matching.gpu.device_pack import PackUnsupported

        packer = getattr(engine, "string_packer", None)
        try:   # default_process and the packed arrays on the GPU (csrc/pack_strings.cu)
            if packer is None:
                raise PackUnsupported("no device string packer")
            dq, ds = packer.pack([[[" ".join(t)] for t in terms], [[s] for s in syn_terms]])
        except PackUnsupported:
            pq, ps = pack.pack_strings([[default_process(" ".join(t))] for t in terms],
                                       [[default_process(s)] for s in syn_terms])
            dq, ds = engine.upload(pq), engine.upload(ps)
        rec = engine.all_pairs(dq, ds, score_threshold, flat=True)
        # per term: best score first (ties keep the synonym frame's order), one row per Id —
        # sort_values(Score, descending) + drop_duplicates("Id") of mesh.py:213-218, for all terms at
        # once: after the sort the FIRST record of every (term, Id) is the one that survives
        order = np.lexsort((rec["right"], -rec["score"], rec["left"]))
        rec = rec[order]
        gid, _ = pd.factorize(ids)
        key = rec["left"].astype(np.int64) * (int(gid.max(initial=0)) + 1) + gid[rec["right"]]
        _, first = np.unique(key, return_index=True)
        rec = rec[np.sort(first)]
        out: List[List[Match]] = [[] for _ in terms]
        bounds = np.searchsorted(rec["left"], np.arange(len(terms) + 1))
        r_ids, r_terms, r_scores = ids[rec["right"]], syn_terms[rec["right"]], rec["score"].tolist()
        for t in range(len(terms)):
            lo, hi = int(bounds[t]), int(bounds[t + 1])
            if hi > lo:
                out[t] = list(zip(r_ids[lo:hi].tolist(), r_terms[lo:hi].tolist(), r_scores[lo:hi]))
        return out


class TerminologyProvider:
    """Combines the providers of the configured terminologies (only MeSH exists)."""

    def __init__(self, config=None, providers: List[MeshProvider] | None = None) -> None:
        self.config = config
        mesh_config = config.get(CONFIG_FIELD_MESH) if isinstance(config, dict) else None
        self.providers: List[MeshProvider] = providers if providers is not None \
            else [MeshProvider(mesh_config)]

    @property
    def initialized(self) -> bool:
        return all(provider.initialized for provider in self.providers)

    def initialize(self) -> None:
        for provider in self.providers:
            provider.initialize()

    @property
    def headings(self) -> pd.DataFrame:
        return pd.concat([provider.headings for provider in self.providers])

    @property
    def synonyms(self) -> pd.DataFrame:
        return pd.concat([provider.synonyms for provider in self.providers])

    def get_matches(self, term: List[str], score_threshold: float = 0.1) -> List[Match] | None:
        results: List[Match] = []
        for provider in self.providers:
            results += provider.get_matches(term, score_threshold)
        return results if results else None

    def get_matches_many(self, terms: Sequence[List[str]], score_threshold: float = 0.1
                         ) -> List[List[Match] | None]:
        merged: List[List[Match]] = [[] for _ in terms]
        for provider in self.providers:
            for acc, part in zip(merged, provider.get_matches_many(terms, score_threshold)):
                acc += part
        return [m if m else None for m in merged]
