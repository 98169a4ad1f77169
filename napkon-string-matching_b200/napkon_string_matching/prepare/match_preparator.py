"""Token enrichment of a whole cohort in one GPU launch.

``MatchPreparator.add_tokens`` fills the ``TokenIds`` / ``Tokens`` / ``TokenMatch`` columns from
the terminology matches of every item's term.  The reference
(/root/reference/napkon_string_matching/prepare/match_preparator.py:34-74) forks a process pool
and scores one term at a time; here ``TerminologyProvider.get_matches_many`` scores all terms
against all synonyms at once and only the kept pairs come back to the host."""
from __future__ import annotations

import logging
from typing import List, Optional, Sequence, Tuple

from napkon_string_matching.terminology.provider import Match, TerminologyProvider
from napkon_string_matching.types.comparable_data import ComparableData

CONFIG_FIELD_TERMINOLOGY = "terminology"

logger = logging.getLogger(__name__)


def _columns(matches: Sequence[Optional[List[Match]]]) -> Tuple[list, list]:
    """(ids per item, terms per item); an item without matches gets None in both."""
    ids, terms = [], []
    for entry in matches:
        if entry:
            entry_ids, entry_terms, _ = zip(*entry)
            ids.append(entry_ids)
            terms.append(entry_terms)
        else:
            ids.append(None)
            terms.append(None)
    return ids, terms


class MatchPreparator:
    """Prepares data for the matching process (terminology lookups)."""

    def __init__(self, config: dict, term_requests=None, heading_requests=None):
        self.config = config
        # table requests describe the MeSH data base layout; kept for signature compatibility
        self.term_requests, self.heading_requests = term_requests, heading_requests
        self.terminology_provider = TerminologyProvider(config[CONFIG_FIELD_TERMINOLOGY])

    def add_tokens(self, cs: ComparableData, score_threshold: float = 0.1, verbose: bool = True,
                   timeout=10):
        """`verbose` / `timeout` belonged to the reference's pool of futures; unused here."""
        provider = self.terminology_provider
        if not provider.initialized:
            provider.initialize()
        if not provider.initialized:
            raise RuntimeError("'terms' and/or 'headings' not initialized")

        logger.info("add tokens for %i items", len(cs))
        matches = provider.get_matches_many(list(cs.term), score_threshold)
        cs.token_ids, cs.tokens = _columns(matches)
        cs.token_match = matches
