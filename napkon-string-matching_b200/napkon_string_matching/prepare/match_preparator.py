"""``MatchPreparator.add_tokens``: adds ``TokenIds`` / ``Tokens`` / ``TokenMatch`` columns from the
terminology matches of every item's term
(/root/reference/napkon_string_matching/prepare/match_preparator.py:34-74).  The reference forks a
process pool and scores term by term; here the whole column goes through one GPU launch."""
from __future__ import annotations

import logging

from napkon_string_matching.terminology.provider import TerminologyProvider
from napkon_string_matching.types.comparable_data import ComparableData

CONFIG_FIELD_TERMINOLOGY = "terminology"

logger = logging.getLogger(__name__)


class MatchPreparator:
    def __init__(self, config: dict, term_requests=None, heading_requests=None):
        self.config = config
        self.term_requests = term_requests
        self.heading_requests = heading_requests
        self.terminology_provider = TerminologyProvider(self.config[CONFIG_FIELD_TERMINOLOGY])

    def add_tokens(self, cs: ComparableData, score_threshold: float = 0.1, verbose: bool = True,
                   timeout=10):
        if not self.terminology_provider.initialized:
            self.terminology_provider.initialize()
        if not self.terminology_provider.initialized:
            raise RuntimeError("'terms' and/or 'headings' not initialized")
        logger.info("add tokens...")
        results = self.terminology_provider.get_matches_many(list(cs.term), score_threshold)
        unpacked = [tuple(zip(*entry)) if entry else (None, None, None) for entry in results]
        cs.token_ids = [ids if ids else None for ids, *_ in unpacked]
        cs.tokens = [tokens if tokens else None for _, tokens, *_ in unpacked]
        cs.token_match = results
        logger.info("...done")
