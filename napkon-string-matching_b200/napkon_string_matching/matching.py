"""``match(config, use_cache)``: the step loop of the CLI
(/root/reference/napkon_string_matching/matching.py:18-38)."""
from __future__ import annotations

import logging
from typing import Dict

from napkon_string_matching.matcher import Matcher

CONFIG_FIELD_PREPARE = "prepare"
CONFIG_FIELD_MATCHING = "matching"
CONFIG_FIELD_STEPS = "steps"

logger = logging.getLogger(__name__)

_STEPS = {
    "variables": Matcher.match_questionnaires_variables,
    "gecco": Matcher.match_gecco_with_questionnaires,
    "questionnaires": Matcher.match_questionnaires,
}


def match(config: Dict, use_cache=True, matcher: Matcher | None = None) -> Matcher:
    matcher = matcher or create_matcher(config, use_cache)
    # the comparisons of all steps are independent: one batch for the multi-GPU scheduler
    matcher.match_steps([step for step in config[CONFIG_FIELD_STEPS] if step in _STEPS])
    matcher.print_analysis()
    matcher.write_results()
    return matcher


def create_matcher(config: Dict, use_cache=True) -> Matcher:
    # token enrichment (prepare/match_preparator.py) needs a MeSH data base and is not part of
    # the comparison path: no preparator is constructed here
    return Matcher(None, config, use_cache=use_cache)
