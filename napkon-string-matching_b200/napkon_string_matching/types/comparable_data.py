"""
``ComparableData``: cohort / definition items that can be compared with one another.

Same public surface as the compare half of
/root/reference/napkon_string_matching/types/comparable_data.py (compare :69-128, gen_comparable
:133-246, compare_terms :248-265, gen_comp_value :283-285, tokenize :287-299 and the module-level
filters :464-574).  What changed is the inside of ``gen_comparable``: the reference materialises
the N_l x N_r cross product as a DataFrame and calls ``compare_terms`` once per row in Python;
here the items are tokenised and packed on the host, every pair is scored on the GPU
(gpu/engine.py -> libnsm_b200.so), and the result frame is gathered from the kept
``(left, right, score)`` records only.  The N x N frame is never built.
"""
from __future__ import annotations

import logging
import os
from enum import Enum
from pathlib import Path
from typing import Dict, List, Tuple

import pandas as pd

import napkon_string_matching.compare.score_functions
from napkon_string_matching.text import tokenize as _tok
from napkon_string_matching.types.comparable import (COLUMN_NAMES, QUESTION_OUTPUT, Columns,  # noqa: F401
                                                     Comparable)
from napkon_string_matching.types.data import Data, gen_hash
from napkon_string_matching.types.mapping import Mapping

PREPARE_REMOVE_SYMBOLS = _tok.PREPARE_REMOVE_SYMBOLS
CACHE_FILE_PATTERN = "compared__score_{}.json"
COMP_COLUMN = "Compare"

logger = logging.getLogger(__name__)
flatten_list = _tok.flatten_list


def _sharded_world():
    """(rank, world) when the ranks of a torch.distributed job share ONE comparison by row
    blocks (gpu/distributed.py); (0, 1) for a single process or when the scheduler has dealt out
    whole comparisons."""
    try:
        from napkon_string_matching.gpu import distributed
    except ImportError:  # torch missing: single process
        return 0, 1
    return distributed._world()


def _cache_decision(local_hit: bool):
    """``(use the cache, this process writes the cache file)``.  With row-sharded ranks every rank
    runs ``compare``; they must take the same branch (the miss branch holds a collective), so
    rank 0's view of the cache directory decides for all and rank 0 alone writes."""
    rank, world = _sharded_world()
    if world == 1:
        return local_hit, True
    import torch.distributed as dist

    from napkon_string_matching.gpu import distributed

    box = [bool(local_hit)]
    dist.broadcast_object_list(box, src=0, group=distributed.host_group())
    return bool(box[0]), rank == 0


def _share_cached(result):
    """Rank 0's cached result for every row-sharded rank (the others may not see the file)."""
    rank, world = _sharded_world()
    if world == 1:
        return result
    import torch.distributed as dist

    from napkon_string_matching.gpu import distributed

    box = [result if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=distributed.host_group())
    return box[0]


class ComparableColumns(Enum):
    TERM = "Term"
    TOKENS = "Tokens"
    TOKEN_IDS = "TokenIds"
    TOKEN_MATCH = "TokenMatch"
    MATCHES = "Matches"
    IDENTIFIER = "Identifier"


class ComparableData(Data):
    __columns__ = list(ComparableColumns)
    __column_mapping__: Dict[str, str] = {}
    __category_column__ = "Category"

    # ------------------------------------------------------------------ compare (cache wrapper)
    def _hash_compare_args(self, other, *args, **kwargs) -> str:
        strings = [self.to_csv(), other.to_csv()]
        strings += [str(arg) for arg in args]
        strings += [str(item) for item in kwargs.items()]
        return gen_hash("".join(strings))

    def compare(
        self,
        other,
        existing_mappings_whitelist: Mapping,
        existing_mappings_blacklist: Mapping,
        compare_column: str,
        score_threshold: float = 0.1,
        cached: bool = True,
        cache_threshold: float = None,
        cache_dir: str | Path | None = None,
        identifier_column_left: str | None = None,
        identifier_column_right: str | None = None,
        *args,
        **kwargs,
    ) -> Comparable:
        """Scores every item of ``self`` against every item of ``other`` (or reads the cached
        result), keeps ``MatchScore >= score_threshold`` and sorts by score, descending."""
        df_hash = self._hash_compare_args(
            other=other,
            existing_mappings_whitelist=existing_mappings_whitelist,
            existing_mappings_blacklist=existing_mappings_blacklist,
            compare_column=compare_column,
            cache_threshold=cache_threshold,
        )
        cache_file = Path(cache_dir if cache_dir else "cache") / CACHE_FILE_PATTERN.format(df_hash)
        # row-sharded over torch.distributed ranks (gpu/distributed.py): every rank is in here, so
        # rank 0 alone decides hit or miss (ranks may see different cache directories) and writes
        use_cache, writer = _cache_decision(bool(cached) and cache_file.exists())
        if use_cache:
            logger.info("using cached result")
            result = Comparable.read_json(cache_file) if cache_file.exists() else None
            result = _share_cached(result)
        else:
            result = self.gen_comparable(
                other,
                existing_mappings_whitelist=existing_mappings_whitelist,
                existing_mappings_blacklist=existing_mappings_blacklist,
                score_threshold=cache_threshold if cache_threshold else score_threshold,
                compare_column=compare_column,
                identifier_column_left=identifier_column_left,
                identifier_column_right=identifier_column_right,
                *args,
                **kwargs,
            )
            if writer:
                cache_file.parent.mkdir(parents=True, exist_ok=True)
                logger.info("write cache to file")
                # a reader never sees a half-written file: write aside, then rename
                from napkon_string_matching.gpu.stages import stage

                tmp = cache_file.with_name(f"{cache_file.name}.{os.getpid()}.tmp")
                with stage("cache JSON"):
                    result.write_json(tmp)
                    os.replace(tmp, cache_file)

        # outside of the caching, so one cache serves several thresholds
        result = result[result.match_score >= score_threshold]
        logger.info("got %i filtered entries", len(result))
        result.sort_by_score()
        return result

    def map_for_comparable(self) -> pd.DataFrame:
        return self._data.rename(columns=self.__column_mapping__)

    # ------------------------------------------------------------------ the hot path
    def gen_comparable(
        self,
        right,
        existing_mappings_whitelist: Mapping,
        existing_mappings_blacklist: Mapping,
        score_func: str,
        compare_column: str,
        category_column: str = "Category",
        score_threshold: float = 0.1,
        left_name: str = None,
        right_name: str = None,
        filter_categories: bool = False,
        identifier_column_left: str | None = None,
        identifier_column_right: str | None = None,
        *args,
        **kwargs,
    ) -> Comparable:
        from napkon_string_matching.gpu import pairing
        from napkon_string_matching.gpu.stages import stage

        # unknown names fail exactly like the reference's getattr (comparable_data.py:150)
        getattr(napkon_string_matching.compare.score_functions, score_func)

        left = self.dropna(subset=[compare_column])
        right = right.dropna(subset=[compare_column])
        logger.info("comparing number of items %i left, %i right, potential %s comparisons",
                    len(left), len(right), "{:,}".format(len(left) * len(right)))

        remove_existing_mappings(left, right, left_name, right_name, existing_mappings_whitelist)
        logger.info("after removing existing whitelisted mappings: %i left, %i right",
                    len(left), len(right))

        left_df, right_df = left.map_for_comparable(), right.map_for_comparable()
        term = ComparableColumns.TERM.value
        with stage("tokenise (gen_comp_value)"):
            left_levels = [self.gen_comp_value(item) for item in left_df[compare_column]]
            right_levels = [self.gen_comp_value(item) for item in right_df[compare_column]]
        left_df = left_df.assign(**{QUESTION_OUTPUT: [":".join(flatten_list(t)) for t in left_df[term]]})
        right_df = right_df.assign(**{QUESTION_OUTPUT: [":".join(flatten_list(t)) for t in right_df[term]]})

        left_prefix, right_prefix = left_name.title(), right_name.title()
        id_left = identifier_column_left or Columns.IDENTIFIER.value
        id_right = identifier_column_right or Columns.IDENTIFIER.value

        # black list (Q8): pure exclusion of (left id, right id) pairs
        blocked = set(flatten_mapping(left_name, right_name, existing_mappings_blacklist))
        skip_pair = None
        if blocked:
            lids, rids = list(left_df[id_left]), list(right_df[id_right])
            skip_pair = lambda li, ri: (lids[li], rids[ri]) in blocked  # noqa: E731

        categories = None
        if filter_categories and len(left_df) and len(right_df):
            categories = pairing.category_masks(left_df[category_column], right_df[category_column])

        logger.info("calculate score")
        records = pairing.score_all_pairs(left_levels, right_levels, score_func, score_threshold,
                                          categories=categories, skip_pair=skip_pair)
        if blocked and len(records):
            keep = pairing.not_blocked(records, left_df[id_left], right_df[id_right], blocked)
            logger.info("removed %i black-listed pairs", int((~keep).sum()))
            records = records[keep]

        with stage("result frame"):
            frame = pairing.result_frame(records, left_df, right_df, left_prefix, right_prefix)
        logger.info("got %s entries", "{:,}".format(len(frame)))
        return Comparable(data=frame, left_name=left_prefix, right_name=right_prefix)

    @classmethod
    def compare_terms(cls, left: List[str], right: List[str], score_func) -> float:
        """Weighted multi-level score of ONE pair with an arbitrary callable (Q1).  The all-pairs
        path evaluates the same schedule inside the CUDA kernels."""
        score, factor = 0, 1
        last_left, last_right = len(left) - 1, len(right) - 1
        for i in range(1, max(len(left), len(right)) + 1):
            factor /= 2
            score += score_func(left[min(i, last_left)], right[min(i, last_right)]) * factor
        return score

    def remove_existing_mappings(self, existing_mappings) -> None:
        drop = set(existing_mappings)
        self._data = self._data[[v not in drop for v in self[ComparableColumns.IDENTIFIER.value]]]

    def add_terms(self, language: str = "german"):
        raise NotImplementedError()

    @staticmethod
    def gen_term(*items: str) -> List[str]:
        return [item for item in items if item]

    @classmethod
    def gen_comp_value(cls, items: List[str]) -> List[str]:
        if cls.tokenize is ComparableData.tokenize:   # not overridden: the incremental builder
            return _tok.gen_comp_value(items)
        return [cls.tokenize(items[-i:]) for i in range(1, len(items) + 1)]

    @staticmethod
    def tokenize(parts: List[str], language: str = "german") -> str:
        return _tok.tokenize(parts, language)

    def filter(self, filter_column: str, filter_prefix: str):
        keep = [entry.startswith(filter_prefix) if pd.notna(entry) else True
                for entry in self[filter_column]]
        self._data = self._data[keep]

    def get_existing_mapping_ids(self, group_name: str, mappings: Mapping):
        own = set(self[Columns.IDENTIFIER.value])
        return [id for id, identifiers in mappings.filter_by_group(group_name).items()
                if own.intersection(identifiers)]


# ----------------------------------------------------------------------------------------------
# module-level helpers with the reference's names
# ----------------------------------------------------------------------------------------------
def categories_matching(df: pd.DataFrame, column_left: str, column_right: str) -> pd.DataFrame:
    """Row filter with the predicate chosen from the types in row 0 (Q9); kept for callers that
    already hold a pair frame.  gen_comparable applies the same predicate as bit masks."""
    from napkon_string_matching.gpu import pairing

    if not len(df):
        return df
    keep = pairing.category_keep_rows(df[column_left], df[column_right])
    return df[keep]


def remove_existing_mappings(left: ComparableData, right: ComparableData, left_name: str,
                             right_name: str, existing_mappings: Mapping):
    """Drops, from both sides, every identifier of a white-list group that has members on both
    sides.  A group lacking one of the two names aborts the whole removal (Q7)."""
    try:
        left_ids = left.get_existing_mapping_ids(left_name, existing_mappings)
        right_ids = right.get_existing_mapping_ids(right_name, existing_mappings)
    except KeyError:
        return
    shared = existing_mappings.get_filtered(list(set(left_ids).intersection(right_ids)))
    left.remove_existing_mappings(get_identifiers_from_mapping(shared, left_name))
    right.remove_existing_mappings(get_identifiers_from_mapping(shared, right_name))


def get_identifiers_from_mapping(mappings: Mapping, group: str) -> List[str]:
    return [identifier for groups in mappings.values() for identifier in groups[group]]


def remove_existing_mapping_from_df(df: pd.DataFrame, left_name: str, right_name: str,
                                    left_prefix: str, right_prefix: str, existing_mappings: Mapping,
                                    identifier_column_left: str | None = None,
                                    identifier_column_right: str | None = None):
    blocked = set(flatten_mapping(left_name, right_name, existing_mappings))
    id_left = left_prefix + (identifier_column_left or Columns.IDENTIFIER.value)
    id_right = right_prefix + (identifier_column_right or Columns.IDENTIFIER.value)
    return df[[pair not in blocked for pair in zip(df[id_left], df[id_right])]]


def flatten_mapping(left_group: str, right_group: str, mapping: Mapping) -> List[Tuple[str, str]]:
    return [(l, r)
            for lefts, rights in mapping.get_all_mapping_for_groups(left_group, right_group)
            for l in lefts for r in rights]
