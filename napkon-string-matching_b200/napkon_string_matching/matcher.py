"""
``Matcher``: runs the comparison steps of one configuration — cohort variables, GECCO vs cohorts,
cohort vs cohort — and collects the results
(same constructor and ``match_*`` / ``print_analysis`` / ``write_results`` methods as
/root/reference/napkon_string_matching/matcher.py:53-337).

The reference's ``__init__`` also parses Excel sheets, queries FHIR Simplifier and a MeSH data
base; that ETL is outside the comparison path.  Here the inputs are the reference's own JSON
caches (``input.prepared`` maps a name to a prepared-items JSON file, see
ComparableData.prepare's ``*_prepared.json``) or objects handed in directly.
"""
from __future__ import annotations

import logging
from itertools import combinations
from pathlib import Path
from string import Template
from typing import Any, Dict, List

from napkon_string_matching.gpu.scheduler import ComparisonTask, run_comparisons
from napkon_string_matching.types.comparable import ComparisonResults
from napkon_string_matching.types.gecco_definition import GeccoDefinition, KdsDefinition
from napkon_string_matching.types.mapping import Mapping
from napkon_string_matching.types.questionnaire import Questionnaire

CONFIG_FIELD_MATCHING = "matching"
CONFIG_VARIABLE_THRESHOLD = "variable_score_threshold"
CONFIG_INPUT = "input"
CONFIG_INPUT_BASE_DIR = "base_dir"
CONFIG_FIELD_MAPPINGS = "mappings"
CONFIG_PREPARED = "prepared"
CONFIG_OUTPUT_DIR = "output_dir"
CONFIG_CACHE_DIR = "cache_dir"
RESULTS_FILE_PATTERN = "result_{score_threshold}_{compare_column}_{score_func}.xlsx"

logger = logging.getLogger(__name__)


class Matcher:
    def __init__(self, preparator, config: Dict, use_cache=True, *, gecco: GeccoDefinition = None,
                 kds: KdsDefinition = None, questionnaires: Dict[str, Questionnaire] = None) -> None:
        self.preparator = preparator
        self.config = config
        self.use_cache = use_cache
        self.input_config: Dict | None = config.get(CONFIG_INPUT)
        self.input_dir = self._input_config(CONFIG_INPUT_BASE_DIR)
        self.cache_dir = config.get(CONFIG_CACHE_DIR)
        self.gecco = gecco
        self.kds = kds
        self.questionnaires: Dict[str, Questionnaire] = dict(questionnaires or {})
        self.mappings_whitelist = Mapping()
        self.mappings_blacklist = Mapping()
        self.results = ComparisonResults()
        self._init_mappings()
        self._init_prepared()

    # ---- inputs -----------------------------------------------------------------------
    def _input_config(self, field_name: str) -> Any:
        return self.input_config.get(field_name) if self.input_config else None

    def _expand_path(self, path: str) -> str:
        return Template(path).substitute(input_base_dir=self.input_dir)

    def _init_mappings(self) -> None:
        folder = self._input_config(CONFIG_FIELD_MAPPINGS)
        if not folder:
            return
        folder = Path(self._expand_path(folder))
        for file in sorted(folder.glob("whitelist/*.json")):
            self.mappings_whitelist.update(Mapping.read_json(file))
        for file in sorted(folder.glob("blacklist/*.json")):
            self.mappings_blacklist.update(Mapping.read_json(file))

    def _init_prepared(self) -> None:
        prepared: Dict[str, str] = self._input_config(CONFIG_PREPARED) or {}
        for name, file in prepared.items():
            path = self._expand_path(file)
            if name == "gecco":
                self.gecco = GeccoDefinition.read_json(path)
            elif name == "kds":
                self.kds = KdsDefinition.read_json(path)
            else:
                self.questionnaires[name] = Questionnaire.read_json(path)

    def clear_results(self) -> None:
        self.results = ComparisonResults()

    # ---- steps ------------------------------------------------------------------------
    # Every step first lists its comparisons (same order, names and arguments as the reference's
    # loops, matcher.py:228-284) and then hands the list to the scheduler, which spreads whole
    # comparisons over the GPUs of the box and returns the results in list order.
    def _gecco_tasks(self) -> List[ComparisonTask]:
        return [ComparisonTask(
            f"gecco vs {name}", self.gecco, questionnaire,
            dict(existing_mappings_whitelist=self.mappings_whitelist,
                 existing_mappings_blacklist=self.mappings_blacklist,
                 left_name="gecco", right_name=name, cache_dir=self.cache_dir,
                 **self.config[CONFIG_FIELD_MATCHING]))
            for name, questionnaire in self.questionnaires.items()]

    def _questionnaire_tasks(self, prefix: str = None, **kwargs) -> List[ComparisonTask]:
        """Every unordered cohort pair once, the lower-case-smaller name on the left."""
        names = sorted(self.questionnaires, key=str.lower)
        return [ComparisonTask(
            f"{prefix if prefix else ''}{name_first} vs {name_second}",
            self.questionnaires[name_first], self.questionnaires[name_second],
            dict(existing_mappings_whitelist=self.mappings_whitelist,
                 existing_mappings_blacklist=self.mappings_blacklist,
                 left_name=name_first, right_name=name_second, cache_dir=self.cache_dir,
                 **{**self.config[CONFIG_FIELD_MATCHING], **kwargs}))
            for name_first, name_second in combinations(names, 2)]

    def _variable_tasks(self) -> List[ComparisonTask]:
        return self._questionnaire_tasks(
            prefix="var_", compare_column="Variable",
            score_threshold=self.config[CONFIG_FIELD_MATCHING][CONFIG_VARIABLE_THRESHOLD])

    def _run(self, tasks: List[ComparisonTask]) -> None:
        for task in tasks:
            logger.info("compare %s", task.name)
        for task, result in zip(tasks, run_comparisons(tasks)):
            self.results[task.name] = result

    def match_gecco_with_questionnaires(self) -> None:
        self._run(self._gecco_tasks())

    def match_questionnaires(self, prefix: str = None, *args, **kwargs) -> None:
        self._run(self._questionnaire_tasks(prefix, **kwargs))

    def match_questionnaires_variables(self) -> None:
        self._run(self._variable_tasks())

    def match_steps(self, steps) -> None:
        """All comparisons of the listed steps ("variables", "gecco", "questionnaires") as ONE
        batch for the scheduler, so that e.g. nine comparisons keep eight GPUs busy; results are
        stored in the order the sequential step loop would have produced them."""
        lists = {"variables": self._variable_tasks, "gecco": self._gecco_tasks,
                 "questionnaires": self._questionnaire_tasks}
        self._run([task for step in steps if step in lists for task in lists[step]()])

    # ---- reporting --------------------------------------------------------------------
    def _analyse(self) -> Dict[str, Dict[str, str]]:
        """Distinct matched variables per comparison, overall and for ``gec_`` variables."""
        prefix = "gec_"
        result = {}
        for name, comp in self.results.items():
            if comp.empty:
                continue
            left_gecco = comp[[prefix in v for v in comp.variable]]
            right_gecco = comp[[prefix in v for v in comp.match_variable]]
            result[name] = {
                "matched": "{}/{}".format(comp.variable.nunique(), comp.match_variable.nunique()),
                "gecco": "{}/{}".format(left_gecco.variable.nunique(),
                                        right_gecco.match_variable.nunique()),
            }
        return result

    @staticmethod
    def _reports() -> bool:
        """Under torch.distributed only rank 0 reports: with row-sharded comparisons it is the
        rank that holds the complete results (gpu/distributed.py:default_gather)."""
        try:
            import torch.distributed as dist
        except ImportError:
            return True
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0

    def print_analysis(self) -> None:
        if not self._reports():
            return
        for name, item in self._analyse().items():
            logger.info("%s\t%s", name, "\t".join(f"{k}: {v}" for k, v in item.items()))

    def write_results(self) -> None:
        if not self._reports():
            return
        matching = self.config[CONFIG_FIELD_MATCHING]
        output_file = RESULTS_FILE_PATTERN.format(
            **{**matching, "score_func": matching["score_func"].replace("_", "-")})
        if output_dir := self.config.get(CONFIG_OUTPUT_DIR):
            output_file = f"{output_dir}/{output_file}"
        self.results.write_excel(output_file)
