"""GPU suite: the drop-in Python API (gen_comparable / compare / score functions / Matcher) on the
CUDA engine against the reference-generated golden vectors."""
import json

import numpy as np
import pytest

import golden_cases
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

SCALARS = json.loads((GOLDEN / "scalar_cases.json").read_text(encoding="utf-8"))


@pytest.mark.parametrize("name", golden_cases.FRAME_CASES)
def test_gen_comparable_matches_reference_run(cuda_engine, name):
    golden_cases.check_case(name)


def _check_scalar(fn, case):
    want = case["result"]
    if want.startswith("raises:"):
        with pytest.raises(Exception) as ei:
            fn(case["left"], case["right"])
        assert type(ei.value).__name__ == want.split(":")[1]
    else:
        assert float(fn(case["left"], case["right"])).hex() == want


def test_scalar_score_functions_run_on_the_gpu(cuda_engine):
    from napkon_string_matching.compare import score_functions as sf

    before = cuda_engine.launches
    for case in SCALARS["jaccard_hand"]["cases"]:
        _check_scalar(sf.intersection_vs_union, case)
    for case in SCALARS["fuzzy_hand"]["cases"]:
        _check_scalar(sf.fuzzy_match, case)
    assert cuda_engine.launches - before >= len(SCALARS["jaccard_hand"]["cases"])
    assert sf.join_sorted(["beta", "Alpha"]) == "Alpha beta"


def test_compare_terms_with_gpu_score_functions(cuda_engine):
    from napkon_string_matching.compare import score_functions as sf
    from napkon_string_matching.types.comparable_data import ComparableData

    for func, fn in (("jaccard", sf.intersection_vs_union), ("fuzzy", sf.fuzzy_match)):
        for case in SCALARS[f"compare_terms_{func}"]["cases"]:
            _check_scalar(lambda l, r: ComparableData.compare_terms(l, r, fn), case)


def test_one_against_many_matches_the_scalar_calls(cuda_engine):
    from napkon_string_matching.compare import score_functions as sf

    synonyms = ["Dialyse", "Renal Dialysis", "Sonstiges", "Hatte Sie Dialyse?", "", "dialyse nach entlassung"]
    many = sf.fuzzy_match_many("Dialyse nach Entlassung", synonyms)
    assert [float(x).hex() for x in many] == \
        [float(sf.fuzzy_match("Dialyse nach Entlassung", s)).hex() for s in synonyms]
    assert many[-1] == 1.0 and many[4] == 0.0
    sets = [["a", "b"], ["b"], ["x", "y", "z"], ["a", "b", "c", "d"]]
    many = sf.intersection_vs_union_many(["a", "b", "c"], sets)
    assert list(many) == [2 / 3, 1 / 3, 0.0, 3 / 4]


def test_terminology_get_matches_on_the_gpu(cuda_engine):
    import terminology_cases

    terminology_cases.check_all()
    terminology_cases.check_add_tokens()


def test_scheduler_threads_on_the_gpu(engine, tmp_path):
    """gpu/scheduler.py: comparisons dealt out to worker threads, each with its own Engine (one
    per visible GPU; two engines on the one GPU when there is only one), give the frames of the
    sequential loop."""
    import torch

    import test_scheduler as ts
    from napkon_string_matching.gpu import scheduler
    from napkon_string_matching.gpu.engine import Engine

    from napkon_string_matching.gpu.engine import use_engine

    seq, _ = ts._matcher(tmp_path, "seq")
    with use_engine(engine):
        for task in seq._variable_tasks() + seq._gecco_tasks() + seq._questionnaire_tasks():
            seq.results[task.name] = task.run()
    n_dev = torch.cuda.device_count()
    engines = [Engine(k % n_dev) for k in range(max(2, n_dev))]
    par, _ = ts._matcher(tmp_path, "par")
    tasks = par._variable_tasks() + par._gecco_tasks() + par._questionnaire_tasks()
    for task, result in zip(tasks, scheduler.run_comparisons(tasks, engines=engines)):
        par.results[task.name] = result
    assert all(e.launches > 0 for e in engines)
    a, b = ts._frames(seq), ts._frames(par)
    assert list(a) == list(b)
    for name in a:
        assert a[name].equals(b[name]), name


def test_affinity_binds_to_gpu_local_cpus_or_leaves_everything_alone():
    import os

    from napkon_string_matching.gpu import affinity

    before = os.sched_getaffinity(0)
    try:
        got = affinity.bind_to_gpu(0)
        now = os.sched_getaffinity(0)
        if got is None:
            assert now == before
        else:
            assert set(got) == now and now < before and now <= set(affinity.gpu_local_cpus(0))
    finally:
        os.sched_setaffinity(0, before)
