"""CPU suite: concurrent scheduling of the comparisons of a matching run — the
LPT plan, the one-thread-per-GPU mode and the one-process-per-GPU mode (gloo, world size 2), with
the C oracle standing in for the CUDA engine (conftest.OracleEngine)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import PKG, ROOT, OracleEngine
from napkon_string_matching.gpu import scheduler


def test_plan_is_lpt_and_deterministic():
    work = [5, 9, 1, 7, 3, 3]
    parts = scheduler.plan(work, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(work)))
    assert all(p == sorted(p) for p in parts)
    loads = [sum(work[i] for i in p) for p in parts]
    assert max(loads) <= 10 and parts == scheduler.plan(work, 3)
    assert scheduler.plan(work, 1) == [list(range(len(work)))]
    assert scheduler.plan([], 4) == [[], [], [], []]
    # more workers than tasks: one task each, the rest idle
    assert sorted(len(p) for p in scheduler.plan([2, 1], 4)) == [0, 0, 1, 1]


def _matcher(tmp_path, tag):
    from napkon_string_matching import synthetic as syn
    from napkon_string_matching.matcher import Matcher
    from napkon_string_matching.types.gecco_definition import GeccoDefinition
    from napkon_string_matching.types.questionnaire import Questionnaire

    vocab = syn.vocabulary(1500)
    qs = {n: Questionnaire(syn.questionnaire_frame(30 + 7 * s, s, vocab, n))
          for n, s in (("pop", 2), ("hap", 1), ("suep", 3))}
    gecco = GeccoDefinition(syn.definitions_frame(15, 4, vocab))
    config = {"matching": {"score_threshold": 0.1, "cache_threshold": 0.05, "compare_column": "Term",
                           "score_func": "intersection_vs_union", "variable_score_threshold": 0.9,
                           "filter_categories": False, "calculate_tokens": False},
              "steps": ["variables", "gecco", "questionnaires"],
              "output_dir": str(tmp_path / f"out_{tag}"), "cache_dir": str(tmp_path / f"cache_{tag}")}
    return Matcher(None, config, gecco=gecco, questionnaires=qs), config


def _frames(matcher):
    return {name: comp.dataframe().sort_values(list(comp.dataframe().columns)).reset_index(drop=True)
            for name, comp in matcher.results.items()}


class CountingEngine(OracleEngine):
    def __init__(self):
        super().__init__()
        self.calls = 0

    def all_pairs(self, *a, **kw):
        self.calls += 1
        return super().all_pairs(*a, **kw)


def test_threaded_mode_equals_sequential(tmp_path):
    from napkon_string_matching.gpu import engine as engine_mod

    seq, _ = _matcher(tmp_path, "seq")
    with engine_mod.use_engine(OracleEngine()):
        for step in ("variables", "gecco", "questionnaires"):
            seq.match_steps([step])
    par, _ = _matcher(tmp_path, "par")
    engines = [CountingEngine() for _ in range(3)]
    tasks = par._variable_tasks() + par._gecco_tasks() + par._questionnaire_tasks()
    assert len(tasks) == 9
    for task, result in zip(tasks, scheduler.run_comparisons(tasks, engines=engines)):
        par.results[task.name] = result
    assert list(par.results.results) == list(seq.results.results)
    assert all(e.calls >= 2 for e in engines), [e.calls for e in engines]
    a, b = _frames(seq), _frames(par)
    for name in a:
        assert a[name].equals(b[name]), name


def test_worker_errors_reach_the_caller(tmp_path):
    class Boom:
        def __len__(self):
            return 3

        def compare(self, other, **kw):
            raise ZeroDivisionError("division by zero")

    tasks = [scheduler.ComparisonTask(f"t{i}", Boom(), Boom()) for i in range(4)]
    try:
        scheduler.run_comparisons(tasks, engines=[OracleEngine(), OracleEngine()])
    except ZeroDivisionError:
        return
    raise AssertionError("the worker's exception was swallowed")


RANKS_WORKER = textwrap.dedent("""
    import os, sys, pathlib
    sys.path.insert(0, {pkg!r}); sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
    import torch.distributed as dist
    import test_scheduler as ts
    from conftest import OracleEngine
    from napkon_string_matching import matching
    from napkon_string_matching.gpu import distributed, engine as engine_mod

    dist.init_process_group("gloo")
    rank = dist.get_rank()
    tmp = pathlib.Path({tmp!r}) / f"rank{{rank}}"
    engine = ts.CountingEngine()
    with engine_mod.use_engine(engine):
        seq, _ = ts._matcher(tmp, "seq")
        with distributed.whole_comparisons():
            for task in seq._variable_tasks() + seq._gecco_tasks() + seq._questionnaire_tasks():
                seq.results[task.name] = task.run()
        calls_seq = engine.calls
        par, config = ts._matcher(tmp, "par")
        matching.match(config, matcher=par)          # 9 comparisons dealt out over 2 ranks
        calls_par = engine.calls - calls_seq
    assert list(par.results.results) == list(seq.results.results)
    a, b = ts._frames(seq), ts._frames(par)
    for name in a:
        assert a[name].equals(b[name]), name
    assert 0 < calls_par < calls_seq, (calls_par, calls_seq)   # only my share ran here
    dist.destroy_process_group()
    sys.stdout.write("rank%sok" % rank + chr(10))
""")


def test_rank_mode_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(RANKS_WORKER.format(pkg=str(PKG), root=str(ROOT), tests=str(ROOT / "tests"),
                                          tmp=str(tmp_path)))
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29581", str(script)],
        capture_output=True, text=True, timeout=300, env={**os.environ, "OMP_NUM_THREADS": "2"})
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2, res.stdout


def test_affinity_is_best_effort_without_a_gpu():
    """gpu/affinity.py never raises: without NVML / a GPU it reports None and changes nothing."""
    import torch

    from napkon_string_matching.gpu import affinity

    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert affinity.gpu_local_cpus(0) is None and affinity.bind_to_gpu(0) is None
    assert os.sched_getaffinity(0) == before or torch.cuda.is_available()


FAILING_RANK_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {pkg!r}); sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
    import torch.distributed as dist
    from napkon_string_matching.gpu import scheduler

    class Item:
        def __init__(self, n, boom=False):
            self.n, self.boom = n, boom
        def __len__(self):
            return self.n
        def compare(self, other, **kw):
            if self.boom:
                raise ZeroDivisionError("division by zero")
            return self.n * len(other)

    dist.init_process_group("gloo")
    # task 1 (the second largest) lands on rank 1 and fails there; rank 0 must not hang in the exchange
    tasks = [scheduler.ComparisonTask(f"t{{i}}", Item(n, boom=(i == 1)), Item(3)) for i, n in enumerate((9, 8, 2, 1))]
    try:
        scheduler.run_comparisons(tasks)
    except ZeroDivisionError:
        sys.stdout.write("rank%sraised" % dist.get_rank() + chr(10))
    dist.destroy_process_group()
""")


def test_a_failing_rank_fails_every_rank_instead_of_hanging(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(FAILING_RANK_WORKER.format(pkg=str(PKG), root=str(ROOT), tests=str(ROOT / "tests")))
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29583", str(script)],
        capture_output=True, text=True, timeout=120, env={**os.environ, "OMP_NUM_THREADS": "2"})
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("raised") == 2, res.stdout + res.stderr
