"""Host-side API surface (CPU): task registry / sampler pinned to the reference, descriptors, the C ABI
library exports every symbol of include/sag_b200.h, product refuses to run without CUDA."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import safe_adaptation_gym_b200 as sag
from safe_adaptation_gym_b200 import _abi, _build, benchmark, tasks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_registry_and_sampler_sequences_match_reference():
    gold = json.load(open(os.path.join(G, "sampler.json")))
    assert list(benchmark.TASKS.keys()) == gold["registry"]
    for name in ("multitask", "task_adaptation"):
        b = benchmark.make(name, 30, 666)
        assert [n for n, _ in b.train_tasks] == gold[name]["train"]
        assert [n for n, _ in b.test_tasks] == gold[name]["test"]
    with pytest.raises(AssertionError):
        benchmark.make("domain_randomization")


def test_task_descriptors_match_reference_layouts():
    layouts = json.load(open(os.path.join(G, "layouts.json")))
    seen = set()
    for L in layouts:
        if L["fail"]:
            continue
        t = benchmark.TASKS[L["task"]]()
        counts = [sum(1 for k in L["order"] if k.startswith(p)) for p in ("hazards", "vases", "gremlins", "pillars")]
        assert list(t.obstacles) == counts, L["task"]
        seen.add(L["task"])
    assert len(seen) == 14
    assert tasks.GoToGoal().task_id == benchmark.TASK_IDS["go_to_goal"] == 3
    assert sorted(benchmark.TASK_IDS.values()) == list(range(14))  # ids = alphabetical registry order
    assert [benchmark.TASK_IDS[k] for k in benchmark.TASKS] == list(range(14))


def test_abi_library_exports_every_declared_symbol():
    lib = _build.build()
    hdr = open(os.path.join(ROOT, "include", "sag_b200.h")).read()
    declared = set(re.findall(r"\b(sag_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_abi.SagLib.SYMBOLS), declared ^ set(_abi.SagLib.SYMBOLS)
    L = ctypes.CDLL(lib)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert L.sag_abi_version() == 3
    cfg = _abi.SagConfig()
    L.sag_default_config(ctypes.byref(cfg))
    assert (cfg.robot_keepout, cfg.hazards_size, cfg.vases_keepout, cfg.action_noise, cfg.max_bound) == (0.4, 0.2, 0.15, 0.01, 25.0)


def test_product_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_abi.SagError):
        sag.make("point", "go_to_goal", num_envs=4)


def test_product_never_imports_oracle_or_hostemu():
    pkg = os.path.join(ROOT, "safe_adaptation_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "hostemu" not in src.replace("tests/hostemu", ""), f


def test_error_conventions():
    with pytest.raises(KeyError):
        sag.make("hexapod", "go_to_goal")
    with pytest.raises(KeyError):
        benchmark.TASKS["fly_to_goal"]
