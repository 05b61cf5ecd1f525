"""CPU tier: the host logic of bench.py that does not need a GPU -- the warm-up / settle loop and the timed loop of
timed_run() (driven with the host-emulated core and a one-rank stand-in for the torch.distributed job), the sample
the CPU baseline is taken on, and the workload table against BASELINE.json."""
import importlib
import json
import os

import numpy as np

from conftest import ROOT


class OneRankJob:
    rank, world, local = 0, 1, 0

    def barrier(self):
        pass

    def max_over_ranks(self, v):
        return v

    def connect(self, s):
        pass


def test_timed_run_counts_steps(emul_lib, pkg):
    bench = importlib.import_module("bench")
    job = OneRankJob()
    s, one_step, ncells = bench.make_plume_solver(pkg, job, emul_lib, (16, 16, 16), 2, 3, 0, True, False)
    calls = {"n": 0}

    def counted():
        calls["n"] += 1
        one_step()

    try:
        ms, launches, clocks = bench.timed_run(job, s, counted, steps=2, warmup=1, settle_s=0.0)
        assert calls["n"] == 3 and ms > 0 and launches > 0 and clocks is None
        calls["n"] = 0
        bench.timed_run(job, s, counted, steps=1, warmup=2, settle_s=1e9)      # settle loop is capped
        assert calls["n"] == 2 + 200 + 1
        assert ncells > 0 and float(s.get_field("density").sum()) > 0.0
    finally:
        s.close()


def test_cpu_sample_is_a_slab_of_the_workload():
    bench = importlib.import_module("bench")
    assert bench.cpu_sample_dims(512) == (512, 512, 64)
    for n in (32, 128):
        assert bench.cpu_sample_dims(n) == (n, n, n)


def test_workloads_follow_baseline_json():
    bench = importlib.import_module("bench")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.WORKLOADS["512"][:3] == (512, 20, 80)
    assert "Gvoxel" in bench.METRIC and "voxel" in json.dumps(base).lower()
    assert bench.step_bytes_per_voxel(20, 80) == 88 * 20 + 26 * 80 + 182
