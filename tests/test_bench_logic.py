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


def test_reference_arm_line_on_cpu():
    """`bench.py --impl reference` needs no GPU: it times the CPU oracle and prints one JSON line whose grid is the
    sample that actually ran (the driver computes the ratio from it)."""
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "32",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["value"] > 0
    assert line["config"]["grid"] == [32, 32, 32] and line["config"]["workload_grid"] == [32, 32, 32]
    assert line["cpu_baseline"]["kind"] == "port-tidy" and line["cpu_baseline"]["faithful"]["kind"] == "port-faithful"
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["unit"] == "Gvoxel-updates/s" and line["metric"] == importlib.import_module("bench").METRIC


def test_bench_main_dry_run_on_the_emulated_core(emul_lib, monkeypatch, capsys):
    """The whole of bench.main() -- timed loop, e2e loops, per-kernel list, roofline, CPU baseline leg, JSON line -- driven
    on CPU: torch.cuda is stubbed out and the build step hands back the host-emulated core.  Numbers are meaningless here;
    the point is that the line the driver parses is produced and carries every contract key."""
    import sys

    import torch

    bench = importlib.import_module("bench")
    bld = importlib.import_module("3dfluidsimulation_b200.build")
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_: None)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, pin_memory=False, **k: real_empty(*a, **k))
    monkeypatch.setattr(bld, "build", lambda *a, **k: emul_lib)
    monkeypatch.setattr(bld, "LIB", emul_lib)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "32", "--grid", "16,16,16", "--steps", "2", "--warmup", "1",
                                      "--no-extra", "--no-graph"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["steps"] == 2 and line["n_gpus"] == 1 and line["value"] > 0 and line["gpu_launches"] > 0
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernels"):
        assert key in line["roofline"], key
    assert {k["kernel"] for k in line["roofline"]["kernels"]} >= {"relax_vec4<JACOBI>", "relax_vec4<SMOOTH>", "divergence_vec4"}
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] == 16 * 16 * 16 + 12
    assert e2e["full_fields_value"] > 0 and e2e["full_fields_d2h_bytes_per_step"] == 2 * 16 ** 3 * 4
    assert line["cpu_baseline"]["kind"] == "port-tidy" and line["config"]["warmup_steps_run"] >= 3
