"""csrc/worker_pool.h (the chunk workers of whisper_b200_full_batch, kept between calls): a stand-alone stress run — 1 200 generations with
growing and shrinking worker counts; every worker of a generation runs the job exactly once, every item is processed exactly once,
run() returns only when all have finished, destruction joins the parked threads."""
import os
import subprocess

from conftest import ROOT


def test_worker_pool_stress(tmp_path):
    exe = str(tmp_path / "worker_pool_stress")
    src = os.path.join(ROOT, "tests", "hostlogic", "worker_pool_stress.cpp")
    inc = os.path.join(ROOT, "godot-whisper_b200", "csrc")
    res = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + inc, src, "-o", exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr[-2000:]
    assert run.stdout.strip() == "ok 1200"
