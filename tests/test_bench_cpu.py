"""Host logic of bench.py that needs no GPU: the clock sampler (one looping nvidia-smi, fallback to one process per
sample) and the reference arm's JSON line (`--impl reference`: the oracle port on the host cores, the contract of the
task's measurement section)."""
import importlib.util
import json
import os
import stat
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("cavp_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _fake_nvidia_smi(dirpath, loop_works):
    path = os.path.join(dirpath, "nvidia-smi")
    loop = ('while true; do echo "0, 1900, 1965, 800.1, Not Active, Not Active, Not Active, Active"; sleep 0.05; done'
            if loop_works else "exec sleep 30")
    with open(path, "w") as f:
        f.write('#!/bin/bash\nif [[ "$*" == *"-lms"* ]]; then ' + loop +
                '; else echo "0, 1800, 1965, 700, Not Active, Not Active, Not Active, Not Active"; fi\n')
    os.chmod(path, os.stat(path).st_mode | stat.S_IEXEC)


@pytest.mark.parametrize("loop_works", [True, False], ids=["looping", "fallback"])
def test_clock_sampler_keeps_only_samples_of_the_armed_region(tmp_path, monkeypatch, loop_works):
    _fake_nvidia_smi(str(tmp_path), loop_works)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    bench = _load_bench()
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.3)          # lines printed before arm() must not count
    s.arm()                  # (fallback mode: waits 3 s for the looping process, then samples one process at a time)
    n0 = len(s.samples)
    time.sleep(0.6)
    s.stop()
    s.join(timeout=2)
    out = s.summary()
    assert n0 <= 1 and out["samples"] >= 2
    assert out["sm_max_mhz"] == 1965.0
    if loop_works:
        assert out["sm_mhz"] == 1900.0 and out["reasons"] == ["sw_power_cap"]
    else:
        assert out["sm_mhz"] == 1800.0 and out["reasons"] == []


def test_reference_arm_line_has_the_contract_keys():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torch.distributed.run exports; the arm must override it
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "0", "--cpu-batch", "2"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 0 and line["gpu_launches"] == 0
    assert line["value"] > 0 and abs(line["value"] - 2 / (line["ms_per_step"] / 1e3)) < 1e-6 * line["value"] + 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and "2 images per step" in cb["sample"]
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert cb["cores"] == ncpu  # every host core although OMP_NUM_THREADS=1 was exported
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "configs[1]" in line["config"]["workload"]


def test_reference_arm_on_other_ranks_exits_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                        "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""
