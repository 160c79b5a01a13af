"""CPU suite, part 7: the reference arm of bench.py (the driver runs it beside the GPU arm) prints one JSON line with
the contract's keys; under torchrun only rank 0 works and prints."""
import json
import os
import subprocess
import sys

import util


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--cpu-queries", "8"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "all-vs-all NN-graph GCUPS" and d["unit"] == "GCUPS"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c2")


def test_reference_arm_other_ranks_do_nothing():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.decode().strip() == ""
