"""bench.py's reference arm on the CPU (the contract of the measurement harness): `--impl reference` prints ONE JSON line with the
metric / unit / config of the GPU arm, `impl`, `cpu_baseline` and a zero-copy `e2e`; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line_config3():
    lines = _run(["--impl", "reference", "--ref-cells", "6,8", "--gpus", "1", "--steps", "2", "--warmup", "3"], {"OMP_NUM_THREADS": "1"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "s/step" and d["higher_is_better"] is False and d["metric"].startswith("time_step_wall_s")
    assert d["config"]["same_config"] is False and d["config"]["extrapolated"] is True and len(d["config"]["sample_seconds"]) == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0
    # torchrun exports OMP_NUM_THREADS=1: the team is sized explicitly all the same
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_stay_silent():
    assert _run(["--impl", "reference", "--ref-cells", "6", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_arm_line_config4():
    d = json.loads(_run(["--impl", "reference", "--config", "4"])[0])
    assert d["impl"] == "reference" and d["config"]["same_config"] is True and "fsi_leaflet_mpi" in d["metric"]
    assert d["cpu_baseline"]["cores"] == 1 and d["value"] > 0
