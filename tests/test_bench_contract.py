"""CPU: the parts of bench.py's contract that need no GPU - the reference arm's JSON line
(both kinds: the unmodified Python reference from oracle/_ref and the C oracle port), the
ranks that stay silent under torchrun, the loud failure of the CUDA arm without a device,
and the algorithmic-bytes formula of SURVEY.md section 8(d)."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None, timeout=300):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=timeout, env=env, cwd=ROOT)


def _check_reference_line(out, kind, steps, warmup):
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout  # ONE JSON line on stdout, everything else on stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "recom_steps_per_sec" and d["unit"] == "steps/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["steps"] == steps and d["warmup"] == warmup and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("Grid 20x20")
    cb = d["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    return d


def test_reference_arm_with_the_oracle_port():
    out = _run(["--impl", "reference", "--ref-kind", "port", "--workload", "config1", "--steps", "2", "--warmup", "1"])
    assert out.returncode == 0, out.stderr
    _check_reference_line(out, "port", 2, 1)


def test_reference_arm_with_the_unmodified_reference():
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "gerrychain")):
        pytest.skip("oracle/_ref not installed (sh oracle/make_ref.sh in the build container)")
    out = _run(["--impl", "reference", "--workload", "config1", "--steps", "1", "--warmup", "1"])
    assert out.returncode == 0, out.stderr
    d = _check_reference_line(out, "reference", 1, 1)
    assert "unmodified mggg/GerryChain" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)


def test_reference_arm_other_ranks_do_no_work():
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
               {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = _run(["--steps", "1", "--warmup", "0", "--extras", "none", "--no-cpu-baseline"], timeout=300)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CPU fallback" in out.stderr


def test_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench

    # B_step = n_sub (2a + 12) + d_sub (4 + a) + 32, a = 1
    assert bench.algorithmic_bytes(2000, 8000, 1) == 2000 * 14 + 8000 * 5 + 32
    assert bench.algorithmic_bytes(1000, 6000, 1, a=2) == 1000 * 16 + 6000 * 6 + 32
    # the product path never imports the oracle
    src = open(BENCH).read()
    assert "from oracle import" in src  # only inside the CPU legs (port_sample / run_reference)
    for path in os.listdir(os.path.join(ROOT, "gerrychain_b200")):
        if path.endswith(".py"):
            text = open(os.path.join(ROOT, "gerrychain_b200", path)).read()
            assert "import oracle" not in text and "from oracle" not in text, path
