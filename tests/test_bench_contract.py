"""bench.py's driver contract, checked where no GPU is needed: the reference arm (CPU port of the reference's algorithm)
prints exactly one JSON line with the keys the driver reads, and the GPU arm refuses to run without a device instead of
falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], {"EXB_REF_READS_PER_THREAD": "20000"})
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    lines = [l for l in out.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("read_fastq + mean-quality filter")
    assert d["config"]["workload"].startswith("C2:")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the port counts what the oracle counts: every generated record, about 59 % of them above mean quality 30
    assert d["records"] == d["cpu_baseline"]["cores"] * 20000 and 0.4 < d["pass"] / d["records"] < 0.8


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.decode().strip() == ""


def test_gpu_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = _run(["--steps", "1", "--warmup", "0"])
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr.decode() + out.stdout.decode())
