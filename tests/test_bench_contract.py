"""The reference arm of bench.py (`--impl reference`: the oracle port of multi_stage_search on the host cores) runs
without a GPU, so its JSON contract is checked here on a small corpus: one line on stdout, the keys the driver reads,
the same `config` object our own arm prints for the same flags, `cpu_baseline` describing the run, zero-byte `e2e`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run("--rows", "20000", "--batch", "64", "--cpu-sample", "16", "--steps", "2", "--warmup", "1")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["higher_is_better"] is True and d["unit"] == "queries/s" and d["value"] > 0
    assert d["metric"].startswith("QPS at recall@10")
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"]
    assert "sample" in cb and "sample" not in d["config"]          # the sample size rides under cpu_baseline
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["rows"] == 20000 and d["config"]["batch"] == 64 and d["config"]["rescore_count"] == 40
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - 16) < 1e-6 * 16 + 1e-9   # value = sampled queries / step time
    assert d["cpu_baseline_select_variant_qps"] > 0                 # the fair CPU number beside the faithful full sort


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run("--gpus", "2", "--rows", "20000", "--batch", "64", "--cpu-sample", "16", "--steps", "1", "--warmup", "0",
                env=env) == []


def test_own_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device our arm must stop with an error and print no result line."""
    import pytest
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--rows", "2000", "--batch", "64", "--steps", "1",
                          "--warmup", "0", "--north-star", "0", "--stream-rows", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
