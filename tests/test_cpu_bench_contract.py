"""bench.py's contract on a CPU-only box: the reference arm prints ONE JSON line with the keys the driver reads (and times only the
CPU reference: no kernel library is loaded), the GPU arm refuses to run without a device instead of falling back."""
import json, os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, timeout=600)


def test_reference_arm_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoded_pcm_sample_frames_per_sec" and d["unit"] == "sample-frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 1e5 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if cb["kind"] == "reference":                              # oracle/_ref present: the other CPU baselines of SURVEY 8d ride along
        v = cb["variants"]
        assert v["O2_ieee_1core"]["cores"] == 1 and 0 < v["O2_ieee_1core"]["value"] <= d["value"] * 1.5
        if "stock_makefile_flags_1core" in v:
            assert v["stock_makefile_flags_1core"]["value"] > 0 and v["stock_makefile_flags_allcores"]["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    """under torchrun (N > 1) rank 0 alone runs the CPU reference; the other ranks print nothing and exit 0"""
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run(["--steps", "1", "--warmup", "0", "--no-cpu", "--no-e2e"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")], "no bench line may be printed without a device"
