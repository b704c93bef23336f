"""BASELINE configs[4] as a product path: N ranks under torch.distributed.run, NCCL scatter / chunked gather called
from the C side (p3_sharded_decode); rank 0's gathered PCM must be bit-identical to the single-GPU decode of the same stream,
in both modes.  Skipped on a box with fewer than 2 GPUs."""
import os, subprocess, sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(n):
    env = dict(os.environ); env.pop("RANK", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tests", "dist_worker.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-4000:]
    for k in range(n):
        assert "DIST_OK rank %d of %d" % (k, n) in r.stdout, r.stdout[-4000:]


def test_sharded_decode_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2)


def test_sharded_decode_all_ranks():
    import torch
    n = torch.cuda.device_count()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    _run(n)


@pytest.mark.parametrize("mode", ["fast", "exact"])
def test_sharded_decode_world_of_one(mode):
    """the rank-0 path alone (device hop of the whole stream, plan, chunked decode into the output buffer) with a world of 1"""
    import numpy as np, pdmp3_b200
    import p3harness as H
    ctx = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST if mode == "fast" else pdmp3_b200.MODE_EXACT)
    d = pdmp3_b200.Dist(ctx, pdmp3_b200.dist_unique_id(), 0, 1)
    one = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST if mode == "fast" else pdmp3_b200.MODE_EXACT)
    for name, n, chunk in (("cfg4_vbr_mixed", 700, 64), ("cfg3_320k_js_ms", 2000, 512), ("cfg3_320k_js_ms", 2, 0)):
        s, _ = H.synth(n, seed=92, **H.CONFIGS[name])
        res = d.sharded_decode(stream=s, chunk_frames=chunk)
        assert res["n_frames_total"] == n and res["n_frames_mine"] == n and res["warmup_mine"] == 0
        one.reset()
        assert np.array_equal(d.pcm(res), one.decode(s, lookahead=0))
    d.close(); ctx.close(); one.close()
