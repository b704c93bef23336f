"""Worker of tests/test_gpu_dist.py, one process per GPU (python -m torch.distributed.run --nproc-per-node N tests/dist_worker.py).
Rank 0 synthesises a stream, all ranks run p3_sharded_decode() (NCCL scatter of byte ranges / chunked gather of PCM, both called
from the C side); rank 0 compares the gathered PCM with the single-GPU decode of the same stream -- bit-identical, both modes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import p3harness as H, pdmp3_b200


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                          # control plane only: the two NCCL ids travel over it
    cases = [("cfg4", dict(H.CONFIGS["cfg4_vbr_mixed"]), 700, 64), ("cfg3", dict(H.CONFIGS["cfg3_320k_js_ms"]), 3000, 256),
             ("garbage", dict(garbage_pm=200, blocks=1), 500, 96), ("mono", dict(mode=3, blocks=1, bitrate_index=7), 400, 64),
             ("tiny", dict(H.CONFIGS["cfg3_320k_js_ms"]), 3, 64)]
    for mode in (pdmp3_b200.MODE_FAST, pdmp3_b200.MODE_EXACT):
        ctx = pdmp3_b200.Context(local, mode)
        ids = [pdmp3_b200.dist_unique_id() if rank == 0 else None]       # (a pair of NCCL ids serves one pair of communicators)
        dist.broadcast_object_list(ids, src=0)
        d = pdmp3_b200.Dist(ctx, ids[0], rank, world)
        for name, kw, n, chunk in cases:
            for on_device in (False, True):
                if rank == 0:
                    s, _ = H.synth(n, seed=91, **kw)
                    if on_device:
                        t = torch.zeros(len(s) + 64, dtype=torch.uint8, device="cuda"); t[:len(s)] = torch.from_numpy(s).cuda()
                        torch.cuda.synchronize()
                        res = d.sharded_decode(device_ptr=t.data_ptr(), nbytes=len(s), chunk_frames=chunk)
                    else:
                        res = d.sharded_decode(stream=s, chunk_frames=chunk)
                    got = d.pcm(res)
                    one = pdmp3_b200.Context(local, mode)
                    want = one.decode(s, lookahead=0)
                    one.close()
                    assert res["n_frames_total"] == n and got.shape == want.shape, (res, want.shape)
                    assert np.array_equal(got, want), "%s mode %d device %d: gathered PCM differs from the single-GPU decode" % (name, mode, on_device)
                    if mode == pdmp3_b200.MODE_EXACT and n <= 1000:
                        o = H.oracle_decode(s, lookahead=0, taps=False)
                        assert np.array_equal(got, o["pcm"]), name
                else:
                    res = d.sharded_decode(chunk_frames=chunk)
                assert res["n_frames_total"] == n
        # the floor measurement entry point at a small size
        ms = d.measure_ingest(1 << 20, 2)
        assert ms > 0
        d.close(); ctx.close()
    dist.barrier()
    print("DIST_OK rank %d of %d" % (rank, world), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
