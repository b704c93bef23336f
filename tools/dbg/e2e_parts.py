"""where the end-to-end time goes: a bare pinned D2H of the PCM in one piece and in batch-sized pieces, then pdmp3_feed + pdmp3_read
(feed=borrow) with P3_TRACE for several batch sizes"""
import sys, os, time, ctypes as C, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch, pdmp3_b200
from bench import make_stream
stream = make_stream(1000000)
L = pdmp3_b200.lib()
n = len(stream)
hin = torch.empty(n, dtype=torch.uint8).pin_memory(); hin.numpy()[:] = stream
out_bytes = 1000000 * 4608
hout = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
dev = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
for pieces in (1, 31, 8):
    step = (out_bytes // pieces + 4607) // 4608 * 4608
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for o in range(0, out_bytes, step): hout[o:o + step].copy_(dev[o:o + step], non_blocking=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print("bare D2H in %2d piece(s): %.1f ms (%.1f GB/s)" % (pieces, 1e3 * (t1 - t0), out_bytes / (t1 - t0) / 1e9), flush=True)
for batch in (32768, 65536, 131072, 262144):
    dec = pdmp3_b200.Decoder("b200:ring=65536,device=0,mode=fast,feed=borrow,batch=%d" % batch)
    ts = []
    for it in range(5):
        dec.open_feed(); torch.cuda.synchronize()
        t0 = time.perf_counter(); rc = L.pdmp3_feed(dec.h, hin.data_ptr(), n)
        done = C.c_size_t(0); rc = L.pdmp3_read(dec.h, hout.data_ptr(), out_bytes, C.byref(done)); torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append(1e3 * (t2 - t0))
    print("batch %6d: e2e %s ms, done %d" % (batch, " ".join("%.1f" % t for t in ts), done.value), flush=True)
    dec.close()
