"""where the end-to-end time goes: feed, read, and a bare pinned D2H / H2D of the same sizes"""
import sys, os, time, ctypes as C, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import torch, pdmp3_b200
from bench import make_stream
stream = make_stream(1000000)
L = pdmp3_b200.lib()
n = len(stream)
hin = torch.empty(n, dtype=torch.uint8).pin_memory(); hin.numpy()[:] = stream
parsed = pdmp3_b200.parse_stream(stream, lookahead=0)
out_bytes = parsed.n_frames * 4608
hout = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
dev = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); hout.copy_(dev, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    d2 = dev[:n]; t2 = time.perf_counter(); d2.copy_(hin, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
print("bare D2H %.1f ms (%.1f GB/s)  bare H2D %.1f ms (%.1f GB/s)" % (1e3 * (t1 - t0), out_bytes / (t1 - t0) / 1e9, 1e3 * (t3 - t2), n / (t3 - t2) / 1e9))
dec = pdmp3_b200.Decoder("b200:ring=%d,device=0,mode=fast" % (n + 4096))
for it in range(4):
    dec.open_feed(); torch.cuda.synchronize()
    t0 = time.perf_counter(); rc = L.pdmp3_feed(dec.h, hin.data_ptr(), n); t1 = time.perf_counter()
    done = C.c_size_t(0); rc = L.pdmp3_read(dec.h, hout.data_ptr(), out_bytes, C.byref(done)); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("feed %.1f ms  read %.1f ms  total %.1f ms  done %d" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0), done.value))
t0 = time.perf_counter(); p2 = pdmp3_b200.parse_stream(stream, lookahead=0); t1 = time.perf_counter()
print("host parse of 1M frames (default threads): %.1f ms; cores %d" % (1e3 * (t1 - t0), os.cpu_count()))
