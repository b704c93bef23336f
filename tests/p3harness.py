"""Test/bench helpers: ctypes wrappers for the synthetic stream generator (tools/libp3synth.so),
the compiled reference with stage taps (oracle/_ref/libref_taps.so) and the oracle restatement
(oracle/libp3_oracle.so).  TEST INFRASTRUCTURE ONLY -- nothing under pdmp3_b200/ imports this."""
import ctypes as C, os, sys, numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

sys.path.insert(0, os.path.join(ROOT, "tools"))
from p3synth import synth, CONFIGS, SynthCfg          # the stream generator lives outside the test harness (bench.py uses it too)

def _lib(path):
    p = os.path.join(ROOT, path)
    if not os.path.exists(p):
        return None
    return C.CDLL(p)

class RefTaps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("side", "hdr", "scf_l", "scf_s", "is_huff", "count1",
                                           "xr_req", "xr_reo", "xr_ste", "xr_ali", "y_hyb", "pcm")]

_ref = None
def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle/_ref/libref_taps.so"))

def ref_lib():
    global _ref
    if _ref is None:
        _ref = _lib("oracle/_ref/libref_taps.so")
        _ref.ref_taps_decode.restype = C.c_long
        _ref.ref_taps_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_long, C.POINTER(RefTaps)]
    return _ref

def ref_decode(stream, max_frames=None, taps=True):
    """Run the UNMODIFIED reference on `stream`; returns dict of numpy tap arrays (trimmed to frames decoded)."""
    lib = ref_lib()
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    mf = int(max_frames if max_frames is not None else len(stream) // 96 + 4)
    shapes = dict(side=((mf, 2, 2, 20), np.int32), hdr=((mf, 8), np.int32), scf_l=((mf, 2, 2, 21), np.uint8),
                  scf_s=((mf, 2, 2, 12, 3), np.uint8), is_huff=((mf, 2, 2, 576), np.int16), count1=((mf, 2, 2), np.int32),
                  xr_req=((mf, 2, 2, 576), np.float32), xr_reo=((mf, 2, 2, 576), np.float32),
                  xr_ste=((mf, 2, 2, 576), np.float32), xr_ali=((mf, 2, 2, 576), np.float32),
                  y_hyb=((mf, 2, 2, 576), np.float32), pcm=((mf, 1152, 2), np.int16))
    if not taps:
        shapes = {k: v for k, v in shapes.items() if k in ("pcm", "hdr")}
    arrs = {k: np.zeros(s, dtype=d) for k, (s, d) in shapes.items()}
    t = RefTaps(**{k: a.ctypes.data for k, a in arrs.items()})
    n = lib.ref_taps_decode(stream.ctypes.data, len(stream), mf, C.byref(t))
    out = {k: a[:n] for k, a in arrs.items()}
    out["n_frames"] = n
    return out

# ---------------------------------------------------------------------------------------------
# host parser (product code, compiled into the oracle .so as well so that CPU-only tests can use it)
class P3Frame(C.Structure):
    _fields_ = [("main_off", C.c_uint64), ("main_pos", C.c_uint64), ("main_size", C.c_uint16), ("main_begin", C.c_uint16),
                ("nch", C.c_uint8), ("mode", C.c_uint8), ("mode_ext", C.c_uint8), ("sfreq", C.c_uint8),
                ("scfsi", C.c_uint8), ("flags", C.c_uint8), ("bitrate_kbps", C.c_uint16), ("pcm_index", C.c_uint32)]
class P3Gc(C.Structure):
    _fields_ = [("w0", C.c_uint32), ("w1", C.c_uint32), ("w2", C.c_uint32), ("w3", C.c_uint32)]
class P3ParseState(C.Structure):
    _fields_ = [("main_pos", C.c_uint64), ("top", C.c_uint32), ("pcm_index", C.c_uint32), ("nch", C.c_int32), ("sfreq", C.c_int32)]
class P3ParseOpts(C.Structure):
    _fields_ = [("max_frames", C.c_int64), ("lookahead", C.c_uint32), ("nthreads", C.c_int32), ("warmup_frames", C.c_uint32), ("hop_only", C.c_uint32), ("iso", C.c_uint32)]
class P3Parsed(C.Structure):
    _fields_ = [("n_frames", C.c_int64), ("frames", C.POINTER(P3Frame)), ("gcs", C.POINTER(P3Gc)),
                ("consumed", C.c_uint64), ("n_pcm_frames", C.c_int64), ("external", C.c_int32), ("stop", C.c_int32),
                ("hop_only", C.c_int32), ("pad_", C.c_int32)]
FRAME_DT = np.dtype([("main_off", "<u8"), ("main_pos", "<u8"), ("main_size", "<u2"), ("main_begin", "<u2"),
                     ("nch", "u1"), ("mode", "u1"), ("mode_ext", "u1"), ("sfreq", "u1"), ("scfsi", "u1"),
                     ("flags", "u1"), ("bitrate_kbps", "<u2"), ("pcm_index", "<u4")])
assert FRAME_DT.itemsize == 32 and C.sizeof(P3Frame) == 32 and C.sizeof(P3Gc) == 16

class OTaps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("is_huff", "count1", "scf_l", "scf_s", "xr_req", "xr_reo", "xr_ste", "xr_ali", "y_hyb")]

_orc = None
def oracle_lib():
    global _orc
    if _orc is None:
        _orc = _lib("oracle/libp3_oracle.so")
        if _orc is None:
            raise RuntimeError("oracle/libp3_oracle.so missing: run __graft_entry__.build()")
        _orc.p3_parse.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(P3ParseOpts), C.POINTER(P3ParseState), C.POINTER(P3Parsed)]
        _orc.p3_parsed_free.argtypes = [C.POINTER(P3Parsed)]
        _orc.p3o_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(OTaps), C.c_void_p]
    return _orc

def parse(stream, lookahead=1152, max_frames=0, warmup=0, nthreads=1, lib=None, iso=False):
    """Run the host parser; returns (frames structured array, gcs uint32 [n,4,4], P3Parsed copy info)."""
    lib = lib or oracle_lib()
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    o = P3ParseOpts(max_frames, lookahead, nthreads, warmup, 0, 1 if iso else 0); st = P3ParseState(0, 0, 0, -1, -1); out = P3Parsed()
    rc = lib.p3_parse(stream.ctypes.data, len(stream), C.byref(o), C.byref(st), C.byref(out))
    assert rc == 0, rc
    n = out.n_frames
    fr = np.frombuffer(C.string_at(out.frames, 32 * n), dtype=FRAME_DT).copy() if n else np.zeros(0, FRAME_DT)
    gc = np.frombuffer(C.string_at(out.gcs, 64 * n), dtype=np.uint32).reshape(n, 4, 4).copy() if n else np.zeros((0, 4, 4), np.uint32)
    info = dict(n_frames=n, consumed=out.consumed, n_pcm_frames=out.n_pcm_frames, stop=out.stop)
    lib.p3_parsed_free(C.byref(out))
    return fr, gc, info

def oracle_decode(stream, lookahead=1152, taps=True, warmup=0, iso=False):
    """Parse with the product parser, decode with the oracle restatement."""
    lib = oracle_lib()
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    fr, gc, info = parse(stream, lookahead, warmup=warmup, iso=iso)
    n = info["n_frames"]; nch = int(fr["nch"][0]) if n else 2
    shapes = dict(is_huff=((n, 2, 2, 576), np.int16), count1=((n, 2, 2), np.int32), scf_l=((n, 2, 2, 21), np.uint8),
                  scf_s=((n, 2, 2, 12, 3), np.uint8), xr_req=((n, 2, 2, 576), np.float32), xr_reo=((n, 2, 2, 576), np.float32),
                  xr_ste=((n, 2, 2, 576), np.float32), xr_ali=((n, 2, 2, 576), np.float32), y_hyb=((n, 2, 2, 576), np.float32))
    arrs = {k: np.zeros(s, dtype=d) for k, (s, d) in shapes.items()} if taps else {}
    pcm = np.zeros((info["n_pcm_frames"], 1152, nch), dtype=np.int16)
    t = OTaps(**{k: a.ctypes.data for k, a in arrs.items()}) if taps else None
    lib.p3o_decode(stream.ctypes.data, fr.ctypes.data, gc.ctypes.data, n, C.byref(t) if taps else None, pcm.ctypes.data)
    arrs["pcm"] = pcm; arrs["frames"] = fr; arrs["gcs"] = gc; arrs["n_frames"] = n
    return arrs

def gc_fields(gc):
    """unpack [n,4,4] uint32 descriptors into the 20-int layout of the reference tap `side`"""
    w0, w1, w2 = gc[..., 0].astype(np.int64), gc[..., 1].astype(np.int64), gc[..., 2].astype(np.int64)
    f = np.zeros(gc.shape[:-1] + (20,), np.int32)
    f[..., 0] = w0 & 0xfff; f[..., 1] = (w0 >> 12) & 0x1ff; f[..., 2] = (w0 >> 21) & 0xff; f[..., 15] = (w0 >> 29) & 1
    f[..., 16] = (w0 >> 30) & 1; f[..., 17] = (w0 >> 31) & 1
    f[..., 3] = w1 & 15; f[..., 4] = (w1 >> 4) & 1; f[..., 5] = (w1 >> 5) & 3; f[..., 6] = (w1 >> 7) & 1
    f[..., 7] = (w1 >> 8) & 31; f[..., 8] = (w1 >> 13) & 31; f[..., 9] = (w1 >> 18) & 31
    f[..., 13] = (w1 >> 23) & 15; f[..., 14] = (w1 >> 27) & 15
    f[..., 10] = w2 & 7; f[..., 11] = (w2 >> 3) & 7; f[..., 12] = (w2 >> 6) & 7
    return f


def empty_some_parts(s, every=7):
    """part2_3_length := 0 for a few granule-channels of a stereo stream, everything else untouched -- in particular
    scalefac_compress, so the reference goes on reading that part's scalefactor bits and the following parts of the frame
    start behind them (pdmp3.c:1379-1435 + 2057-2061), and count1 of the slot stays stale (Q6)."""
    t = s.copy()
    fr, _, _ = parse(s, lookahead=0)
    for i, f in enumerate(fr):
        if i % every != 3 or f["nch"] != 2: continue
        si = int(f["main_off"]) - 32
        k = (i // every) % 4
        pos = 20 + 59 * k                                   # the 12 bits of part2_3_length of granule-channel k
        for b in range(pos, pos + 12):
            t[si + (b >> 3)] &= ~(0x80 >> (b & 7)) & 0xff
    return t


def check_against_golden(g, pcm, is_huff, count1, xr_ali, y_hyb):
    """One decoder's stage outputs against a committed fixture of the UNMODIFIED reference (tests/golden/*.npz, tools/make_golden.py):
    Huffman output, count1 and PCM equal, the float stages equal bit for bit (+0 == -0).  Used with the oracle's arrays on the CPU
    (tests/test_cpu_oracle.py) and with the CUDA path's taps on the GPU (tests/test_gpu_parity.py)."""
    n = int(g["n_frames"]); nch = g["pcm"].shape[2]
    assert pcm.shape == (n, 1152, nch), (pcm.shape, n, nch)
    assert np.array_equal(is_huff[:n, :, :nch], g["is_huff"][:, :, :nch]), "Huffman output"
    assert np.array_equal(count1[:n, :, :nch], g["count1"][:, :, :nch]), "count1"
    for name, bits, got in (("requantize..antialias", g["xr_ali_bits"], xr_ali), ("hybrid synthesis", g["y_hyb_bits"], y_hyb)):
        a = bits[:, :, :nch]; b = np.ascontiguousarray(got[:n, :, :nch], dtype=np.float32).view(np.uint32)
        assert ((a == b) | (((a & 0x7fffffff) == 0) & ((b & 0x7fffffff) == 0))).all(), name
    assert np.array_equal(pcm, g["pcm"]), "PCM"

