"""ctypes binding of the synthetic stream generator (tools/libp3synth.so, built from tools/p3_synth.c).
Bench and test infrastructure; it knows nothing about the oracle or the reference, so that bench.py's
GPU arm can produce its workload without importing the test harness."""
import ctypes as C, os, numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class SynthCfg(C.Structure):
    _fields_ = [("seed", C.c_uint64)] + [(n, C.c_int32) for n in (
        "bitrate_index", "mode", "mode_ext", "sfreq", "blocks", "reservoir", "scalefacs", "gain",
        "fill_pm", "crc", "count1_b_pm", "overrun_pm", "max_table", "garbage_pm", "iso", "peak_pm")]


_DEF = dict(seed=1, bitrate_index=9, mode=0, mode_ext=0, sfreq=0, blocks=0, reservoir=1, scalefacs=1,
            gain=172, fill_pm=850, crc=0, count1_b_pm=0, overrun_pm=0, max_table=31, garbage_pm=0, iso=0, peak_pm=0)

# the BASELINE.json configurations
CONFIGS = {
    "cfg1_128k_stereo_long": dict(bitrate_index=9, mode=0, blocks=0),
    "cfg3_320k_js_ms":       dict(bitrate_index=14, mode=1, mode_ext=2, blocks=0),
    "cfg4_vbr_mixed":        dict(bitrate_index=0, mode=1, mode_ext=-1, blocks=1, overrun_pm=30),
}

_synth = None


def synth(n_frames, want_is=False, **kw):
    """-> (stream bytes as np.uint8, optional encoded spectra [n,2,2,576] int16)"""
    global _synth
    if _synth is None:
        p = os.path.join(ROOT, "tools", "libp3synth.so")
        if not os.path.exists(p):
            raise RuntimeError("tools/libp3synth.so missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _synth = C.CDLL(p)
        _synth.p3_synth.restype = C.c_int64
        _synth.p3_synth.argtypes = [C.POINTER(SynthCfg), C.c_int64, C.c_void_p, C.c_uint64, C.c_void_p]
    d = dict(_DEF); d.update(kw)
    cfg = SynthCfg(**d)
    cap = int(n_frames) * 1500 + 4096
    buf = np.zeros(cap, dtype=np.uint8)
    iso = np.zeros((n_frames, 2, 2, 576), dtype=np.int16) if want_is else None
    n = _synth.p3_synth(C.byref(cfg), n_frames, buf.ctypes.data, cap, iso.ctypes.data if want_is else None)
    if n < 0:
        raise RuntimeError("p3_synth failed: %d" % n)
    return buf[:n].copy(), iso
