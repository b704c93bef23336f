"""The pdmp3_* streaming API of libpdmp3_b200.so against the UNMODIFIED reference library
(oracle/_ref/libpdmp3_ref.so) driven by the same caller loop -- the CLI loop of pdmp3.c:2564-2584
(pdmp3_read 16 KiB, pdmp3_feed 4096 B on NEED_MORE): same return codes, same `done` counts,
same number of frames (Q7: the last 1-2 frames are never output), same PCM."""
import ctypes as C, os
import numpy as np, pytest
import p3harness as H

pytestmark = pytest.mark.gpu
REFLIB = os.path.join(H.ROOT, "oracle", "_ref", "libpdmp3_ref.so")


class RefApi:
    """ctypes view of the reference's own API (checker)."""
    def __init__(self):
        L = C.CDLL(REFLIB)
        L.pdmp3_new.restype = C.c_void_p; L.pdmp3_new.argtypes = [C.c_char_p, C.c_void_p]
        L.pdmp3_open_feed.argtypes = [C.c_void_p]; L.pdmp3_delete.argtypes = [C.c_void_p]
        L.pdmp3_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.pdmp3_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.pdmp3_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.pdmp3_getformat.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.L = L
        self.h = L.pdmp3_new(None, None)
        C.memset(self.h, 0, 39912)                       # G0: the reference's pdmp3_new is a bare malloc (pdmp3.c:2352)
    def open_feed(self): return self.L.pdmp3_open_feed(self.h)
    def feed(self, b):
        b = np.ascontiguousarray(b, dtype=np.uint8); return self.L.pdmp3_feed(self.h, b.ctypes.data if len(b) else None, len(b))
    def read(self, n):
        out = np.zeros(n, np.uint8); done = C.c_size_t(0)
        rc = self.L.pdmp3_read(self.h, out.ctypes.data, n, C.byref(done)); return rc, out[:done.value]
    def getformat(self):
        r = C.c_long(); c = C.c_int(); e = C.c_int(); rc = self.L.pdmp3_getformat(self.h, C.byref(r), C.byref(c), C.byref(e))
        return rc, r.value, c.value, e.value
    def close(self): self.L.pdmp3_delete(self.h)


def cli_loop(dec, stream, outsize=16384, feedsize=4096, getformat_at=None):
    """The reference CLI loop; returns (PCM bytes, trace of (rc, done))."""
    dec.open_feed()
    fed, pcm, trace, it = 0, [], [], 0
    while True:
        rc, out = dec.read(outsize)
        trace.append((rc, len(out)))
        if rc == -1: break
        pcm.append(out.copy())
        if getformat_at is not None and it == getformat_at: trace.append(("fmt",) + tuple(dec.getformat()))
        if rc == -10:
            chunk = stream[fed:fed + feedsize]
            if len(chunk) == 0: break
            assert dec.feed(chunk) == 0
            fed += len(chunk)
        it += 1
    return np.concatenate(pcm) if pcm else np.zeros(0, np.uint8), trace


@pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,kw,outsize", [
    ("cfg1", dict(H.CONFIGS["cfg1_128k_stereo_long"]), 16384),
    ("cfg3", dict(H.CONFIGS["cfg3_320k_js_ms"]), 16384),
    ("cfg4", dict(H.CONFIGS["cfg4_vbr_mixed"]), 16384),
    ("mono", dict(mode=3, blocks=1, bitrate_index=7), 16384),
    ("odd_outsize", dict(H.CONFIGS["cfg1_128k_stereo_long"]), 5000),
])
def test_cli_loop_matches_reference(name, kw, outsize):
    import pdmp3_b200
    s, _ = H.synth(60, seed=41, **kw)
    ref = RefApi(); a_pcm, a_tr = cli_loop(ref, s, outsize, getformat_at=3); ref.close()
    dec = pdmp3_b200.Decoder("b200:mode=exact"); b_pcm, b_tr = cli_loop(dec, s, outsize, getformat_at=3); dec.close()
    assert a_tr == b_tr, "return codes / done counts differ"
    assert len(a_pcm) == len(b_pcm) and len(a_pcm) > 0
    assert np.array_equal(a_pcm, b_pcm), "PCM differs from the reference API"
    dec = pdmp3_b200.Decoder(); c_pcm, c_tr = cli_loop(dec, s, outsize, getformat_at=3); dec.close()     # default: FAST mode
    assert a_tr == c_tr
    assert np.abs(a_pcm.view(np.int16).astype(np.int32) - c_pcm.view(np.int16).astype(np.int32)).max() <= 1


def test_api_codes_and_options():
    import pdmp3_b200
    P = pdmp3_b200
    s, _ = H.synth(40, seed=2)
    d = P.Decoder()
    assert d.open_feed() == P.PDMP3_OK
    assert d.feed(np.zeros(0, np.uint8)) == P.PDMP3_ERR                      # zero size (pdmp3.c:2392,2422)
    assert d.feed(s[:16384]) == P.PDMP3_OK
    assert d.feed(s[:1]) == P.PDMP3_NO_SPACE                                 # ring of INBUF_SIZE is full (pdmp3.c:2394,2420)
    rc, out = d.read(4608 * 2)
    assert rc == P.PDMP3_NEW_FORMAT and len(out) == 9216                     # NEW_FORMAT replaces OK until getformat (2470)
    assert d.getformat() == (0, 44100, 2, P.PDMP3_ENC_SIGNED_16)
    rc, out = d.read(4608)
    assert rc == P.PDMP3_OK and len(out) == 4608
    d.close()
    # a large ring takes a whole stream at once: one feed + one read
    big, _ = H.synth(500, seed=3, **H.CONFIGS["cfg3_320k_js_ms"])
    d = P.Decoder("b200:mode=exact,ring=%d" % (len(big) + 10))
    d.open_feed()
    assert d.feed(big) == P.PDMP3_OK
    rc, out = d.read(500 * 4608)
    assert rc == P.PDMP3_NEED_MORE and len(out) == 499 * 4608                # the 1152-byte rule leaves the last frame (Q7)
    o = H.oracle_decode(big, lookahead=1152)
    assert np.array_equal(out.view(np.int16).reshape(-1, 1152, 2), o["pcm"])
    # decode(): header peek without output buffer -> NEW_FORMAT (pdmp3.c:2507-2516)
    d2 = P.Decoder(); d2.open_feed()
    rc, out = d2.decode(big[:3000], 0)
    assert rc == P.PDMP3_NEW_FORMAT
    assert d2.getformat()[1:3] == (44100, 2)
    d2.close(); d.close()


def test_cli_entry_writes_raw_file(tmp_path):
    """pdmp3(char*const*) (pdmp3.c:2540-2589): decodes files to <file>.raw with the reference's read/feed loop."""
    import pdmp3_b200
    s, _ = H.synth(50, seed=9, **H.CONFIGS["cfg1_128k_stereo_long"])
    f = tmp_path / "a.mp3"; f.write_bytes(s.tobytes())
    L = pdmp3_b200.lib()
    argv = (C.c_char_p * 2)(str(f).encode(), None)
    L.pdmp3.argtypes = [C.POINTER(C.c_char_p)]; L.pdmp3.restype = None
    L.pdmp3(argv)
    raw = np.fromfile(str(f) + ".raw", dtype=np.int16)
    o = H.oracle_decode(s, lookahead=1152, taps=False)["pcm"]
    assert raw.size == o.size
    assert np.abs(raw.astype(np.int32) - o.reshape(-1).astype(np.int32)).max() <= 1      # default FAST mode


@pytest.mark.parametrize("opts", ["b200:ring=4000000", "b200:ring=4000000,hop=host"])
def test_feed_borrow_equals_copying_feed(opts):
    """feed=borrow (no copy in pdmp3_feed: the handle decodes out of the caller's buffer) delivers the same bytes and return codes
    as the copying feed: one big feed + one big read, a second feed while a rest of the borrowed buffer is pending (the rest moves
    into the handle's own buffer), and small reads in the reference's CLI pattern."""
    import pdmp3_b200
    s, _ = H.synth(700, seed=71, **H.CONFIGS["cfg4_vbr_mixed"])
    half = len(s) // 2
    got = {}
    for mode in ("", ",feed=borrow"):
        d = pdmp3_b200.Decoder(opts + mode); d.open_feed()
        a, b = np.ascontiguousarray(s[:half]), np.ascontiguousarray(s[half:])     # kept alive below: borrowed buffers belong to the caller
        trace = [d.feed(a)]
        rc, p1 = d.read(700 * 4608); trace.append((rc, len(p1)))
        trace.append(d.feed(b))
        rc, p2 = d.read(700 * 4608); trace.append((rc, len(p2)))
        rc, p3 = d.read(700 * 4608); trace.append((rc, len(p3)))
        got[mode] = (np.concatenate([p1, p2, p3]), trace)
        d.close()
    assert got[""][1] == got[",feed=borrow"][1], (got[""][1], got[",feed=borrow"][1])
    assert np.array_equal(got[""][0], got[",feed=borrow"][0]) and len(got[""][0]) >= 690 * 4608
    # the CLI pattern on a borrowed whole-stream feed: small reads out of the caller's buffer
    d = pdmp3_b200.Decoder("b200:ring=65536,feed=borrow"); d.open_feed()
    whole = np.ascontiguousarray(s); assert d.feed(whole) == 0
    parts = []
    while True:
        rc, out = d.read(16384)
        parts.append(out.copy())
        if rc in (-1, -10): break
    d.close()
    assert np.array_equal(np.concatenate(parts), got[""][0])
