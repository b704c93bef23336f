"""Read_Audio_L3 on the device (k_sideinfo + k_q6_chain): the descriptors the kernels see after the
device parser must equal, bit for bit, what the host parser (p3_parse.c) produces -- every field of the four
granule-channels, scfsi, the validity flag of malformed side info (Q10) and the stale-count1 chain (Q6) -- and a
decode through either path gives the same PCM."""
import numpy as np, pytest
import p3harness as H
from test_gpu_parity import VARIANTS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import pdmp3_b200
    c = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST)
    yield c
    c.close()


def empty_some_parts(s, every=7):
    """part2_3_length := 0 for a few granule-channels (reference quirk Q6: count1 of the slot stays stale)"""
    import pdmp3_b200
    t = s.copy()
    fr = pdmp3_b200.parse_stream(s, lookahead=0).frames()
    for i, f in enumerate(fr):
        if i % every != 3 or f["nch"] != 2: continue
        si = int(f["main_off"]) - 32
        k = (i // every) % 4
        pos = 20 + 59 * k                                   # the 12 bits of part2_3_length of granule-channel k
        for b in range(pos, pos + 12):
            t[si + (b >> 3)] &= ~(0x80 >> (b & 7)) & 0xff
    return t


def compare_descriptors(ctx, s):
    import pdmp3_b200
    host = pdmp3_b200.parse_stream(s, lookahead=0)
    n = host.n_frames
    dev = pdmp3_b200.parse_stream(s, lookahead=0, hop_only=True)
    assert dev.n_frames == n and dev.c.hop_only == 1
    assert not dev.gcs().any(), "hop-only parse must not fill the granule descriptors on the host"
    ctx.reset(); ctx.upload(dev); ctx.run(); ctx.sync()
    fr_d, gc_d = ctx.download_desc(n)
    fr_h, gc_h = host.frames(), host.gcs()
    for name in fr_h.dtype.names:
        assert np.array_equal(fr_h[name], fr_d[name]), "frame field %s" % name
    nch = int(fr_h["nch"][0]) if n else 2
    live = [k for k in range(4) if (k & 1) < nch]
    assert np.array_equal(gc_h[:, live], gc_d[:, live]), "granule-channel descriptors"
    return n


@pytest.mark.parametrize("name", list(VARIANTS))
def test_device_side_info_equals_host_parser(ctx, name):
    s, _ = H.synth(400, seed=41, **VARIANTS[name])
    assert compare_descriptors(ctx, s) >= 398


def test_stale_count1_chain_on_device(ctx):
    s, _ = H.synth(600, seed=42, **H.CONFIGS["cfg4_vbr_mixed"])
    t = empty_some_parts(s)
    import pdmp3_b200
    g = pdmp3_b200.parse_stream(t, lookahead=0).gcs()
    assert (g[:, :, 3] != 0).sum() > 20, "the test stream must contain empty parts"
    compare_descriptors(ctx, t)
    # batches: an empty part at the start of a batch refers to the carried state (w3 = 0x7fffffff)
    compare_descriptors(ctx, t[int(pdmp3_b200.parse_stream(t, lookahead=0).frames()["main_off"][3]) - 36:])


@pytest.mark.parametrize("seed", range(4))
def test_corrupted_side_info(ctx, seed):
    rng = np.random.default_rng(seed)
    s, _ = H.synth(300, seed=50 + seed, **H.CONFIGS["cfg4_vbr_mixed"])
    t = s.copy()
    for pos in rng.integers(0, len(t), size=len(t) // 300):
        t[pos] ^= 1 << int(rng.integers(0, 8))
    import pdmp3_b200
    p = pdmp3_b200.parse_stream(t, lookahead=0)
    if p.n_frames < 10: pytest.skip("corruption hit the first header")
    compare_descriptors(ctx, t[: int(p.consumed)])


@pytest.mark.parametrize("mode", ["fast", "exact"])
@pytest.mark.parametrize("name", ["cfg3", "cfg4", "mono", "k48", "crc", "garbage"])
def test_decode_same_pcm_either_parser(name, mode):
    import pdmp3_b200
    s, _ = H.synth(260, seed=43, **VARIANTS[name])
    if name == "cfg4": s = empty_some_parts(s)
    c = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST if mode == "fast" else pdmp3_b200.MODE_EXACT)
    a = c.decode(s, lookahead=0)
    c.reset(); b = c.decode(s, lookahead=0, hop_only=True)
    c.close()
    assert np.array_equal(a, b)


def test_streaming_api_parses_on_the_device_by_default():
    """pdmp3_read() with the default options (device side info) == the same handle with sideinfo=host"""
    import pdmp3_b200
    s, _ = H.synth(500, seed=44, **H.CONFIGS["cfg4_vbr_mixed"])
    outs = []
    for opt in ("b200:ring=2000000", "b200:ring=2000000,sideinfo=host"):
        d = pdmp3_b200.Decoder(opt); d.open_feed(); d.feed(s)
        rc, out = d.read(500 * 4608)
        d.close(); outs.append(out.copy())
    assert len(outs[0]) == len(outs[1]) and len(outs[0]) >= 495 * 4608       # the 1152-byte look-ahead rule keeps the last frames back
    assert np.array_equal(outs[0], outs[1])
