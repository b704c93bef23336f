"""ISO mode (P3_FRAME_ISO) on the GPU against the oracle's ISO switch (itself pinned by tests/test_cpu_iso.py):
bit-exact at every stage in EXACT mode, within 1 LSB in FAST mode (both synthesis kernels), the device side-info parser
and the streaming API ("b200:iso") included."""
import numpy as np, pytest
import p3harness as H
from test_gpu_parity import feq
from test_gpu_api import cli_loop

pytestmark = pytest.mark.gpu

ISO_VARIANTS = dict(
    tabB_js=dict(iso=1, mode=1, mode_ext=-1, blocks=1, count1_b_pm=400, bitrate_index=11),
    vbr=dict(iso=1, mode=1, mode_ext=-1, blocks=1, count1_b_pm=200, bitrate_index=0),
    is_long=dict(iso=1, mode=1, mode_ext=3, blocks=0, bitrate_index=7),
    k48=dict(iso=1, sfreq=1, mode=1, mode_ext=-1, blocks=1, count1_b_pm=300, bitrate_index=9),
    mono=dict(iso=1, mode=3, blocks=1, count1_b_pm=500, bitrate_index=6),
    compat_stream=dict(H.CONFIGS["cfg4_vbr_mixed"]),          # a stream from the reference's envelope decodes in ISO mode too
)


@pytest.mark.parametrize("name", list(ISO_VARIANTS))
def test_iso_exact_mode_bit_exact_vs_oracle(gpu_ctx, name):
    s, enc = H.synth(150, seed=41, want_is=True, **ISO_VARIANTS[name])
    o = H.oracle_decode(s, lookahead=1152, iso=True)
    gpu_ctx.reset()
    pcm, t = gpu_ctx.decode(s, lookahead=1152, taps=True, iso=True)
    n = o["n_frames"]; nch = pcm.shape[2]
    assert pcm.shape[0] == n
    assert np.array_equal(t["is_huff"][:, :, :nch], o["is_huff"][:, :, :nch]), "Huffman output"
    if ISO_VARIANTS[name].get("iso"):
        assert np.array_equal(t["is_huff"][:, :, :nch], enc[:n, :, :nch]), "decoded spectra != encoded spectra"
    assert np.array_equal(t["count1"][:, :, :nch], o["count1"][:, :, :nch]), "count1"
    assert feq(t["xr"][:, :, :nch], o["xr_ali"][:, :, :nch]).all(), "requantize/reorder/stereo/antialias"
    assert np.array_equal(pcm, o["pcm"]), "PCM"
    if name != "mono" and ISO_VARIANTS[name].get("iso"):
        c = H.oracle_decode(s, lookahead=1152, iso=False)
        assert not np.array_equal(c["pcm"], o["pcm"]), "the stream does not exercise the switch"


@pytest.mark.parametrize("name", list(ISO_VARIANTS))
def test_iso_fast_mode_within_one_lsb(name):
    import pdmp3_b200
    s, _ = H.synth(300, seed=43, **ISO_VARIANTS[name])
    o = H.oracle_decode(s, lookahead=0, iso=True)
    c = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST)
    try:
        got = {}
        for which in (0, 1):                                   # k_synth_warp (stereo default) / k_synth_fast
            c.reset(); c.set_synth_kernel(which); got[which] = c.decode(s, lookahead=0, iso=True)
        c.set_synth_kernel(0); c.reset(); c.set_frames_per_cta(3); short_runs = c.decode(s, lookahead=0, iso=True)
        c.reset(); hop = c.decode(s, lookahead=0, iso=True, hop_only=True)     # side info parsed on the device
    finally:
        c.close()
    for which in (0, 1):
        d = np.abs(got[which].astype(np.int32) - o["pcm"].astype(np.int32))
        assert d.max() <= 1, "kernel %d: max |diff| = %d LSB" % (which, d.max())
        assert (d == 0).mean() > 0.90
    assert np.array_equal(got[0], short_runs), "result depends on the run length"
    assert np.array_equal(got[0], hop), "device side-info parse differs"


def test_iso_streaming_api():
    """pdmp3_new("b200:iso,mode=exact") through the reference's CLI loop == the oracle in ISO mode, PCM bit for bit"""
    import pdmp3_b200
    s, _ = H.synth(200, seed=47, **ISO_VARIANTS["tabB_js"])
    o = H.oracle_decode(s, lookahead=1152, iso=True)
    for opts in ("b200:iso,mode=exact", "b200:iso,mode=exact,sideinfo=host"):
        d = pdmp3_b200.Decoder(opts)
        pcm, _ = cli_loop(d, s)
        d.close()
        got = pcm.view(np.int16).reshape(-1, 1152, 2)
        assert got.shape[0] == o["n_frames"]
        assert np.array_equal(got, o["pcm"]), opts
