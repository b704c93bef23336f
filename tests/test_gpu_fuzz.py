"""Robustness: corrupted streams (bit flips in headers, side info and main data, truncation, junk) must
decode without a CUDA error or a crash, in both modes and through the streaming API.  No parity claim:
such streams are outside the envelope where the reference itself is well defined."""
import numpy as np, pytest
import p3harness as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(6))
def test_bit_flips_do_not_break_the_decoder(seed):
    import pdmp3_b200
    rng = np.random.default_rng(seed)
    s, _ = H.synth(120, seed=100 + seed, **H.CONFIGS["cfg4_vbr_mixed"])
    t = s.copy()
    n = len(t)
    for pos in rng.integers(0, n, size=n // 40):              # ~2.5 % of the bytes get a flipped bit
        t[pos] ^= 1 << int(rng.integers(0, 8))
    t = t[: n - int(rng.integers(0, 700))]                    # and the tail is cut somewhere
    for mode in (pdmp3_b200.MODE_FAST, pdmp3_b200.MODE_EXACT):
        ctx = pdmp3_b200.Context(0, mode)
        pcm = ctx.decode(t, lookahead=0)
        assert pcm.shape[1:] == (1152, pcm.shape[2]) and pcm.shape[0] > 20
        good = ctx.decode(s, lookahead=0)                      # the context is still healthy afterwards
        assert good.shape[0] == 120
        ctx.close()
    d = pdmp3_b200.Decoder(); d.open_feed()
    fed = 0
    for _ in range(400):
        rc, out = d.read(16384)
        if rc == pdmp3_b200.PDMP3_ERR: break
        if rc == pdmp3_b200.PDMP3_NEED_MORE:
            if fed >= len(t): break
            d.feed(t[fed:fed + 4096]); fed += 4096
    d.close()


def test_oracle_survives_the_same_garbage():
    rng = np.random.default_rng(7)
    s, _ = H.synth(60, seed=3, **H.CONFIGS["cfg4_vbr_mixed"])
    t = s.copy()
    for pos in rng.integers(0, len(t), size=len(t) // 40): t[pos] ^= 1 << int(rng.integers(0, 8))
    o = H.oracle_decode(t, lookahead=0, taps=False)
    assert o["n_frames"] > 10
