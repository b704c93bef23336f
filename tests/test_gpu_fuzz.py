"""Robustness: corrupted streams (bit flips in headers, side info and main data, truncation, junk) must
decode without a CUDA error or a crash, in both modes and through the streaming API.  No parity claim:
such streams are outside the envelope where the reference itself is well defined."""
import numpy as np, pytest
import p3harness as H

pytestmark = pytest.mark.gpu


def decode_all(ctx, data):
    """batch after batch until the parser finds nothing more (a corrupt header may end a batch: stop=3)"""
    import pdmp3_b200
    pos, frames, guard = 0, 0, 0
    while pos < len(data) and guard < 500:
        guard += 1
        p = pdmp3_b200.parse_stream(data[pos:], lookahead=0)
        if p.n_frames == 0:
            if p.stop == 2: pos += 1152                       # junk: skip what the reference would scan (pdmp3.c:1337)
            elif p.stop == 3: pos += max(int(p.consumed), 1)
            else: break
            continue
        pcm = ctx.decode_parsed(p)
        assert pcm.shape[0] == p.n_pcm_frames
        frames += p.n_frames; pos += int(p.consumed)
    return frames


@pytest.mark.parametrize("seed", range(6))
def test_bit_flips_do_not_break_the_decoder(seed):
    import pdmp3_b200
    rng = np.random.default_rng(seed)
    s, _ = H.synth(120, seed=100 + seed, **H.CONFIGS["cfg4_vbr_mixed"])
    t = s.copy()
    n = len(t)
    rate = (400, 40)[seed % 2]                                 # 0.25 % or 2.5 % of the bytes get a flipped bit
    for pos in rng.integers(0, n, size=n // rate):
        t[pos] ^= 1 << int(rng.integers(0, 8))
    t = t[: n - int(rng.integers(0, 700))]                    # and the tail is cut somewhere
    for mode in (pdmp3_b200.MODE_FAST, pdmp3_b200.MODE_EXACT):
        ctx = pdmp3_b200.Context(0, mode)
        assert decode_all(ctx, t) > 10
        ctx.reset()
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
    for pos in rng.integers(0, len(t), size=len(t) // 400): t[pos] ^= 1 << int(rng.integers(0, 8))
    o = H.oracle_decode(t, lookahead=0, taps=False)
    assert o["n_frames"] > 10
