"""ISO mode (P3_FRAME_ISO): the oracle's ISO switch checked against things that do not depend on it --
the spectra the forward encoder put into the stream (count1 table B, empty parts) and a numpy restatement of the
ISO stereo rules.  The reference itself cannot pin this mode: it is exactly where pdmp3.c deviates (Q1-Q4, Q6)."""
import numpy as np
import p3harness as H

ISO = dict(iso=1, mode=1, mode_ext=-1, blocks=1, count1_b_pm=400, bitrate_index=11)


import pytest


@pytest.mark.parametrize("kw", [dict(), dict(sfreq=1), dict(sfreq=2, bitrate_index=8), dict(mode=3, bitrate_index=6), dict(bitrate_index=0),
                                dict(crc=1, garbage_pm=100)], ids=["44k", "48k", "32k", "mono", "vbr", "crc_junk"])
def test_table_b_and_empty_parts_round_trip(kw):
    cfg = dict(ISO); cfg.update(kw)
    s, enc = H.synth(120, seed=5, want_is=True, **cfg)
    o = H.oracle_decode(s, lookahead=0, iso=True)
    n = o["n_frames"]; nch = 1 if cfg["mode"] == 3 else 2
    f = H.gc_fields(o["gcs"]).reshape(n, 2, 2, 20)[:, :, :nch]
    assert (f[..., 17] == 1).any() and (f[..., 0] == 0).any(), "stream must hold table-B granules and empty parts"
    assert np.array_equal(o["is_huff"][:, :, :nch], enc[:n, :, :nch]), "decoded spectra != encoded spectra"
    assert (o["count1"][:, :, :nch][f[..., 0] == 0] == 0).all(), "an empty part has count1 = 0 in ISO mode"
    if kw:
        return
    # the same stream decoded the reference's way differs (table B quads are garbage there, Q1)
    c = H.oracle_decode(s, lookahead=0, iso=False)
    assert not np.array_equal(c["is_huff"], enc[:n])


def _iso_stereo(T, fr, gcf, c1, scf_l, scf_s, l, r):
    """ISO stereo of one granule, straight from the rules stated in include/pdmp3_b200.h (P3_FRAME_ISO)."""
    l, r = l.copy(), r.copy()
    if fr["mode"] != 1 or fr["mode_ext"] == 0:
        return l, r
    sfb_l, sfb_s, is_l, is_r = T
    done = np.zeros(576, bool)
    c1r = int(c1[1])
    if fr["mode_ext"] & 1:
        short = gcf[0][4] == 1 and gcf[0][5] == 2
        mixed = short and gcf[0][6] == 1
        if not short or mixed:
            for sfb in range(8 if mixed else 21):
                a, b = sfb_l[sfb], sfb_l[sfb + 1]
                p = int(scf_l[1][sfb])
                if a >= c1r and p < 7:
                    x = l[a:b].copy(); l[a:b] = is_l[p] * x; r[a:b] = is_r[p] * x; done[a:b] = True
        if short:
            for sfb in range(3 if mixed else 0, 12):
                a, wl = 3 * sfb_s[sfb], sfb_s[sfb + 1] - sfb_s[sfb]
                for w in range(3):
                    p = int(scf_s[1][sfb][w])
                    if a >= c1r and p < 7:
                        lo = a + wl * w
                        x = l[lo:lo + wl].copy(); l[lo:lo + wl] = is_l[p] * x; r[lo:lo + wl] = is_r[p] * x; done[lo:lo + wl] = True
    if fr["mode_ext"] & 2:
        m = ~done & (np.arange(576) < max(int(c1[0]), c1r))
        a, b = l[m] + r[m], l[m] - r[m]
        l[m] = (a.astype(np.float64) * 0.70710678118654752440).astype(np.float32)
        r[m] = (b.astype(np.float64) * 0.70710678118654752440).astype(np.float32)
    return l, r


def test_iso_stereo_matches_numpy_restatement():
    import math
    s, _ = H.synth(100, seed=9, **dict(ISO, count1_b_pm=0))
    o = H.oracle_decode(s, lookahead=0, iso=True)
    fr, f = o["frames"], H.gc_fields(o["gcs"]).reshape(o["n_frames"], 2, 2, 20)
    sfb_l = [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576]
    sfb_s = [0, 4, 8, 12, 16, 22, 30, 40, 52, 66, 84, 106, 136, 192]
    ratio = [np.float32(float("%f" % math.tan(i * math.pi / 12))) for i in range(6)]          # is_ratios (pdmp3.c:575)
    is_l = [np.float32(ratio[i] / (np.float32(1) + ratio[i])) for i in range(6)] + [np.float32(1)]
    is_r = [np.float32(np.float32(1) / (np.float32(1) + ratio[i])) for i in range(6)] + [np.float32(0)]
    T = (sfb_l, sfb_s, is_l, is_r)
    seen_is = seen_short_is = 0
    for k in range(o["n_frames"]):
        for gr in range(2):
            l, r = _iso_stereo(T, fr[k], f[k, gr], o["count1"][k, gr], o["scf_l"][k, gr], o["scf_s"][k, gr],
                               o["xr_reo"][k, gr, 0], o["xr_reo"][k, gr, 1])
            assert np.array_equal(l, o["xr_ste"][k, gr, 0]) and np.array_equal(r, o["xr_ste"][k, gr, 1]), (k, gr)
            if fr[k]["mode_ext"] & 1 and not np.array_equal(o["xr_reo"][k, gr, 1], o["xr_ste"][k, gr, 1]):
                seen_is += 1; seen_short_is += int(f[k, gr, 0, 5] == 2)
    assert seen_is > 10 and seen_short_is > 0, (seen_is, seen_short_is)


def test_compat_mode_is_untouched_by_the_switch():
    """iso=False leaves every descriptor and result as before (the flag is the only difference in the descriptors)"""
    s, _ = H.synth(60, seed=3, **H.CONFIGS["cfg4_vbr_mixed"])
    a, b = H.oracle_decode(s, iso=False), H.oracle_decode(s, iso=True)
    assert (a["frames"]["flags"] & 16 == 0).all() and (b["frames"]["flags"] & 16 == 16).all()
    assert np.array_equal(a["gcs"][..., :3], b["gcs"][..., :3])
