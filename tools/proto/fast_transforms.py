"""Numerical prototype (float64) of the fast transforms used by the FAST-mode kernel.
Checks the index/sign algebra against the direct formulas the reference evaluates
(pdmp3.c:1689-1698 IMDCT, 1990-1993/2010-2026 polyphase)."""
import numpy as np
pi = np.pi

# ---------- IMDCT-36 through an 18-point DCT-IV ----------
def imdct36_direct(x):
    p = np.arange(36)[:, None]; m = np.arange(18)[None, :]
    return (np.cos(pi / 72 * (2 * p + 19) * (2 * m + 1)) * x[None, :]).sum(1)

def dct4_18(x):
    k = np.arange(18)[:, None]; m = np.arange(18)[None, :]
    return (np.cos(pi / 18 * (k + 0.5) * (m + 0.5)) * x[None, :]).sum(1)

def imdct36_fast(x):
    t = dct4_18(x)
    y = np.empty(36)
    y[0:9] = t[9:18]
    y[9:18] = -t[17:8:-1]
    y[18:27] = -t[8::-1]
    y[27:36] = -t[0:9]
    return y

x = np.random.randn(18)
assert np.allclose(imdct36_direct(x), imdct36_fast(x), atol=1e-12)

# ---------- DCT-II 32 (Lee) and the 64-entry matrixing vector ----------
def matrixing_direct(s):
    i = np.arange(64)[:, None]; j = np.arange(32)[None, :]
    return (np.cos((16 + i) * (2 * j + 1) * pi / 64) * s[None, :]).sum(1)

def dct2_direct(s, n):
    k = np.arange(n)[:, None]; j = np.arange(n)[None, :]
    return (np.cos((2 * j + 1) * k * pi / (2 * n)) * s[None, :]).sum(1)

def dct2_lee(x):
    """X[k] = sum_j x[j] cos((2j+1)k pi/(2n)), recursive Lee decomposition."""
    n = len(x)
    if n == 1:
        return x.copy()
    h = n // 2
    a = x[:h] + x[::-1][:h]
    b = (x[:h] - x[::-1][:h]) / (2 * np.cos((2 * np.arange(h) + 1) * pi / (2 * n)))
    A = dct2_lee(a); B = dct2_lee(b)
    X = np.empty(n)
    X[0::2] = A
    X[1::2] = B + np.append(B[1:], 0.0)
    return X

s = np.random.randn(32)
assert np.allclose(dct2_direct(s, 32), dct2_lee(s), atol=1e-10)

def matrixing_from_X(X):
    V = np.empty(64)
    V[0:16] = X[16:32]
    V[16] = 0.0
    i = np.arange(17, 48); V[17:48] = -X[48 - i]
    i = np.arange(48, 64); V[48:64] = -X[i - 48]
    return V

assert np.allclose(matrixing_direct(s), matrixing_from_X(dct2_direct(s, 32)), atol=1e-10)

# ---------- window: pcm(t)[j] = sum_k D[32k+j] * V(t-k)[(k odd)*32 + j]  expressed on X ----------
def idx_sign(j, odd):
    """V[(odd?32:0)+j] = sign * X[idx] (sign 0 -> zero)"""
    i = j + (32 if odd else 0)
    if i < 16: return 16 + i, 1.0
    if i == 16: return 0, 0.0
    if i < 48: return 48 - i, -1.0
    return i - 48, -1.0

T = 40
S = np.random.randn(T, 32); D = np.random.randn(512)
V = np.array([matrixing_direct(S[t]) for t in range(T)]); X = np.array([dct2_direct(S[t], 32) for t in range(T)])
for t in range(15, T):
    for j in range(32):
        ref = sum(D[32 * k + j] * V[t - k][(32 if k & 1 else 0) + j] for k in range(16))
        acc = 0.0
        for k in range(16):
            ix, sg = idx_sign(j, k & 1)
            acc += D[32 * k + j] * sg * X[t - k][ix]
        assert abs(ref - acc) < 1e-9
print("fast transform algebra OK")

# ---------- 18-point DCT-IV through two 9-point DCT-IIs (stage D of k_synth_fast) ----------
def dct9(x):
    c = lambda deg: np.cos(np.deg2rad(deg))
    s0, s1, s2, s3, x4 = x[0] + x[8], x[1] + x[7], x[2] + x[6], x[3] + x[5], x[4]
    d0, d1, d2, d3 = x[0] - x[8], x[1] - x[7], x[2] - x[6], x[3] - x[5]
    X = np.empty(9)
    X[0] = s0 + s1 + s2 + s3 + x4
    X[2] = s0 * c(20) + s1 * 0.5 - s2 * c(80) - s3 * c(40) - x4
    X[4] = s0 * c(40) - s1 * 0.5 - s2 * c(20) + s3 * c(80) + x4
    X[6] = (s0 + s2 + s3) * 0.5 - s1 - x4
    X[8] = s0 * c(80) - s1 * 0.5 + s2 * c(40) - s3 * c(20) + x4
    X[1] = d0 * c(10) + d1 * c(30) + d2 * c(50) + d3 * c(70)
    X[3] = (d0 - d2 - d3) * c(30)
    X[5] = d0 * c(50) - d1 * c(30) - d2 * c(70) + d3 * c(10)
    X[7] = d0 * c(70) - d1 * c(30) + d2 * c(10) - d3 * c(50)
    return X

x9 = np.random.randn(9)
assert np.allclose(dct9(x9), dct2_direct(x9, 9), atol=1e-12)

def dct4_18_fast(x):
    m = np.arange(18)
    y = x * 2 * np.cos(pi * (2 * m + 1) / 72)
    a = y[:9] + y[::-1][:9]
    b = (y[:9] - y[::-1][:9]) / (2 * np.cos(pi * (2 * np.arange(9) + 1) / 36))
    A, B = dct9(a), dct9(b)
    Y = np.empty(18); Y[0::2] = A; Y[1::2] = B + np.append(B[1:], 0.0)
    t = np.empty(18); t[0] = Y[0] / 2
    for k in range(1, 18): t[k] = Y[k] - t[k - 1]
    return t

x = np.random.randn(18)
assert np.allclose(dct4_18(x), dct4_18_fast(x), atol=1e-10)
# single precision error of the recursion, relative to the output scale
xf = np.random.randn(2000, 18).astype(np.float32)
def f32(v): return np.asarray(v, dtype=np.float32)
err = 0
for r in xf[:200]:
    ref = dct4_18(r.astype(np.float64))
    m = np.arange(18)
    y = f32(r * f32(2 * np.cos(pi * (2 * m + 1) / 72)))
    a = f32(y[:9] + y[::-1][:9]); b = f32(f32(y[:9] - y[::-1][:9]) * f32(1 / (2 * np.cos(pi * (2 * np.arange(9) + 1) / 36))))
    A, B = f32(dct9(a.astype(np.float64))), f32(dct9(b.astype(np.float64)))
    Y = np.empty(18, np.float32); Y[0::2] = A; Y[1::2] = f32(B + np.append(B[1:], np.float32(0)))
    t = np.empty(18, np.float32); t[0] = Y[0] * np.float32(0.5)
    for k in range(1, 18): t[k] = np.float32(Y[k] - t[k - 1])
    err = max(err, np.abs(t - ref).max() / np.abs(ref).max())
print("fast DCT-IV-18 OK; fp32 relative error (max over 200 vectors): %.2e" % err)
