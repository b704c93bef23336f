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
