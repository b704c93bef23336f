"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares;
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C, os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("pdmp3.h", "pdmp3_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(pdmp3(?:_[a-z_]+)?|p3_[a-z0-9_]+)\s*\(", src))
    return sorted(n for n in names if not n.startswith("p3_gc") and n not in ("p3_frame", "p3_ctx"))


def test_library_exports_every_declared_symbol():
    import pdmp3_b200
    L = pdmp3_b200.lib()
    syms = declared_symbols()
    assert {"pdmp3_new", "pdmp3_read", "pdmp3_feed", "pdmp3_decode", "pdmp3_getformat", "pdmp3_open_feed", "pdmp3_delete",
            "pdmp3", "p3_parse", "p3_decode_batch", "p3_ctx_create"} <= set(syms)
    for s in syms:
        assert hasattr(L, s), s


def test_no_cpu_fallback():
    import torch, pdmp3_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pdmp3_b200.P3Error):
        pdmp3_b200.Context(0)
    # the streaming API accepts data but cannot decode without a device: PDMP3_ERR, never CPU output
    import p3harness as H
    s, _ = H.synth(8, seed=1)
    d = pdmp3_b200.Decoder()
    assert d.open_feed() == 0 and d.feed(s[:4096]) == 0
    rc, out = d.read(16384)
    assert rc == pdmp3_b200.PDMP3_ERR and len(out) == 0


def test_product_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", os.path.join(ROOT, "pdmp3_b200", "libpdmp3_b200.so")], stdout=subprocess.PIPE, text=True).stdout
    assert "p3o_" not in out and "ref_taps" not in out
    for f in os.listdir(os.path.join(ROOT, "pdmp3_b200")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "pdmp3_b200", f)).read().replace("oracle/", "").replace("oracle restatement", "").replace("(oracle", "") or f == "build.py"
