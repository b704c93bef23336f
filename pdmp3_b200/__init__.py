"""pdmp3_b200 -- B200-native MPEG-1 Layer III decoder behind the pdmp3_* (libmpg123-subset) API.

The product is the C-ABI shared library pdmp3_b200/libpdmp3_b200.so (plain-C host side +
hand-written sm_100a CUDA kernels).  This Python package is only a ctypes binding to it:
  * `Decoder`  mirrors the reference's streaming API one to one (pdmp3.c:2351-2535),
  * `Context`  exposes the batch C-ABI of include/pdmp3_b200.h (used by tests and bench.py),
  * `Dist`     the frame-sharded multi-GPU decode (BASELINE configs[4]; NCCL is called from the C side).
There is NO CPU fallback: if the library or a sm_100 GPU is missing, calls raise.
"""
from ._binding import (lib, Context, Decoder, Dist, dist_unique_id, parse_stream, P3Error,
                       PDMP3_OK, PDMP3_ERR, PDMP3_NEED_MORE, PDMP3_NEW_FORMAT, PDMP3_NO_SPACE,
                       PDMP3_ENC_SIGNED_16, MODE_EXACT, MODE_FAST)

__all__ = ["lib", "Context", "Decoder", "Dist", "dist_unique_id", "parse_stream", "P3Error", "PDMP3_OK", "PDMP3_ERR", "PDMP3_NEED_MORE",
           "PDMP3_NEW_FORMAT", "PDMP3_NO_SPACE", "PDMP3_ENC_SIGNED_16", "MODE_EXACT", "MODE_FAST"]
