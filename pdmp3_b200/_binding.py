"""ctypes binding to libpdmp3_b200.so (see include/pdmp3.h and include/pdmp3_b200.h)."""
import ctypes as C
import os
import numpy as np

PDMP3_OK, PDMP3_ERR, PDMP3_NEED_MORE, PDMP3_NEW_FORMAT, PDMP3_NO_SPACE = 0, -1, -10, -11, 7
PDMP3_ENC_SIGNED_16 = 0x80 | 0x40 | 0x10
MODE_EXACT, MODE_FAST = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.environ.get("P3_LIB") or os.path.join(_HERE, "libpdmp3_b200.so")     # P3_LIB: kernel-variant experiments


class P3Error(RuntimeError):
    pass


class P3Frame(C.Structure):
    _fields_ = [("main_off", C.c_uint64), ("main_pos", C.c_uint64), ("main_size", C.c_uint16), ("main_begin", C.c_uint16),
                ("nch", C.c_uint8), ("mode", C.c_uint8), ("mode_ext", C.c_uint8), ("sfreq", C.c_uint8),
                ("scfsi", C.c_uint8), ("flags", C.c_uint8), ("bitrate_kbps", C.c_uint16), ("pcm_index", C.c_uint32)]


class P3Gc(C.Structure):
    _fields_ = [("w0", C.c_uint32), ("w1", C.c_uint32), ("w2", C.c_uint32), ("w3", C.c_uint32)]


class P3ParseState(C.Structure):
    _fields_ = [("main_pos", C.c_uint64), ("top", C.c_uint32), ("pcm_index", C.c_uint32), ("nch", C.c_int32), ("sfreq", C.c_int32)]


class P3ParseOpts(C.Structure):
    _fields_ = [("max_frames", C.c_int64), ("lookahead", C.c_uint32), ("nthreads", C.c_int32), ("warmup_frames", C.c_uint32), ("hop_only", C.c_uint32), ("iso", C.c_uint32)]


class P3Parsed(C.Structure):
    _fields_ = [("n_frames", C.c_int64), ("frames", C.POINTER(P3Frame)), ("gcs", C.POINTER(P3Gc)),
                ("consumed", C.c_uint64), ("n_pcm_frames", C.c_int64), ("external", C.c_int32), ("stop", C.c_int32),
                ("hop_only", C.c_int32), ("pad_", C.c_int32)]


class P3ShardResult(C.Structure):
    _fields_ = [("n_frames_total", C.c_int64), ("n_frames_mine", C.c_int64), ("warmup_mine", C.c_int64), ("chunks", C.c_int64),
                ("nch", C.c_int32), ("stop", C.c_int32), ("launches", C.c_int32), ("pad_", C.c_int32),
                ("consumed", C.c_uint64), ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64), ("ms", C.c_float), ("ms_scatter", C.c_float),
                ("ms_staged", C.c_float), ("ms_decoded", C.c_float)]


class P3Taps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("is_huff", "count1", "scf", "xr", "y")]


FRAME_DT = np.dtype([("main_off", "<u8"), ("main_pos", "<u8"), ("main_size", "<u2"), ("main_begin", "<u2"),
                     ("nch", "u1"), ("mode", "u1"), ("mode_ext", "u1"), ("sfreq", "u1"), ("scfsi", "u1"),
                     ("flags", "u1"), ("bitrate_kbps", "<u2"), ("pcm_index", "<u4")])

_lib = None


def lib():
    """Load the product library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise P3Error("%s not found: build it with `python -m pdmp3_b200.build` (there is no CPU fallback)" % _LIBPATH)
    L = C.CDLL(_LIBPATH)
    L.p3_last_error.restype = C.c_char_p
    L.p3_parse.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(P3ParseOpts), C.POINTER(P3ParseState), C.POINTER(P3Parsed)]
    L.p3_parsed_free.argtypes = [C.POINTER(P3Parsed)]
    L.p3_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.p3_ctx_destroy.argtypes = [C.c_void_p]
    L.p3_ctx_reset.argtypes = [C.c_void_p]
    L.p3_ctx_set_mode.argtypes = [C.c_void_p, C.c_int]
    L.p3_ctx_set_taps.argtypes = [C.c_void_p, C.c_int]
    L.p3_ctx_set_frames_per_cta.argtypes = [C.c_void_p, C.c_int]
    L.p3_ctx_set_synth_kernel.argtypes = [C.c_void_p, C.c_int]
    L.p3_ctx_set_overlap.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int]
    L.p3_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(P3Parsed), C.c_void_p, C.POINTER(P3Taps)]
    L.p3_synth_from_xr.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(P3Parsed), C.c_void_p]
    L.p3_batch_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(P3Parsed)]
    L.p3_batch_upload_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(P3ParseOpts), C.POINTER(P3ParseState), C.POINTER(P3Parsed)]
    L.p3_decode_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(P3ParseOpts), C.POINTER(P3ParseState), C.POINTER(P3Parsed),
                                C.c_void_p, C.c_int64, C.POINTER(P3Taps)]
    L.p3_hop_rounds.argtypes = [C.c_void_p]
    L.p3_hop_ms.argtypes = [C.c_void_p]
    L.p3_hop_ms.restype = C.c_float
    L.p3_batch_channels.argtypes = [C.c_void_p]
    L.p3_dist_unique_id.argtypes = [C.c_void_p]
    L.p3_dist_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.p3_dist_destroy.argtypes = [C.c_void_p]
    L.p3_dist_gather_transport.argtypes = [C.c_void_p]
    L.p3_sharded_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(P3ParseOpts), C.c_int64, C.POINTER(P3ShardResult)]
    L.p3_dist_measure_ingest.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float)]
    L.p3_batch_run.argtypes = [C.c_void_p]
    L.p3_batch_sync.argtypes = [C.c_void_p]
    L.p3_batch_download.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(P3Taps)]
    L.p3_batch_download_desc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.p3_batch_pcm_device.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.p3_batch_pcm_device.restype = C.c_void_p
    L.p3_ctx_stream.argtypes = [C.c_void_p]
    L.p3_ctx_stream.restype = C.c_void_p
    L.p3_batch_time.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.p3_batch_time_xr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.p3_kernel_launch_count.argtypes = [C.c_void_p]
    for name in ("pdmp3_new", "pdmp3_delete", "pdmp3_open_feed", "pdmp3_feed", "pdmp3_read", "pdmp3_decode", "pdmp3_getformat"):
        if not hasattr(L, name):
            break
    else:
        L.pdmp3_new.restype = C.c_void_p
        L.pdmp3_new.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.pdmp3_delete.argtypes = [C.c_void_p]
        L.pdmp3_open_feed.argtypes = [C.c_void_p]
        L.pdmp3_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.pdmp3_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.pdmp3_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.pdmp3_getformat.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise P3Error("%s failed (%d): %s" % (what, rc, lib().p3_last_error().decode(errors="replace")))


class Parsed:
    """Owns a p3_parsed (host descriptors of one batch)."""

    def __init__(self, stream, lookahead=0, max_frames=0, warmup=0, nthreads=0, state=None, hop_only=False, iso=False):
        self.stream = np.ascontiguousarray(stream, dtype=np.uint8)
        self.c = P3Parsed()
        self.state = state if state is not None else P3ParseState(0, 0, 0, -1, -1)
        o = P3ParseOpts(max_frames, lookahead, nthreads, warmup, 1 if hop_only else 0, 1 if iso else 0)   # iso: P3_FRAME_ISO on every frame
        _check(lib().p3_parse(self.stream.ctypes.data, len(self.stream), C.byref(o), C.byref(self.state), C.byref(self.c)), "p3_parse")

    n_frames = property(lambda s: s.c.n_frames)
    n_pcm_frames = property(lambda s: s.c.n_pcm_frames)
    consumed = property(lambda s: s.c.consumed)
    stop = property(lambda s: s.c.stop)

    @property
    def nch(self):
        return int(self.c.frames[0].nch) if self.c.n_frames else 2

    def frames(self):
        n = self.c.n_frames
        return np.frombuffer(C.string_at(self.c.frames, 32 * n), dtype=FRAME_DT).copy() if n else np.zeros(0, FRAME_DT)

    def gcs(self):
        n = self.c.n_frames
        return np.frombuffer(C.string_at(self.c.gcs, 64 * n), dtype=np.uint32).reshape(n, 4, 4).copy() if n else np.zeros((0, 4, 4), np.uint32)

    def __del__(self):
        try:
            lib().p3_parsed_free(C.byref(self.c))
        except Exception:
            pass


def parse_stream(stream, **kw):
    return Parsed(stream, **kw)


class Context:
    """Device context of the batch C-ABI (p3_ctx)."""

    def __init__(self, device=0, mode=MODE_EXACT):
        self.h = C.c_void_p()
        _check(lib().p3_ctx_create(device, C.byref(self.h)), "p3_ctx_create")
        self.set_mode(mode)

    def set_mode(self, mode):
        _check(lib().p3_ctx_set_mode(self.h, mode), "p3_ctx_set_mode")

    def set_frames_per_cta(self, n):
        _check(lib().p3_ctx_set_frames_per_cta(self.h, n), "p3_ctx_set_frames_per_cta")

    def set_synth_kernel(self, which):
        """FAST mode: 0 = k_synth_warp / k_synth_warp_lean by content class (default), 1 = always k_synth_fast, 2 = k_synth_warp only (no classes)."""
        _check(lib().p3_ctx_set_synth_kernel(self.h, which), "p3_ctx_set_synth_kernel")

    def set_overlap(self, chunk_frames, prio=0, synth_pad=0, k1_pad=0):
        """FAST mode: K0 + K1 of chunk i+1 under the synthesis of chunk i (0 = off); call before upload (sizes the intermediates)."""
        _check(lib().p3_ctx_set_overlap(self.h, chunk_frames, prio, synth_pad, k1_pad), "p3_ctx_set_overlap")

    def reset(self):
        _check(lib().p3_ctx_reset(self.h), "p3_ctx_reset")

    def close(self):
        if self.h:
            lib().p3_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _taps(self, n):
        arrs = dict(is_huff=np.zeros((n, 2, 2, 576), np.int16), count1=np.zeros((n, 2, 2), np.int32),
                    scf=np.zeros((n, 2, 2, 64), np.uint8), xr=np.zeros((n, 2, 2, 576), np.float32),
                    y=np.zeros((n, 2, 2, 576), np.float32))
        return arrs, P3Taps(**{k: a.ctypes.data for k, a in arrs.items()})

    def decode_parsed(self, parsed, taps=False):
        """Host buffers in, host buffers out (copies inside). -> pcm [n_pcm_frames,1152,nch] (+ taps dict)."""
        n = parsed.n_frames
        pcm = np.zeros((parsed.n_pcm_frames, 1152, parsed.nch), np.int16)
        arrs, t = self._taps(n) if taps else ({}, None)
        _check(lib().p3_decode_batch(self.h, parsed.stream.ctypes.data, len(parsed.stream), C.byref(parsed.c),
                                     pcm.ctypes.data, C.byref(t) if taps else None), "p3_decode_batch")
        if taps:
            # y is produced slot-major [18][32]; present it in the reference's [sb][18] layout
            arrs["y"] = np.ascontiguousarray(arrs["y"].reshape(n, 2, 2, 18, 32).transpose(0, 1, 2, 4, 3)).reshape(n, 2, 2, 576)
            arrs["scf_l"] = arrs["scf"][..., :21].copy()
            arrs["scf_s"] = arrs["scf"][..., 24:60].reshape(n, 2, 2, 12, 3).copy()
            return pcm, arrs
        return pcm

    def decode(self, stream, lookahead=0, taps=False, **kw):
        return self.decode_parsed(Parsed(stream, lookahead=lookahead, **kw), taps=taps)

    def synth_from_xr(self, xr, parsed):
        """BASELINE configs[1]: IMDCT + polyphase only, spectra (after antialias) from the host. -> pcm"""
        xr = np.ascontiguousarray(xr, dtype=np.float32)
        assert xr.shape == (parsed.n_frames, 2, 2, 576)
        pcm = np.zeros((parsed.n_pcm_frames, 1152, parsed.nch), np.int16)
        _check(lib().p3_synth_from_xr(self.h, xr.ctypes.data, C.byref(parsed.c), pcm.ctypes.data), "p3_synth_from_xr")
        return pcm

    # batches staged from raw bytes: the frame hop runs on the device (p3_hop.cu)
    def upload_raw(self, stream, lookahead=0, max_frames=0, warmup=0, iso=False, state=None):
        """-> dict(n_frames, n_pcm_frames, consumed, stop, nch, rounds); the descriptors stay on the device (download_desc)"""
        self._raw = np.ascontiguousarray(stream, dtype=np.uint8)
        o = P3ParseOpts(max_frames, lookahead, 0, warmup, 1, 1 if iso else 0)
        st = state if state is not None else P3ParseState(0, 0, 0, -1, -1)
        info = P3Parsed()
        _check(lib().p3_batch_upload_raw(self.h, self._raw.ctypes.data if len(self._raw) else None, len(self._raw), 0, C.byref(o), C.byref(st), C.byref(info)), "p3_batch_upload_raw")
        self._up = None
        return dict(n_frames=info.n_frames, n_pcm_frames=info.n_pcm_frames, consumed=info.consumed, stop=info.stop, rounds=lib().p3_hop_rounds(self.h),
                    nch=lib().p3_batch_channels(self.h), hop_ms=lib().p3_hop_ms(self.h))

    def download_raw(self, info):
        """PCM of the batch staged by upload_raw() (after run())"""
        pcm = np.zeros((info["n_pcm_frames"], 1152, info["nch"]), np.int16)
        _check(lib().p3_batch_download(self.h, pcm.ctypes.data, None), "p3_batch_download")
        return pcm

    def decode_raw(self, stream, lookahead=0, max_frames=0, warmup=0, iso=False, state=None, taps=False):
        """device hop + decode: -> (pcm [n_pcm_frames,1152,nch], info dict[, taps])"""
        raw = np.ascontiguousarray(stream, dtype=np.uint8)
        o = P3ParseOpts(max_frames, lookahead, 0, warmup, 1, 1 if iso else 0)
        st = state if state is not None else P3ParseState(0, 0, 0, -1, -1)
        info = P3Parsed()
        cap = max_frames if max_frames > 0 else len(raw) // 96 + 16
        pcm = np.zeros(cap * 1152 * 2, np.int16)
        arrs, t = self._taps(cap) if taps else ({}, None)
        _check(lib().p3_decode_raw(self.h, raw.ctypes.data if len(raw) else None, len(raw), C.byref(o), C.byref(st), C.byref(info), pcm.ctypes.data, cap,
                                   C.byref(t) if taps else None), "p3_decode_raw")
        nch = lib().p3_batch_channels(self.h)
        d = dict(n_frames=info.n_frames, n_pcm_frames=info.n_pcm_frames, consumed=info.consumed, stop=info.stop, rounds=lib().p3_hop_rounds(self.h), nch=nch)
        pcm = pcm[:info.n_pcm_frames * 1152 * nch].reshape(info.n_pcm_frames, 1152, nch)
        if taps:
            return pcm, d, {k: a[:info.n_frames] for k, a in arrs.items()}
        return pcm, d

    # device-resident path (bench)
    def upload(self, parsed):
        _check(lib().p3_batch_upload(self.h, parsed.stream.ctypes.data, len(parsed.stream), C.byref(parsed.c)), "p3_batch_upload")
        self._up = parsed

    def run(self):
        _check(lib().p3_batch_run(self.h), "p3_batch_run")

    def sync(self):
        _check(lib().p3_batch_sync(self.h), "p3_batch_sync")

    def download(self):
        p = self._up
        pcm = np.zeros((p.n_pcm_frames, 1152, p.nch), np.int16)
        _check(lib().p3_batch_download(self.h, pcm.ctypes.data, None), "p3_batch_download")
        return pcm

    def download_desc(self, n_frames):
        """the frame / granule-channel descriptors as the kernels see them (after the device side-info parser)"""
        fr = np.zeros(n_frames, FRAME_DT); gc = np.zeros((n_frames, 4, 4), np.uint32)
        _check(lib().p3_batch_download_desc(self.h, fr.ctypes.data, gc.ctypes.data), "p3_batch_download_desc")
        return fr, gc

    def time(self, iters=5):
        tot = C.c_float(); st = (C.c_float * 8)()
        _check(lib().p3_batch_time(self.h, iters, C.byref(tot), st), "p3_batch_time")
        return tot.value, [st[i] for i in range(5)]

    def time_xr(self, iters=5):
        """configs[1]: k_imdct + k_polyphase over the device-resident spectra of the last EXACT run -> (ms, [ms_imdct, ms_polyphase])"""
        tot = C.c_float(); st = (C.c_float * 2)()
        _check(lib().p3_batch_time_xr(self.h, iters, C.byref(tot), st), "p3_batch_time_xr")
        return tot.value, [st[0], st[1]]

    def launch_count(self):
        return lib().p3_kernel_launch_count(self.h)


def dist_unique_id():
    """rank 0: the 256 bytes (two NCCL unique ids) every rank needs for Dist()"""
    buf = (C.c_uint8 * 256)()
    _check(lib().p3_dist_unique_id(buf), "p3_dist_unique_id")
    return bytes(buf)


class Dist:
    """BASELINE configs[4]: the frame-sharded decode of one stream held on rank 0, PCM gathered to rank 0 (p3_dist.cuh; NCCL from C)."""

    def __init__(self, ctx, ids, rank, world):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.h = C.c_void_p()
        b = (C.c_uint8 * 256).from_buffer_copy(ids)
        _check(lib().p3_dist_init(ctx.h, b, rank, world, C.byref(self.h)), "p3_dist_init")

    def sharded_decode(self, stream=None, device_ptr=None, nbytes=0, chunk_frames=0, iso=False):
        """rank 0: `stream` (np.uint8, host) or (device_ptr, nbytes) of a device buffer with 64 bytes of slack; others: nothing.
        -> P3ShardResult fields as a dict; rank 0 then reads the PCM with pcm()"""
        o = P3ParseOpts(0, 0, 0, 0, 1, 1 if iso else 0)
        res = P3ShardResult()
        if device_ptr is not None:
            rc = lib().p3_sharded_decode(self.h, device_ptr, nbytes, 1, C.byref(o), chunk_frames, C.byref(res))
        elif stream is not None:
            self._raw = np.ascontiguousarray(stream, dtype=np.uint8)
            rc = lib().p3_sharded_decode(self.h, self._raw.ctypes.data, len(self._raw), 0, C.byref(o), chunk_frames, C.byref(res))
        else:
            rc = lib().p3_sharded_decode(self.h, None, 0, 0, C.byref(o), chunk_frames, C.byref(res))
        _check(rc, "p3_sharded_decode")
        return {k: getattr(res, k) for k, _ in P3ShardResult._fields_}

    def pcm(self, res):
        """rank 0: the PCM of the whole stream [n_frames_total, 1152, nch]"""
        out = np.zeros((res["n_frames_total"], 1152, res["nch"]), np.int16)
        _check(lib().p3_batch_download(self.ctx.h, out.ctypes.data, None), "p3_batch_download")
        return out

    def gather_transport(self):
        return "cuda-ipc copy engine" if lib().p3_dist_gather_transport(self.h) == 1 else "nccl send/recv"

    def measure_ingest(self, bytes_per_rank, iters=3):
        ms = C.c_float()
        _check(lib().p3_dist_measure_ingest(self.h, bytes_per_rank, iters, C.byref(ms)), "p3_dist_measure_ingest")
        return ms.value

    def close(self):
        if self.h:
            lib().p3_dist_destroy(self.h)
            self.h = C.c_void_p()


class Decoder:
    """The reference's streaming API, one method per function (pdmp3.c:2351-2535)."""

    def __init__(self, options=None):
        L = lib()
        if not hasattr(L, "pdmp3_new"):
            raise P3Error("libpdmp3_b200.so lacks the pdmp3_* streaming API")
        err = C.c_int(0)
        self.h = L.pdmp3_new(options.encode() if options else None, C.byref(err))
        if not self.h:
            raise P3Error("pdmp3_new failed: %s" % L.p3_last_error().decode(errors="replace"))

    def open_feed(self):
        return lib().pdmp3_open_feed(self.h)

    def feed(self, data):
        b = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data, dtype=np.uint8)
        return lib().pdmp3_feed(self.h, b.ctypes.data if len(b) else None, len(b))

    def read(self, outsize):
        out = np.zeros(outsize, np.uint8); done = C.c_size_t(0)
        rc = lib().pdmp3_read(self.h, out.ctypes.data if outsize else None, outsize, C.byref(done))
        return rc, out[:done.value]

    def read_into(self, out):
        done = C.c_size_t(0)
        rc = lib().pdmp3_read(self.h, out.ctypes.data, out.nbytes, C.byref(done))
        return rc, done.value

    def decode(self, data, outsize):
        b = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data, dtype=np.uint8)
        out = np.zeros(max(outsize, 1), np.uint8); done = C.c_size_t(0)
        rc = lib().pdmp3_decode(self.h, b.ctypes.data, len(b), out.ctypes.data if outsize else None, outsize, C.byref(done))
        return rc, out[:done.value]

    def getformat(self):
        rate = C.c_long(); ch = C.c_int(); enc = C.c_int()
        rc = lib().pdmp3_getformat(self.h, C.byref(rate), C.byref(ch), C.byref(enc))
        return rc, rate.value, ch.value, enc.value

    def close(self):
        if getattr(self, "h", None):
            lib().pdmp3_delete(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
