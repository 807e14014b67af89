"""ctypes binding of the product library libvorbis_b200.so (C ABI: include/vorbis_b200.h).

This module is plumbing for tests and bench.py: it only marshals numpy arrays / raw
device pointers into the C entry points.  There is NO CPU fallback: if the CUDA
library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvorbis_b200.so")

PCM_F32_PLANAR = 1
PCM_S16_INTERLEAVED = 2
IWORK_S32 = 0
IWORK_S16 = 1

EXPORTS = [
    "vb200_ctx_create", "vb200_ctx_destroy", "vb200_device_count", "vb200_last_error",
    "vb200_ctx_table", "vb200_launch_count", "vb200_set_profiling", "vb200_phaseA_kernel_ms", "vb200_debug_phase_cycles",
    "vb200_encode_dsp_kernel_ms",
    "vb200_mdct_forward_dev", "vb200_mdct_forward", "vb200_mdct_backward_dev", "vb200_mdct_backward",
    "vb200_apply_window", "vb200_drft_forward",
    "vb200_noisemask", "vb200_tonemask", "vb200_offset_and_mix",
    "vb200_analysis_phaseA_dev", "vb200_analysis_phaseA", "vb200_analysis_phaseA_streams_dev",
    "vb200_analysis_phaseA_pcmstream_dev", "vb200_synthesis_s16_dev",
    "vb200_couple_quantize_normalize_dev", "vb200_couple_quantize_normalize",
    "vb200_synthesis_dev", "vb200_synthesis", "vb200_decouple_dev", "vb200_decouple",
    "vb200_floor1_fit_dev", "vb200_floor1_fit", "vb200_floor1_render_dev", "vb200_floor1_render",
    "vb200_encode_dsp_dev", "vb200_encode_dsp", "vb200_encode_dsp_managed_dev", "vb200_encode_dsp_managed",
    "vb200_envelope_search_dev", "vb200_envelope_search", "vb200_envelope_search_var", "vb200_envelope_apply_marks",
    "vb200_floor1_inverse2_dev", "vb200_floor1_inverse2", "vb200_decode_dsp_dev", "vb200_decode_dsp",
    "vb200_residue_partvals", "vb200_residue_classify_dev", "vb200_residue_classify",
    "vb200_plan_blocks", "vb200_encode_streams_dev", "vb200_encode_streams",
    "vb200_malloc_device", "vb200_free_device", "vb200_memcpy_h2d", "vb200_memcpy_d2h", "vb200_synchronize",
]

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
vp = C.c_void_p


class VB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load libvorbis_b200.so; raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VB200Error("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                         % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.vb200_last_error.restype = C.c_char_p
    L.vb200_launch_count.restype = C.c_uint64
    L.vb200_launch_count.argtypes = [vp]
    L.vb200_set_profiling.argtypes = [vp, C.c_int]
    L.vb200_phaseA_kernel_ms.argtypes = [vp, C.POINTER(C.c_float * 3)]
    L.vb200_encode_dsp_kernel_ms.argtypes = [vp, C.POINTER(C.c_float * 6)]
    L.vb200_debug_phase_cycles.argtypes = [vp, C.POINTER(C.c_ulonglong * 16), C.c_int]
    L.vb200_ctx_create.argtypes = [C.POINTER(abi.Setup), C.c_int, C.POINTER(vp)]
    L.vb200_ctx_destroy.argtypes = [vp]
    L.vb200_ctx_table.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
    L.vb200_mdct_forward_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.vb200_mdct_backward_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.vb200_mdct_forward.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.vb200_mdct_backward.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.vb200_apply_window.argtypes = [vp, C.c_int, C.c_int, vp, vp, f32p]
    L.vb200_drft_forward.argtypes = [vp, C.c_int, C.c_int, f32p]
    L.vb200_noisemask.argtypes = [vp, C.c_int, C.c_int, f32p, f32p]
    L.vb200_tonemask.argtypes = [vp, C.c_int, C.c_int, f32p, f32p, f32p, f32p]
    L.vb200_offset_and_mix.argtypes = [vp, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p]
    L.vb200_analysis_phaseA_dev.argtypes = [vp, C.c_int, C.c_int, C.POINTER(abi.PhaseAIO), vp]
    L.vb200_analysis_phaseA.argtypes = [vp, C.c_int, C.c_int, C.POINTER(abi.PhaseAIO)]
    L.vb200_analysis_phaseA_streams_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(abi.PhaseAIO), vp, vp]
    L.vb200_analysis_phaseA_pcmstream_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int64, C.c_int,
                                                      C.POINTER(abi.PhaseAIO), vp, vp]
    L.vb200_synthesis_s16_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int64, vp]
    L.vb200_couple_quantize_normalize_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.vb200_couple_quantize_normalize.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.vb200_synthesis_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int64, vp]
    L.vb200_synthesis.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int64, vp, vp, C.c_int64]
    L.vb200_decouple_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.vb200_decouple.argtypes = [vp, C.c_int, C.c_int, vp]
    L.vb200_floor1_fit_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.vb200_floor1_fit.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.vb200_floor1_render_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.vb200_floor1_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.vb200_encode_dsp_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(abi.EncodeIO), vp]
    L.vb200_encode_dsp.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(abi.EncodeIO)]
    L.vb200_encode_dsp_managed.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(abi.EncodeIO)]
    L.vb200_encode_dsp_managed_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(abi.EncodeIO), vp]
    L.vb200_residue_partvals.argtypes = [vp, C.c_int]
    L.vb200_residue_classify_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp]
    L.vb200_residue_classify.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int]
    L.vb200_floor1_inverse2_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.vb200_floor1_inverse2.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.vb200_decode_dsp_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int64, vp]
    L.vb200_decode_dsp.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int64, vp, vp, vp, vp, C.c_int, C.c_int64]
    L.vb200_envelope_search_dev.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int64, C.c_int, C.c_int, vp, vp, vp]
    L.vb200_envelope_search.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int64, C.c_int, C.c_int, vp, vp]
    L.vb200_envelope_apply_marks.argtypes = [vp, C.c_int, C.c_int, vp]
    L.vb200_envelope_apply_marks.restype = None
    L.vb200_plan_blocks.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int, vp, vp, C.c_int, vp, vp]
    L.vb200_encode_streams.argtypes = [vp, C.c_int, C.c_int, C.POINTER(abi.StreamsIO)]
    L.vb200_encode_streams_dev.argtypes = [vp, C.c_int, C.c_int, C.POINTER(abi.StreamsIO), vp]
    L.vb200_malloc_device.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.vb200_free_device.argtypes = [vp, vp]
    L.vb200_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.vb200_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.vb200_synchronize.argtypes = [vp]
    _lib = L
    return L


def _ptr(a):
    """numpy array / int (device pointer) / None -> c_void_p value"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class Context:
    """One vb200_ctx: device tables for a (channels, rate, quality) setup on one GPU."""

    def __init__(self, setup, device=0):
        self.L = load()
        self.setup = setup
        self.channels = setup.channels
        self.bs = [setup.blocksize(0), setup.blocksize(1)]
        h = vp()
        self._chk(self.L.vb200_ctx_create(C.byref(setup.c), device, C.byref(h)))
        self.h = h

    def _chk(self, rc):
        if rc != 0:
            raise VB200Error("vb200 error %d: %s" % (rc, (self.L.vb200_last_error() or b"").decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.vb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self):
        return int(self.L.vb200_launch_count(self.h))

    def set_profiling(self, on=True):
        self._chk(self.L.vb200_set_profiling(self.h, 1 if on else 0))

    def phaseA_kernel_ms(self):
        ms = (C.c_float * 3)()
        self._chk(self.L.vb200_phaseA_kernel_ms(self.h, C.byref(ms)))
        return [float(x) for x in ms]

    def encode_dsp_kernel_ms(self):
        ms = (C.c_float * 6)()
        self._chk(self.L.vb200_encode_dsp_kernel_ms(self.h, C.byref(ms)))
        return list(ms)

    def debug_phase_cycles(self, reset=True):
        out = (C.c_ulonglong * 16)()
        self._chk(self.L.vb200_debug_phase_cycles(self.h, C.byref(out), 1 if reset else 0))
        return [int(x) for x in out]

    def table(self, W, which):
        N = self.bs[W]
        out = np.zeros(N // 4, np.int32) if which == 1 else np.zeros(2 * N, np.float32)
        k = self.L.vb200_ctx_table(self.h, W, which, out.ctypes.data, out.size)
        if k <= 0:
            raise VB200Error("vb200_ctx_table failed")
        return out[:k].copy()

    # ---- transforms: host buffers -------------------------------------------
    def mdct_forward(self, W, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.bs[W])
        out = np.empty((x.shape[0], self.bs[W] // 2), np.float32)
        self._chk(self.L.vb200_mdct_forward(self.h, W, x.shape[0], _ptr(x), _ptr(out)))
        return out

    def mdct_backward(self, W, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.bs[W] // 2)
        out = np.empty((x.shape[0], self.bs[W]), np.float32)
        self._chk(self.L.vb200_mdct_backward(self.h, W, x.shape[0], _ptr(x), _ptr(out)))
        return out

    def apply_window(self, W, x, lW=None, nW=None):
        x = np.array(x, np.float32).reshape(-1, self.bs[W])
        lWa = None if lW is None else np.ascontiguousarray(lW, np.int32)
        nWa = None if nW is None else np.ascontiguousarray(nW, np.int32)
        self._chk(self.L.vb200_apply_window(self.h, W, x.shape[0], _ptr(lWa), _ptr(nWa), x))
        return x

    def drft_forward(self, W, x):
        x = np.array(x, np.float32).reshape(-1, self.bs[W])
        self._chk(self.L.vb200_drft_forward(self.h, W, x.shape[0], x))
        return x

    # ---- transforms: device pointers ----------------------------------------
    def mdct_forward_dev(self, W, nvec, d_in, d_out, stream=None):
        self._chk(self.L.vb200_mdct_forward_dev(self.h, W, nvec, _ptr(d_in), _ptr(d_out), _ptr(stream)))

    def mdct_backward_dev(self, W, nvec, d_in, d_out, stream=None):
        self._chk(self.L.vb200_mdct_backward_dev(self.h, W, nvec, _ptr(d_in), _ptr(d_out), _ptr(stream)))

    # ---- psy stages -----------------------------------------------------------
    def noisemask(self, look, logmdct):
        x = np.ascontiguousarray(logmdct, np.float32)
        x = x.reshape(-1, x.shape[-1])
        out = np.empty_like(x)
        self._chk(self.L.vb200_noisemask(self.h, look, x.shape[0], x, out))
        return out

    def tonemask(self, look, logfft, gmax, lmax):
        x = np.ascontiguousarray(logfft, np.float32)
        x = x.reshape(-1, x.shape[-1])
        out = np.empty_like(x)
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(gmax, np.float32), (x.shape[0],)))
        l = np.ascontiguousarray(np.broadcast_to(np.asarray(lmax, np.float32), (x.shape[0],)))
        self._chk(self.L.vb200_tonemask(self.h, look, x.shape[0], x, g, l, out))
        return out

    def offset_and_mix(self, look, sel, noise, tone, mdct, logmdct):
        noise = np.ascontiguousarray(noise, np.float32)
        noise = noise.reshape(-1, noise.shape[-1])
        tone = np.ascontiguousarray(tone, np.float32).reshape(noise.shape)
        mdct = np.array(mdct, np.float32).reshape(noise.shape)
        logmdct = np.ascontiguousarray(logmdct, np.float32).reshape(noise.shape)
        logmask = np.empty_like(noise)
        self._chk(self.L.vb200_offset_and_mix(self.h, look, noise.shape[0], sel, noise, tone, mdct,
                                              logmdct, logmask))
        return logmask, mdct

    # ---- Phase A ---------------------------------------------------------------
    def phaseA(self, W, pcm, desc, taps=False):
        ch, N = self.channels, self.bs[W]
        pcm = np.ascontiguousarray(pcm, np.float32).reshape(-1, ch, N)
        nb = pcm.shape[0]
        desc = np.ascontiguousarray(desc, abi.BLOCKDESC_DTYPE)
        out = {k: np.empty((nb, ch, N // 2), np.float32) for k in ("mdct", "logmdct", "logmask")}
        out["ampmax_out"] = np.empty(nb, np.float32)
        io = abi.PhaseAIO()
        io.pcm, io.desc = pcm.ctypes.data, desc.ctypes.data
        io.mdct, io.logmdct, io.logmask = (out[k].ctypes.data for k in ("mdct", "logmdct", "logmask"))
        io.ampmax_out = out["ampmax_out"].ctypes.data
        if taps:
            for k in ("noise", "tone", "logfft", "mdct_raw"):
                out[k] = np.empty((nb, ch, N // 2), np.float32)
                setattr(io, "tap_" + k, out[k].ctypes.data)
        self._chk(self.L.vb200_analysis_phaseA(self.h, W, nb, C.byref(io)))
        return out

    def phaseA_dev(self, W, nblocks, io, stream=None, streams=None, d_ampmax0=None):
        """io: abi.PhaseAIO holding DEVICE pointers."""
        if streams is None:
            self._chk(self.L.vb200_analysis_phaseA_dev(self.h, W, nblocks, C.byref(io), _ptr(stream)))
        else:
            self._chk(self.L.vb200_analysis_phaseA_streams_dev(self.h, W, streams[0], streams[1], C.byref(io),
                                                               _ptr(d_ampmax0), _ptr(stream)))

    def phaseA_pcmstream_dev(self, W, nstreams, bps, d_pcm, fmt, stream_stride, hop, io, d_ampmax0=None, stream=None):
        self._chk(self.L.vb200_analysis_phaseA_pcmstream_dev(self.h, W, nstreams, bps, _ptr(d_pcm), fmt, stream_stride,
                                                             hop, C.byref(io), _ptr(d_ampmax0), _ptr(stream)))

    def synthesis_s16_dev(self, nstreams, nblk, d_Wseq, d_coef_off, d_coef, d_pcm_off, d_pcm16, pcm_stride, stream=None):
        self._chk(self.L.vb200_synthesis_s16_dev(self.h, nstreams, nblk, _ptr(d_Wseq), _ptr(d_coef_off), _ptr(d_coef),
                                                 _ptr(d_pcm_off), _ptr(d_pcm16), pcm_stride, _ptr(stream)))

    # ---- Phase B ---------------------------------------------------------------
    def couple_quantize_normalize(self, W, blocktype, blobno, mdct, iwork, nonzero):
        mdct = np.ascontiguousarray(mdct, np.float32)
        iwork = np.array(iwork, np.int32)
        nonzero = np.array(nonzero, np.int32)
        self._chk(self.L.vb200_couple_quantize_normalize(self.h, W, blocktype, blobno, mdct.shape[0],
                                                         _ptr(mdct), _ptr(iwork), _ptr(nonzero)))
        return iwork, nonzero

    def couple_quantize_normalize_dev(self, W, blocktype, blobno, nblocks, d_mdct, d_iwork, d_nonzero, stream=None):
        self._chk(self.L.vb200_couple_quantize_normalize_dev(self.h, W, blocktype, blobno, nblocks,
                                                             _ptr(d_mdct), _ptr(d_iwork), _ptr(d_nonzero),
                                                             _ptr(stream)))

    # ---- floor 1 (lib/floor1.c:576 floor1_fit, :765 floor1_encode minus the bit packing) -------
    def floor1_fit(self, W, logmdct, logmask, floor_sel=-1):
        """logmdct, logmask [rows][n] -> (posts [rows][FLOOR1_STRIDE] int32, fit_nonzero [rows])"""
        n = self.bs[W] // 2
        a = np.ascontiguousarray(logmdct, np.float32).reshape(-1, n)
        b = np.ascontiguousarray(logmask, np.float32).reshape(-1, n)
        posts = np.zeros((a.shape[0], abi.FLOOR1_STRIDE), np.int32)
        nz = np.zeros(a.shape[0], np.int32)
        self._chk(self.L.vb200_floor1_fit(self.h, W, floor_sel, a.shape[0], _ptr(a), _ptr(b), _ptr(posts), _ptr(nz)))
        return posts, nz

    def floor1_render(self, W, posts, fit_nonzero, floor_sel=-1):
        """posts from floor1_fit -> (posts as floor1_encode leaves them, ilogmask [rows][n], nonzero)"""
        n = self.bs[W] // 2
        posts = np.array(posts, np.int32).reshape(-1, abi.FLOOR1_STRIDE)
        fz = np.ascontiguousarray(fit_nonzero, np.int32)
        ilog = np.zeros((posts.shape[0], n), np.int32)
        nz = np.zeros(posts.shape[0], np.int32)
        self._chk(self.L.vb200_floor1_render(self.h, W, floor_sel, posts.shape[0], _ptr(posts), _ptr(fz),
                                             _ptr(ilog), _ptr(nz)))
        return posts, ilog, nz

    def floor1_fit_dev(self, W, nrows, d_logmdct, d_logmask, d_posts, d_fit_nonzero, floor_sel=-1, stream=None):
        self._chk(self.L.vb200_floor1_fit_dev(self.h, W, floor_sel, nrows, _ptr(d_logmdct), _ptr(d_logmask),
                                              _ptr(d_posts), _ptr(d_fit_nonzero), _ptr(stream)))

    def floor1_render_dev(self, W, nrows, d_posts, d_fit_nonzero, d_ilogmask, d_nonzero, floor_sel=-1, stream=None):
        self._chk(self.L.vb200_floor1_render_dev(self.h, W, floor_sel, nrows, _ptr(d_posts), _ptr(d_fit_nonzero),
                                                 _ptr(d_ilogmask), _ptr(d_nonzero), _ptr(stream)))

    # ---- whole streams: envelope marks -> block plan -> both block sizes, ampmax chain across sizes ----
    def plan_blocks(self, mark, nsteps, pcm_len, eof=None, max_blocks=None):
        """mark [streams][stride] int32 (timeline steps), pcm_len/eof [streams] int64 -> (plan, nblocks)"""
        mark = np.ascontiguousarray(mark, np.int32)
        ns, stride = mark.shape
        pcm_len = np.ascontiguousarray(pcm_len, np.int64)
        eofp = None if eof is None else np.ascontiguousarray(eof, np.int64)
        if max_blocks is None:
            max_blocks = int(pcm_len.max()) // (self.bs[0] // 2) + 8
        plan = np.zeros((ns, max_blocks), abi.STREAM_BLOCK_DTYPE)
        nb = np.zeros(ns, np.int32)
        self._chk(self.L.vb200_plan_blocks(self.h, ns, mark.ctypes.data, stride, int(nsteps), pcm_len.ctypes.data,
                                        None if eofp is None else eofp.ctypes.data, max_blocks, plan.ctypes.data,
                                        nb.ctypes.data))
        return plan, nb

    def encode_streams(self, pcm, pcm_len, eof=None, fmt=PCM_F32_PLANAR, max_blocks=None, cap=None, blobno=7):
        """pcm: timeline buffers, PCM_F32_PLANAR [streams][ch][stride] float32 or PCM_S16_INTERLEAVED
        [streams][stride][ch] int16.  Returns plan, nblocks and per block size W the batch outputs."""
        ch = self.channels
        if fmt == PCM_F32_PLANAR:
            pcm = np.ascontiguousarray(pcm, np.float32)
            ns, stride = pcm.shape[0], pcm.shape[2]
            assert pcm.shape[1] == ch
        else:
            pcm = np.ascontiguousarray(pcm, np.int16)
            ns, stride = pcm.shape[0], pcm.shape[1]
            assert pcm.shape[2] == ch
        pcm_len = np.ascontiguousarray(pcm_len, np.int64)
        eofp = None if eof is None else np.ascontiguousarray(eof, np.int64)
        if max_blocks is None:
            max_blocks = stride // (self.bs[0] // 2) + 8
        if cap is None:
            cap = [ns * max_blocks, ns * (stride // (self.bs[1] // 2) + 8)]
        io = abi.StreamsIO()
        io.pcm, io.pcm_fmt, io.max_blocks, io.stream_stride = pcm.ctypes.data, fmt, max_blocks, stride
        io.pcm_len = pcm_len.ctypes.data
        io.eof = None if eofp is None else eofp.ctypes.data
        plan = np.zeros((ns, max_blocks), abi.STREAM_BLOCK_DTYPE)
        nb = np.zeros(ns, np.int32)
        io.plan, io.nblocks = plan.ctypes.data, nb.ctypes.data
        out = {}
        for w in range(2):
            n = self.bs[w] // 2
            io.cap[w] = int(cap[w])
            out[w] = {"posts": np.zeros((cap[w], ch, abi.FLOOR1_STRIDE), np.int32), "nonzero": np.zeros((cap[w], ch), np.int32),
                      "iwork": np.zeros((cap[w], ch, n), np.int32), "ampmax_out": np.zeros(cap[w], np.float32)}
            io.posts[w], io.nonzero[w] = out[w]["posts"].ctypes.data, out[w]["nonzero"].ctypes.data
            io.iwork[w], io.ampmax_out[w] = out[w]["iwork"].ctypes.data, out[w]["ampmax_out"].ctypes.data
        self._chk(self.L.vb200_encode_streams(self.h, ns, blobno, C.byref(io)))
        for w in range(2):
            out[w] = {k: v[:io.count[w]] for k, v in out[w].items()}
        return {"plan": plan, "nblocks": nb, "count": [io.count[0], io.count[1]], 0: out[0], 1: out[1]}

    # ---- whole per-block encode DSP (Phase A -> floor1 -> Phase B) in one call ---------------
    def encode_dsp(self, W, pcm, desc, nstreams=None, fmt=0, hop=0, ampmax0=None, independent=None, blobno=7,
                   floats=False, iwork_s16=False, classes=False):
        """Host buffers.  fmt 0: pcm [nblocks][ch][N] float; PCM_F32_PLANAR: [streams][ch][stride] float;
        PCM_S16_INTERLEAVED: [streams][stride][ch] int16.  nstreams None = every block its own stream.
        independent None = True when the blocks are not grouped in streams."""
        ch, N = self.channels, self.bs[W]
        n = N // 2
        desc = np.ascontiguousarray(desc, abi.BLOCKDESC_DTYPE)
        nb = desc.shape[0]
        if nstreams is None:
            nstreams = nb
            if independent is None:
                independent = True
        bps = nb // nstreams
        assert bps * nstreams == nb
        io = abi.EncodeIO()
        if fmt == 0:
            pcm = np.ascontiguousarray(pcm, np.float32).reshape(nb, ch, N)
        elif fmt == PCM_F32_PLANAR:
            pcm = np.ascontiguousarray(pcm, np.float32)
            assert pcm.shape[:2] == (nstreams, ch)
            io.stream_stride = pcm.shape[2]
        else:
            pcm = np.ascontiguousarray(pcm, np.int16)
            assert pcm.shape[0] == nstreams and pcm.shape[2] == ch
            io.stream_stride = pcm.shape[1]
        io.pcm, io.pcm_fmt, io.hop = pcm.ctypes.data, fmt, hop
        io.desc = desc.ctypes.data
        io.independent = 1 if independent else 0
        if ampmax0 is not None:
            ampmax0 = np.ascontiguousarray(ampmax0, np.float32)
            io.ampmax0 = ampmax0.ctypes.data
        out = {"posts": np.zeros((nb, ch, abi.FLOOR1_STRIDE), np.int32), "nonzero": np.zeros((nb, ch), np.int32),
               "iwork": np.zeros((nb, ch, n), np.int16 if iwork_s16 else np.int32),
               "ampmax_out": np.zeros(nb, np.float32)}
        if iwork_s16:
            io.iwork_fmt = IWORK_S16
            out["overflow"] = np.full(nb, -1, np.int32)
        if classes:
            io.class_stride = self.residue_partvals(W)
            out["classes"] = np.full((nb, ch, int(io.class_stride)), -1, np.int32)
        if floats:
            for k in ("mdct", "logmdct", "logmask"):
                out[k] = np.zeros((nb, ch, n), np.float32)
        for k, v in out.items():
            setattr(io, k, v.ctypes.data)
        self._chk(self.L.vb200_encode_dsp(self.h, W, nstreams, bps, blobno, C.byref(io)))
        return out

    def encode_dsp_managed(self, W, pcm, desc, nstreams=None, fmt=0, hop=0, ampmax0=None, independent=None):
        """Bitrate-managed mode (vb200_encode_dsp_managed): the 15 curves of every block.  Inputs as encode_dsp;
        returns posts [15][nb][ch][FLOOR1_STRIDE], nonzero [15][nb][ch], iwork [15][nb][ch][n], ampmax_out [nb]."""
        ch, N = self.channels, self.bs[W]
        n = N // 2
        desc = np.ascontiguousarray(desc, abi.BLOCKDESC_DTYPE)
        nb = desc.shape[0]
        if nstreams is None:
            nstreams = nb
            if independent is None:
                independent = True
        bps = nb // nstreams
        assert bps * nstreams == nb
        io = abi.EncodeIO()
        if fmt == 0:
            pcm = np.ascontiguousarray(pcm, np.float32).reshape(nb, ch, N)
        elif fmt == PCM_F32_PLANAR:
            pcm = np.ascontiguousarray(pcm, np.float32)
            assert pcm.shape[:2] == (nstreams, ch)
            io.stream_stride = pcm.shape[2]
        else:
            pcm = np.ascontiguousarray(pcm, np.int16)
            assert pcm.shape[0] == nstreams and pcm.shape[2] == ch
            io.stream_stride = pcm.shape[1]
        io.pcm, io.pcm_fmt, io.hop = pcm.ctypes.data, fmt, hop
        io.desc = desc.ctypes.data
        io.independent = 1 if independent else 0
        if ampmax0 is not None:
            ampmax0 = np.ascontiguousarray(ampmax0, np.float32)
            io.ampmax0 = ampmax0.ctypes.data
        NB = abi.PACKETBLOBS
        out = {"posts": np.full((NB, nb, ch, abi.FLOOR1_STRIDE), -1, np.int32), "nonzero": np.full((NB, nb, ch), -1, np.int32),
               "iwork": np.full((NB, nb, ch, n), -1, np.int32), "ampmax_out": np.zeros(nb, np.float32)}
        for k, v in out.items():
            setattr(io, k, v.ctypes.data)
        self._chk(self.L.vb200_encode_dsp_managed(self.h, W, nstreams, bps, C.byref(io)))
        return out

    def encode_dsp_dev(self, W, nstreams, bps, io, blobno=7, stream=None):
        """io: abi.EncodeIO holding DEVICE pointers."""
        self._chk(self.L.vb200_encode_dsp_dev(self.h, W, nstreams, bps, blobno, C.byref(io), _ptr(stream)))

    # ---- residue partition classification (lib/res0.c:412-532) --------------------------------
    def residue_partvals(self, W):
        return int(self.L.vb200_residue_partvals(self.h, W))

    def residue_classify(self, W, iwork, nonzero, stride=None):
        ch, n = self.channels, self.bs[W] // 2
        iwork = np.ascontiguousarray(iwork, np.int32).reshape(-1, ch, n)
        nonzero = np.ascontiguousarray(nonzero, np.int32).reshape(-1, ch)
        stride = self.residue_partvals(W) if stride is None else stride
        classes = np.full((iwork.shape[0], ch, stride), -1, np.int32)
        self._chk(self.L.vb200_residue_classify(self.h, W, iwork.shape[0], _ptr(iwork), _ptr(nonzero), _ptr(classes),
                                                stride))
        return classes

    def residue_classify_dev(self, W, nblocks, d_iwork, d_nonzero, d_classes, stride, stream=None):
        self._chk(self.L.vb200_residue_classify_dev(self.h, W, nblocks, _ptr(d_iwork), _ptr(d_nonzero), _ptr(d_classes),
                                                    stride, _ptr(stream)))

    # ---- decode: floor multiply and the whole decode DSP in one call ----------------------------
    def floor1_inverse2(self, W, posts, present, data, floor_sel=-1):
        """floor1_inverse2 (lib/floor1.c:1041): rows [block][channel] of n floats, multiplied in place"""
        n = self.bs[W] // 2
        posts = np.ascontiguousarray(posts, np.int32).reshape(-1, abi.FLOOR1_STRIDE)
        present = np.ascontiguousarray(present, np.int32).reshape(-1)
        data = np.array(data, np.float32).reshape(posts.shape[0], n)
        self._chk(self.L.vb200_floor1_inverse2(self.h, W, floor_sel, posts.shape[0], _ptr(posts), _ptr(present),
                                               _ptr(data)))
        return data

    def decode_dsp(self, Wseq, coef_off, res, posts, present, pcm_off, pcm_stride, s16=False):
        """de-couple + floor multiply + IMDCT + overlap-add in one call; layout as synthesis()"""
        Wseq = np.ascontiguousarray(Wseq, np.int32)
        ns, nblk = Wseq.shape
        res = np.array(res, np.float32)
        posts = np.ascontiguousarray(posts, np.int32)
        present = np.ascontiguousarray(present, np.int32)
        coef_off = np.ascontiguousarray(coef_off, np.int64)
        pcm_off = np.ascontiguousarray(pcm_off, np.int64)
        pcm = (np.zeros((ns, pcm_stride, self.channels), np.int16) if s16
               else np.zeros((ns, self.channels, pcm_stride), np.float32))
        self._chk(self.L.vb200_decode_dsp(self.h, ns, nblk, _ptr(Wseq), _ptr(coef_off), _ptr(res), res.size,
                                          _ptr(posts), _ptr(present), _ptr(pcm_off), _ptr(pcm), 1 if s16 else 0,
                                          pcm_stride))
        return pcm

    # ---- envelope / block-switch detector (lib/envelope.c) --------------------------------------
    def envelope_search(self, pcm, first_step, nsteps, state=None, fmt=PCM_F32_PLANAR):
        """Host buffers.  pcm: float [streams][ch][stride] (PCM_F32_PLANAR) or int16 [streams][stride][ch].
        Returns (ret uint8 [streams][nsteps], state int32 [streams][ve_state_words])."""
        ch = self.channels
        if fmt == PCM_F32_PLANAR:
            pcm = np.ascontiguousarray(pcm, np.float32)
            ns, stride = pcm.shape[0], pcm.shape[2]
            assert pcm.shape[1] == ch
        else:
            pcm = np.ascontiguousarray(pcm, np.int16)
            ns, stride = pcm.shape[0], pcm.shape[1]
            assert pcm.shape[2] == ch
        state = (np.zeros((ns, abi.ve_state_words(ch)), np.int32) if state is None
                 else np.array(state, np.int32).reshape(ns, abi.ve_state_words(ch)))
        ret = np.zeros((ns, nsteps), np.uint8)
        self._chk(self.L.vb200_envelope_search(self.h, ns, _ptr(pcm), fmt, stride, first_step, nsteps,
                                               _ptr(state), _ptr(ret)))
        return ret, state

    def envelope_search_dev(self, nstreams, d_pcm, fmt, stride, first_step, nsteps, d_state, d_ret, stream=None):
        self._chk(self.L.vb200_envelope_search_dev(self.h, nstreams, _ptr(d_pcm), fmt, stride, first_step, nsteps,
                                                   _ptr(d_state), _ptr(d_ret), _ptr(stream)))

    def envelope_marks(self, ret, first_step=0, mark=None):
        """replay lib/envelope.c:254-264 for one stream's trigger bits (plain C helper, no GPU)"""
        ret = np.ascontiguousarray(ret, np.uint8)
        if mark is None:
            mark = np.zeros(first_step + len(ret) + 2, np.int32)
        self.L.vb200_envelope_apply_marks(_ptr(ret), first_step, len(ret), _ptr(mark))
        return mark

    # ---- decode ------------------------------------------------------------------
    def decouple(self, W, res):
        res = np.array(res, np.float32)
        self._chk(self.L.vb200_decouple(self.h, W, res.shape[0], _ptr(res)))
        return res

    def synthesis(self, Wseq, coef_off, coef, pcm_off, pcm_stride):
        Wseq = np.ascontiguousarray(Wseq, np.int32)
        ns, nblk = Wseq.shape
        coef_off = np.ascontiguousarray(coef_off, np.int64)
        pcm_off = np.ascontiguousarray(pcm_off, np.int64)
        coef = np.ascontiguousarray(coef, np.float32)
        pcm = np.zeros((ns, self.channels, pcm_stride), np.float32)
        self._chk(self.L.vb200_synthesis(self.h, ns, nblk, _ptr(Wseq), _ptr(coef_off), _ptr(coef), coef.size,
                                         _ptr(pcm_off), _ptr(pcm), pcm_stride))
        return pcm

    def synthesis_dev(self, nstreams, nblk, d_Wseq, d_coef_off, d_coef, d_pcm_off, d_pcm, pcm_stride, stream=None):
        self._chk(self.L.vb200_synthesis_dev(self.h, nstreams, nblk, _ptr(d_Wseq), _ptr(d_coef_off), _ptr(d_coef),
                                             _ptr(d_pcm_off), _ptr(d_pcm), pcm_stride, _ptr(stream)))


def synthesis_layout(Wseq, bs, channels):
    """Offsets for vb200_synthesis: Wseq [nstreams][nblk] -> (coef_off, pcm_off, coef_len, pcm_len)
    with every stream's spectra packed back to back (block-major, channel-minor)."""
    Wseq = np.asarray(Wseq, np.int32)
    ns, nblk = Wseq.shape
    N = np.where(Wseq == 1, bs[1], bs[0]).astype(np.int64)
    per_block = channels * (N // 2)
    coef_off = np.zeros((ns, nblk), np.int64)
    flat = per_block.reshape(-1)
    coef_off.reshape(-1)[1:] = np.cumsum(flat)[:-1]
    fin = np.zeros((ns, nblk), np.int64)
    fin[:, 1:] = N[:, :-1] // 4 + N[:, 1:] // 4
    pcm_off = np.cumsum(fin, axis=1) - fin      # finished samples before block k's contribution
    # block k's finished samples start where block k-1's ended
    pcm_off = np.concatenate([np.zeros((ns, 1), np.int64), np.cumsum(fin, axis=1)[:, :-1]], axis=1)
    pcm_len = int(np.cumsum(fin, axis=1)[:, -1].max())
    return coef_off, pcm_off, int(flat.sum()), pcm_len
