"""Python mirror of the reference's C++ interface for the hot path, on top of the C ABI.

Reference interface being mirrored (file:line of /root/reference):
  class MF              motion_framework.h:9-54   ctor (:12), calcMotionBlockMatching (:13), public ints (:16-19)
  class PyramidLevel    pyramid_level.h:7-16
  class BlockPosition   block_position.h:4-9
  class Flow            rw_flow.h:9-38            ReadFlowFile, WriteFlowFile, CalculateMSE
Error behaviour: where the reference prints and exit(1)s, these raise BbmeError carrying the same text.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import BbmeOptions, BbmeShape, BbmeStats


class BbmeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"[bbme status {status}] {message}")
        self.status = status


def _int_array(values):
    return (C.c_int * len(values))(*[int(v) for v in values])


def _check(lib, ctx, rc, what):
    if rc != 0:
        msg = lib.bbme_last_error(ctx)
        msg = msg.decode() if msg else ""
        if not msg:
            msg = lib.bbme_status_string(rc).decode()
        raise BbmeError(rc, f"{what}: {msg}")


def _shape_dict(sh):
    L = sh.num_levels
    return {
        "width": sh.width, "height": sh.height,
        "padded_width": sh.padded_width, "padded_height": sh.padded_height,
        "padding_x": sh.padding_x, "padding_y": sh.padding_y, "num_levels": L,
        "level_width": list(sh.level_width[:L]), "level_height": list(sh.level_height[:L]),
        "block_size": list(sh.block_size[:L]), "search_size": list(sh.search_size[:L]),
    }


def plan_shape(width, height, search_size, block_size, num_levels=None):
    """Padding search of MF::MF (motion_framework.cpp:15-54); needs no GPU."""
    lib = _lib.load()
    L = len(block_size) if num_levels is None else int(num_levels)
    sh = BbmeShape()
    rc = lib.bbme_plan_shape(int(width), int(height), L, _int_array(search_size[:L]), _int_array(block_size[:L]), C.byref(sh))
    if rc != 0:
        raise BbmeError(rc, lib.bbme_status_string(rc).decode())
    return _shape_dict(sh)


def search_geometry(level_width, level_height, block_size, search_size, allow_copies=True):
    """How the search planner would run one pyramid level (kernel family, ring depth, shared memory); needs no GPU."""
    lib = _lib.load()
    g = _lib.BbmeSearchGeometry()
    rc = lib.bbme_debug_search_geometry(int(level_width), int(level_height), int(block_size), int(search_size), int(bool(allow_copies)), C.byref(g))
    if rc != 0:
        raise BbmeError(rc, lib.bbme_status_string(rc).decode())
    return {k: getattr(g, k) for k, _ in _lib.BbmeSearchGeometry._fields_}


def div_magic(divisor):
    """(magic, shift) of the search kernel's multiply-high division: x // d == ((x * magic) >> 32) >> shift for 0 <= x < 2**31."""
    lib = _lib.load()
    m, s = C.c_uint(), C.c_uint()
    rc = lib.bbme_debug_div_magic(int(divisor), C.byref(m), C.byref(s))
    if rc != 0:
        raise BbmeError(rc, lib.bbme_status_string(rc).decode())
    return m.value, s.value


class PyramidLevel:
    """pyramid_level.h:7-16 -- per-level state; images/flow are fetched from the device on demand."""

    def __init__(self, block_size, search_size, lam, width, height):
        self.block_size = block_size
        self.search_size = search_size
        self.lambda_ = lam  # `lambda` is a Python keyword
        self.width = width
        self.height = height
        self.image1 = None
        self.image2 = None
        self.level_flow = None


class BlockPosition:
    """block_position.h:4-9"""

    def __init__(self, pos_x=0, pos_y=0):
        self.pos_x = pos_x
        self.pos_y = pos_y


class Estimator:
    """A planned context: one GPU, fixed geometry, batched estimation (bbme_plan / bbme_estimate*)."""

    def __init__(self, width, height, search_size, block_size, num_levels=None, sweeps=2, device=0, chunk_pairs=1,
                 slots=1, search_kernel=0, collect_stats=False, keep_search_mv=False, search_variant=0):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        rc = self._lib.bbme_create(C.byref(self._ctx), int(device))
        if rc != 0:
            msg = self._lib.bbme_last_error(None)
            raise BbmeError(rc, "bbme_create: " + (msg.decode() if msg else ""))
        L = len(block_size) if num_levels is None else int(num_levels)
        opt = BbmeOptions()
        self._lib.bbme_default_options(C.byref(opt))
        opt.sweeps = int(sweeps)
        opt.chunk_pairs = int(chunk_pairs)
        opt.slots = int(slots)
        opt.search_kernel = int(search_kernel)
        opt.collect_stats = int(bool(collect_stats))
        opt.keep_search_mv = int(bool(keep_search_mv))
        opt.search_variant = int(search_variant)
        sh = BbmeShape()
        try:
            _check(self._lib, self._ctx,
                   self._lib.bbme_plan(self._ctx, int(width), int(height), L, _int_array(search_size[:L]),
                                       _int_array(block_size[:L]), C.byref(opt), C.byref(sh)), "bbme_plan")
        except Exception:
            self.close()
            raise
        self._sh = sh
        self.shape = _shape_dict(sh)
        self.chunk_pairs = int(chunk_pairs)
        self.device = int(device)

    # -- lifetime
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.bbme_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- host-buffer API (H2D + pipeline + D2H inside the call)
    def flow_shape(self):
        return (self.shape["padded_height"], self.shape["padded_width"], 2)

    def estimate(self, im1, im2, out=None):
        return self.estimate_batch([im1], [im2], None if out is None else [out])[0]

    def estimate_batch(self, im1_list, im2_list, out_list=None):
        n = len(im1_list)
        h, w = self.shape["height"], self.shape["width"]
        pitch = None
        keep = []
        p1 = (C.c_void_p * n)()
        p2 = (C.c_void_p * n)()
        po = (C.c_void_p * n)()
        if out_list is None:
            out_list = [np.empty(self.flow_shape(), np.float32) for _ in range(n)]
        for i in range(n):
            a, b, o = im1_list[i], im2_list[i], out_list[i]
            for img in (a, b):
                if img.dtype != np.uint8 or img.ndim != 2 or img.shape != (h, w) or img.strides[1] != 1:
                    raise BbmeError(-1, f"frames must be uint8 {h}x{w} with unit column stride")
            if pitch is None:
                pitch = a.strides[0]
            if a.strides[0] != pitch or b.strides[0] != pitch:
                raise BbmeError(-1, "all frames of a batch must share one row pitch")
            if o.dtype != np.float32 or o.shape != self.flow_shape() or not o.flags["C_CONTIGUOUS"]:
                raise BbmeError(-1, "flow buffers must be C-contiguous float32 (padded_h, padded_w, 2)")
            keep += [a, b, o]
            p1[i], p2[i], po[i] = a.ctypes.data, b.ctypes.data, o.ctypes.data
        _check(self._lib, self._ctx, self._lib.bbme_estimate_batch(self._ctx, n, p1, p2, pitch, po), "bbme_estimate_batch")
        return out_list

    # -- video sequences: n frames -> n - 1 pairs (t, t + 1), every frame's pyramid built once
    def estimate_sequence(self, frames):
        n = len(frames)
        if n < 2:
            raise BbmeError(-1, "a sequence needs at least two frames")
        h, w = self.shape["height"], self.shape["width"]
        PA = C.c_void_p * n
        pf, po = PA(), PA()
        outs = [np.empty(self.flow_shape(), np.float32) for _ in range(n - 1)]
        keep = []
        for i in range(n):
            a = np.ascontiguousarray(frames[i])
            if a.dtype != np.uint8 or a.shape != (h, w):
                raise BbmeError(-1, f"frames must be uint8 {h}x{w}")
            keep.append(a)
            pf[i] = a.ctypes.data
            po[i] = outs[i].ctypes.data if i < n - 1 else None
        _check(self._lib, self._ctx, self._lib.bbme_estimate_sequence(self._ctx, n, pf, w, po), "bbme_estimate_sequence")
        return outs

    def estimate_sequence_device(self, n_frames, d_frames, pitch, plane, d_flow, flow_plane):
        _check(self._lib, self._ctx,
               self._lib.bbme_estimate_sequence_device(self._ctx, int(n_frames), C.c_void_p(d_frames), int(pitch), int(plane),
                                                       C.c_void_p(d_flow), int(flow_plane)), "bbme_estimate_sequence_device")

    # -- main()'s quarter-pel wrapper on the device (main_class.cpp:32-33, 58-70)
    def estimate_upsampled(self, im1_list, im2_list, factor=4):
        """Frames of (height / factor) x (width / factor) -- width, height as planned -- are up-sampled with
        cv::resize(INTER_LINEAR) arithmetic, run through the path, and come back as main()'s `subpix_MVs`: padding
        stripped, every factor-th pixel, vectors / factor.  Returns a list of (h, w, 2) float32 arrays."""
        single = isinstance(im1_list, np.ndarray) and im1_list.ndim == 2
        if single:
            im1_list, im2_list = [im1_list], [im2_list]
        n = len(im1_list)
        if n == 0 or n != len(im2_list):
            raise BbmeError(-1, "estimate_upsampled needs equally many (and at least one) first and second frames")
        factor = int(factor)
        if factor < 2 or self.shape["width"] % factor or self.shape["height"] % factor:
            raise BbmeError(-1, f"factor {factor} does not divide the planned size")
        h, w = self.shape["height"] // factor, self.shape["width"] // factor
        PA = C.c_void_p * n
        p1, p2, po = PA(), PA(), PA()
        outs = [np.empty((h, w, 2), np.float32) for _ in range(n)]
        keep = []
        for i in range(n):
            a = np.ascontiguousarray(im1_list[i])
            b = np.ascontiguousarray(im2_list[i])
            for img in (a, b):
                if img.dtype != np.uint8 or img.shape != (h, w):
                    raise BbmeError(-1, f"frames must be uint8 {h}x{w} (planned size / factor)")
            keep += [a, b]
            p1[i], p2[i], po[i] = a.ctypes.data, b.ctypes.data, outs[i].ctypes.data
        _check(self._lib, self._ctx, self._lib.bbme_estimate_upsampled(self._ctx, n, factor, p1, p2, w, po),
               "bbme_estimate_upsampled")
        return outs[0] if single else outs

    def estimate_upsampled_device(self, n, factor, d_im1, d_im2, pitch, plane, d_flow, flow_plane):
        _check(self._lib, self._ctx,
               self._lib.bbme_estimate_upsampled_device(self._ctx, int(n), int(factor), C.c_void_p(d_im1), C.c_void_p(d_im2),
                                                        int(pitch), int(plane), C.c_void_p(d_flow), int(flow_plane)),
               "bbme_estimate_upsampled_device")

    # -- device-resident API (raw device pointers, e.g. torch tensors' data_ptr())
    def estimate_device(self, n, d_im1, d_im2, pitch, plane, d_flow, flow_plane):
        _check(self._lib, self._ctx,
               self._lib.bbme_estimate_device(self._ctx, int(n), C.c_void_p(d_im1), C.c_void_p(d_im2), int(pitch),
                                              int(plane), C.c_void_p(d_flow), int(flow_plane)), "bbme_estimate_device")

    def estimate_device_compact(self, n, d_im1, d_im2, pitch, plane, d_mv, mv_plane):
        _check(self._lib, self._ctx,
               self._lib.bbme_estimate_device_compact(self._ctx, int(n), C.c_void_p(d_im1), C.c_void_p(d_im2), int(pitch),
                                                      int(plane), C.c_void_p(d_mv), int(mv_plane)),
               "bbme_estimate_device_compact")

    def estimate_device_both(self, n, d_im1, d_im2, pitch, plane, d_flow, flow_plane, d_mv, mv_plane):
        _check(self._lib, self._ctx,
               self._lib.bbme_estimate_device_both(self._ctx, int(n), C.c_void_p(d_im1), C.c_void_p(d_im2), int(pitch),
                                                   int(plane), C.c_void_p(d_flow), int(flow_plane), C.c_void_p(d_mv),
                                                   int(mv_plane)), "bbme_estimate_device_both")

    def set_streams(self, handles):
        """Run slot i on the caller's CUDA stream handles[i] (e.g. torch.cuda.Stream().cuda_stream)."""
        arr = (C.c_void_p * len(handles))(*[C.c_void_p(int(h)) for h in handles])
        _check(self._lib, self._ctx, self._lib.bbme_set_streams(self._ctx, len(handles), arr), "bbme_set_streams")

    def measure_int_peak(self):
        """(|a-b| per second, SM MHz) of a live VABSDIFF4 issue-rate micro-benchmark on this device."""
        v, f = C.c_double(0), C.c_double(0)
        _check(self._lib, self._ctx, self._lib.bbme_measure_int_peak(self._ctx, C.byref(v), C.byref(f)), "bbme_measure_int_peak")
        return v.value, f.value

    def measure_host_link(self, nbytes=256 << 20):
        """Pinned H2D / D2H bandwidth and the worker threads' host write bandwidth (GB/s): the host-buffer path's ceilings."""
        hl = _lib.BbmeHostLink()
        _check(self._lib, self._ctx, self._lib.bbme_measure_host_link(self._ctx, int(nbytes), C.byref(hl)), "bbme_measure_host_link")
        return {k: getattr(hl, k) for k, _ in _lib.BbmeHostLink._fields_}

    def sync(self):
        _check(self._lib, self._ctx, self._lib.bbme_sync(self._ctx), "bbme_sync")

    def stats(self):
        st = BbmeStats()
        _check(self._lib, self._ctx, self._lib.bbme_get_stats(self._ctx, C.byref(st)), "bbme_get_stats")
        return {k: getattr(st, k) for k, _ in BbmeStats._fields_}

    # -- state of the last call (per-stage parity tests)
    def level_image(self, level, frame, pair=0):
        out = np.empty((self.shape["level_height"][level], self.shape["level_width"][level]), np.uint8)
        _check(self._lib, self._ctx, self._lib.bbme_debug_level_image(self._ctx, pair, frame, level, out.ctypes.data),
               "bbme_debug_level_image")
        return out

    def level_mv(self, level, which=0, pair=0):
        g = 2 if which == 0 else self.shape["block_size"][level]
        out = np.empty((self.shape["level_height"][level] // g, self.shape["level_width"][level] // g, 2), np.int16)
        _check(self._lib, self._ctx, self._lib.bbme_debug_level_mv(self._ctx, pair, level, which, out.ctypes.data),
               "bbme_debug_level_mv")
        return out

    # -- single stages on host arrays
    def stage_pyrdown(self, src):
        src = np.ascontiguousarray(src, np.uint8)
        h, w = src.shape
        dst = np.empty((h // 2, w // 2), np.uint8)
        _check(self._lib, self._ctx, self._lib.bbme_stage_pyrdown(self._ctx, src.ctypes.data, w, h, dst.ctypes.data), "stage_pyrdown")
        return dst

    def stage_resize(self, src, factor):
        src = np.ascontiguousarray(src, np.uint8)
        h, w = src.shape
        dst = np.empty((h * factor, w * factor), np.uint8)
        _check(self._lib, self._ctx, self._lib.bbme_stage_resize(self._ctx, src.ctypes.data, w, h, int(factor), dst.ctypes.data),
               "bbme_stage_resize")
        return dst

    def stage_search(self, im1, im2, block_size, search_size, pred=None, kernel=0):
        im1 = np.ascontiguousarray(im1, np.uint8)
        im2 = np.ascontiguousarray(im2, np.uint8)
        h, w = im1.shape
        mv = np.zeros((h // block_size, w // block_size, 2), np.int16) if pred is None else np.ascontiguousarray(pred, np.int16).copy()
        st = BbmeStats()
        _check(self._lib, self._ctx,
               self._lib.bbme_stage_search(self._ctx, im1.ctypes.data, im2.ctypes.data, w, h, block_size, search_size,
                                           mv.ctypes.data, kernel, C.byref(st)), "stage_search")
        return mv, {k: getattr(st, k) for k, _ in BbmeStats._fields_}

    def stage_search_raster(self, im1, im2, block_size, search_size, pred=None):
        """MF::calcLevelBM with MF::find_min_block (motion_framework.cpp:246-294) instead of the spiral search."""
        im1 = np.ascontiguousarray(im1, np.uint8)
        im2 = np.ascontiguousarray(im2, np.uint8)
        h, w = im1.shape
        mv = np.zeros((h // block_size, w // block_size, 2), np.int16) if pred is None else np.ascontiguousarray(pred, np.int16).copy()
        _check(self._lib, self._ctx,
               self._lib.bbme_stage_search_raster(self._ctx, im1.ctypes.data, im2.ctypes.data, w, h, block_size, search_size,
                                                  mv.ctypes.data), "stage_search_raster")
        return mv

    def stage_compensate(self, im2, block_size, mv):
        """MF::draw_MVimage (motion_framework.cpp:887-905): the motion-compensated frame from image 2 and a block-granular field."""
        im2 = np.ascontiguousarray(im2, np.uint8)
        h, w = im2.shape
        mv = np.ascontiguousarray(mv, np.int16)
        out = np.empty((h, w), np.uint8)
        _check(self._lib, self._ctx,
               self._lib.bbme_stage_compensate(self._ctx, im2.ctypes.data, w, h, block_size, mv.ctypes.data, out.ctypes.data),
               "stage_compensate")
        return out

    def stage_regularize(self, im1, im2, block_size, lam, mult, mv):
        im1 = np.ascontiguousarray(im1, np.uint8)
        im2 = np.ascontiguousarray(im2, np.uint8)
        h, w = im1.shape
        mv = np.ascontiguousarray(mv, np.int16).copy()
        rounds = C.c_uint32(0)
        _check(self._lib, self._ctx,
               self._lib.bbme_stage_regularize(self._ctx, im1.ctypes.data, im2.ctypes.data, w, h, block_size, float(lam),
                                               int(mult), mv.ctypes.data, C.byref(rounds)), "stage_regularize")
        return mv, rounds.value

    def stage_divide(self, mv):
        mv = np.ascontiguousarray(mv, np.int16)
        gh, gw, _ = mv.shape
        out = np.empty((2 * gh, 2 * gw, 2), np.int16)
        _check(self._lib, self._ctx, self._lib.bbme_stage_divide(self._ctx, mv.ctypes.data, gw, gh, out.ctypes.data), "stage_divide")
        return out

    def stage_copy_mvs(self, coarse_mv2, coarse_block_size, fine_block_size):
        coarse_mv2 = np.ascontiguousarray(coarse_mv2, np.int16)
        ch2, cw2, _ = coarse_mv2.shape
        cw, ch = 2 * cw2, 2 * ch2
        out = np.empty((2 * ch // fine_block_size, 2 * cw // fine_block_size, 2), np.int16)
        _check(self._lib, self._ctx,
               self._lib.bbme_stage_copy_mvs(self._ctx, coarse_mv2.ctypes.data, cw, ch, coarse_block_size, fine_block_size,
                                             out.ctypes.data), "stage_copy_mvs")
        return out


class Pool:
    """All GPUs of the box behind one call (bbme_pool_*): the batch is cut into contiguous shards, one host thread per GPU
    runs its shard, the fields land in host arrays.  Frame pairs are independent, so nothing moves between GPUs."""

    def __init__(self, width, height, search_size, block_size, num_levels=None, sweeps=2, devices=None, chunk_pairs=8, slots=2):
        self._lib = _lib.load()
        self._pool = C.c_void_p()
        n = 0 if devices is None else len(devices)
        rc = self._lib.bbme_pool_create(C.byref(self._pool), n, None if devices is None else _int_array(devices))
        if rc != 0:
            msg = self._lib.bbme_last_error(None)
            raise BbmeError(rc, "bbme_pool_create: " + (msg.decode() if msg else ""))
        L = len(block_size) if num_levels is None else int(num_levels)
        opt = BbmeOptions()
        self._lib.bbme_default_options(C.byref(opt))
        opt.sweeps, opt.chunk_pairs, opt.slots = int(sweeps), int(chunk_pairs), int(slots)
        sh = BbmeShape()
        rc = self._lib.bbme_pool_plan(self._pool, int(width), int(height), L, _int_array(search_size[:L]),
                                      _int_array(block_size[:L]), C.byref(opt), C.byref(sh))
        if rc != 0:
            msg = self._lib.bbme_pool_last_error(self._pool).decode()
            self.close()
            raise BbmeError(rc, "bbme_pool_plan: " + msg)
        self.shape = _shape_dict(sh)
        self.device_count = self._lib.bbme_pool_device_count(self._pool)

    def close(self):
        if getattr(self, "_pool", None) is not None and self._pool:
            self._lib.bbme_pool_destroy(self._pool)
            self._pool = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def estimate_batch(self, im1_list, im2_list, out_list=None):
        """out_list: optional C-contiguous float32 (padded_h, padded_w, 2) arrays to fill (reused buffers avoid the page faults
        of fresh ones; pinned buffers from bbme_host_alloc make the copies asynchronous)."""
        n = len(im1_list)
        h, w = self.shape["height"], self.shape["width"]
        PA = C.c_void_p * n
        p1, p2, po = PA(), PA(), PA()
        fshape = (self.shape["padded_height"], self.shape["padded_width"], 2)
        outs = out_list if out_list is not None else [np.empty(fshape, np.float32) for _ in range(n)]
        for o in outs:
            if o.dtype != np.float32 or o.shape != fshape or not o.flags["C_CONTIGUOUS"]:
                raise BbmeError(-1, "flow buffers must be C-contiguous float32 (padded_h, padded_w, 2)")
        keep = []
        for i in range(n):
            a = np.ascontiguousarray(im1_list[i], np.uint8)
            b = np.ascontiguousarray(im2_list[i], np.uint8)
            if a.shape != (h, w) or b.shape != (h, w):
                raise BbmeError(-1, f"frames must be uint8 {h}x{w}")
            keep += [a, b]
            p1[i], p2[i], po[i] = a.ctypes.data, b.ctypes.data, outs[i].ctypes.data
        rc = self._lib.bbme_pool_estimate_batch(self._pool, n, p1, p2, w, po)
        if rc != 0:
            raise BbmeError(rc, "bbme_pool_estimate_batch: " + self._lib.bbme_pool_last_error(self._pool).decode())
        return outs


class MF:
    """Mirror of `class MF` (motion_framework.h:9-54).

    MF(image1, image2, search_size, block_size, num_levels) builds the padded pyramid on the GPU (the
    reference's constructor, motion_framework.cpp:4-111, does pad + pyrDown); calcMotionBlockMatching()
    returns the padded CV_32FC2-shaped field (motion_framework.cpp:218).  `sweeps` exposes the hard-coded 2
    of motion_framework.cpp:143,184.
    """

    def __init__(self, image1, image2, search_size, block_size, num_levels, sweeps=2, device=0, **opts):
        if num_levels <= 0:
            raise BbmeError(-1, "num_levels must be > 0")  # assert(num_levels > 0), motion_framework.cpp:7
        image1 = np.asarray(image1)
        image2 = np.asarray(image2)
        if image1.shape != image2.shape:
            raise BbmeError(-1, "image1 and image2 differ in size")  # assert, motion_framework.cpp:8
        h, w = image1.shape
        self._est = Estimator(w, h, list(search_size)[:num_levels], list(block_size)[:num_levels], num_levels,
                              sweeps=sweeps, device=device, **opts)
        sh = self._est.shape
        self.padded_height = sh["padded_height"]
        self.padded_width = sh["padded_width"]
        self.padding_x = sh["padding_x"]
        self.padding_y = sh["padding_y"]
        self._im1 = np.ascontiguousarray(image1, np.uint8)
        self._im2 = np.ascontiguousarray(image2, np.uint8)
        self.level_data = [
            PyramidLevel(sh["block_size"][l], sh["search_size"][l], float(sh["block_size"][l] // 2), sh["level_width"][l],
                         sh["level_height"][l]) for l in range(num_levels)
        ]

    def calcMotionBlockMatching(self):
        return self._est.estimate(self._im1, self._im2)

    def draw_MVimage(self):
        """Mirror of MF::draw_MVimage (motion_framework.cpp:887-905) as the commented call site :213-216 would use it after
        calcMotionBlockMatching(): block size 2, level 0 -- the motion-compensated padded frame."""
        return self._est.stage_compensate(self._est.level_image(0, 1), 2, self._est.level_mv(0, which=0))

    def stats(self):
        return self._est.stats()

    def close(self):
        self._est.close()


class Flow:
    """Mirror of `class Flow` (rw_flow.h:9-38) for the parts on the path: .flo codec and the AEE metric."""

    def __init__(self):
        self._lib = _lib.load()

    def ReadFlowFile(self, filename):
        if filename is None:
            raise BbmeError(-1, "ReadFlowFile: empty filename")
        w, h = C.c_int(0), C.c_int(0)
        rc = self._lib.bbme_flo_read_header(str(filename).encode(), C.byref(w), C.byref(h))
        if rc != 0:
            raise BbmeError(rc, "ReadFlowFile: " + self._lib.bbme_status_string(rc).decode())
        img = np.empty((h.value, w.value, 2), np.float32)
        rc = self._lib.bbme_flo_read(str(filename).encode(), img.ctypes.data, w.value, h.value)
        if rc != 0:
            raise BbmeError(rc, "ReadFlowFile: " + self._lib.bbme_status_string(rc).decode())
        return img

    def WriteFlowFile(self, img, filename):
        if filename is None:
            raise BbmeError(-1, "WriteFlowFile: empty filename")
        img = np.ascontiguousarray(img, np.float32)
        h, w, _ = img.shape
        rc = self._lib.bbme_flo_write(str(filename).encode(), img.ctypes.data, w, h)
        if rc != 0:
            raise BbmeError(rc, "WriteFlowFile: " + self._lib.bbme_status_string(rc).decode())

    def CalculateMSE(self, gtruth, flow):
        gt = np.ascontiguousarray(gtruth, np.float32)
        fl = np.ascontiguousarray(flow, np.float32)
        h, w, _ = gt.shape
        return float(self._lib.bbme_flow_aee(gt.ctypes.data, fl.ctypes.data, w, h))

    def MotionToColor(self, input_img, maxmotion=-1.0):
        """Flow::MotionToColor (rw_flow.cpp:202-249): (h, w, 2) float32 field -> (h, w, 3) uint8 image in OpenCV's BGR order."""
        f = np.ascontiguousarray(input_img, np.float32)
        h, w = f.shape[:2]
        out = np.empty((h, w, 3), np.uint8)
        rng = (C.c_float * 5)()
        rc = self._lib.bbme_flow_to_color(f.ctypes.data, w, h, C.c_float(maxmotion), out.ctypes.data, rng)
        if rc != 0:
            raise BbmeError(rc, "bbme_flow_to_color")
        self.last_motion_range = tuple(rng)
        return out

    def StripAndSubsample(self, padded_flow, shape, factor):
        """main()'s post-processing (main_class.cpp:58-70)."""
        sh = BbmeShape()
        for k in ("width", "height", "padded_width", "padded_height", "padding_x", "padding_y", "num_levels"):
            setattr(sh, k, shape[k])
        pf = np.ascontiguousarray(padded_flow, np.float32)
        out = np.empty((shape["height"] // factor, shape["width"] // factor, 2), np.float32)
        rc = self._lib.bbme_flow_strip_subsample(pf.ctypes.data, C.byref(sh), int(factor), out.ctypes.data)
        if rc != 0:
            raise BbmeError(rc, "StripAndSubsample: " + self._lib.bbme_status_string(rc).decode())
        return out
