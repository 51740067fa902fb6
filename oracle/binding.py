"""ctypes binding of oracle/libbbme_oracle.so and, when built, oracle/_ref/libbbme_ref.so.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libbbme_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbbme_ref.so")
MAXL = 16


class OrcShape(C.Structure):
    _fields_ = [("padded_w", C.c_int), ("padded_h", C.c_int), ("pad_x", C.c_int), ("pad_y", C.c_int),
                ("num_levels", C.c_int), ("level_w", C.c_int * MAXL), ("level_h", C.c_int * MAXL)]


class OrcStats(C.Structure):
    _fields_ = [("t_ctor_s", C.c_double), ("t_run_s", C.c_double),
                ("search_sad_calls", C.c_uint64), ("search_absdiffs", C.c_uint64),
                ("reg_sad_calls", C.c_uint64), ("reg_absdiffs", C.c_uint64),
                ("level_search_absdiffs", C.c_uint64 * MAXL), ("level_reg_absdiffs", C.c_uint64 * MAXL)]


_lib = None
_ref = None


def build(force=False):
    if force or not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", HERE, "libbbme_oracle.so"], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(ORACLE_SO)
        lib.orc_aee.restype = C.c_double
        lib.orc_regularize_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_int]
        lib.orc_estimate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_estimate_debug.argtypes = lib.orc_estimate.argtypes + [C.c_void_p] * 4
        lib.orc_estimate_raster.argtypes = lib.orc_estimate.argtypes[:-1]
        lib.orc_estimate_many.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.orc_pad_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        lib.orc_pyrdown.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.orc_search_level.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_int]
        lib.orc_search_level_raster.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.orc_compensate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_divide_blocks.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.orc_copy_mvs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.orc_copy_to_all_pixels.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.orc_flo_read.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        lib.orc_flo_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        lib.orc_aee.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.orc_resize_linear.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.orc_strip_subsample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def _ia(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def stats_dict(st):
    d = {k: getattr(st, k) for k, _ in OrcStats._fields_ if not k.startswith("level_")}
    d["level_search_absdiffs"] = list(st.level_search_absdiffs)
    d["level_reg_absdiffs"] = list(st.level_reg_absdiffs)
    return d


def plan_shape(w, h, block_size):
    lib = load()
    sh = OrcShape()
    rc = lib.orc_plan_shape(int(w), int(h), len(block_size), _ia(block_size), C.byref(sh))
    L = len(block_size)
    return rc, {"padded_width": sh.padded_w, "padded_height": sh.padded_h, "padding_x": sh.pad_x, "padding_y": sh.pad_y,
                "level_width": list(sh.level_w[:L]), "level_height": list(sh.level_h[:L])}


def pyrdown(src):
    lib = load()
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h // 2, w // 2), np.uint8)
    lib.orc_pyrdown(src.ctypes.data, w, h, dst.ctypes.data)
    return dst


def pad_image(src, pad_x, pad_y):
    lib = load()
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h + 2 * pad_y, w + 2 * pad_x), np.uint8)
    lib.orc_pad_image(src.ctypes.data, w, h, w, pad_x, pad_y, dst.ctypes.data)
    return dst


def spiral_walk(shift):
    lib = load()
    cap = (shift + 3) * (shift + 3) + 8
    buf = (C.c_int * (2 * cap))()
    n = lib.orc_spiral_walk(int(shift), buf, cap)
    return np.array(buf[:2 * n], np.int32).reshape(n, 2)


def spiral_rank(dx, dy):
    return load().orc_spiral_rank(int(dx), int(dy))


def estimate(im1, im2, search_size, block_size, sweeps=2, debug=False):
    """Whole pair.  Returns (flow[Hp,Wp,2] float32, stats dict[, debug dict])."""
    lib = load()
    im1 = np.ascontiguousarray(im1, np.uint8)
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im1.shape
    L = len(block_size)
    rc, sh = plan_shape(w, h, block_size)
    if rc != 0:
        raise ValueError(f"oracle plan_shape failed: {rc}")
    flow = np.empty((sh["padded_height"], sh["padded_width"], 2), np.float32)
    st = OrcStats()
    if not debug:
        rc = lib.orc_estimate(im1.ctypes.data, im2.ctypes.data, w, h, w, L, _ia(search_size), _ia(block_size), sweeps,
                              flow.ctypes.data, C.byref(st))
        if rc != 0:
            raise ValueError(f"oracle estimate failed: {rc}")
        return flow, stats_dict(st)
    pyr1 = [np.empty((sh["level_height"][l], sh["level_width"][l]), np.uint8) for l in range(L)]
    pyr2 = [np.empty_like(p) for p in pyr1]
    a_s = [np.empty((sh["level_height"][l], sh["level_width"][l], 2), np.float32) for l in range(L)]
    a_r = [np.empty_like(p) for p in a_s]

    def ptrs(arrs):
        return (C.c_void_p * L)(*[a.ctypes.data for a in arrs])

    rc = lib.orc_estimate_debug(im1.ctypes.data, im2.ctypes.data, w, h, w, L, _ia(search_size), _ia(block_size), sweeps,
                                flow.ctypes.data, C.byref(st), ptrs(pyr1), ptrs(pyr2), ptrs(a_s), ptrs(a_r))
    if rc != 0:
        raise ValueError(f"oracle estimate failed: {rc}")
    return flow, stats_dict(st), {"pyr1": pyr1, "pyr2": pyr2, "after_search": a_s, "after_reg": a_r, "shape": sh}


def estimate_raster(im1, im2, search_size, block_size, sweeps=2):
    """Whole pair with find_min_block (raster scan, L1-distance tie-break) as the per-level search."""
    lib = load()
    im1 = np.ascontiguousarray(im1, np.uint8)
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im1.shape
    rc, sh = plan_shape(w, h, block_size)
    if rc != 0:
        raise ValueError(f"oracle plan_shape failed: {rc}")
    flow = np.empty((sh["padded_height"], sh["padded_width"], 2), np.float32)
    rc = lib.orc_estimate_raster(im1.ctypes.data, im2.ctypes.data, w, h, w, len(block_size), _ia(search_size), _ia(block_size),
                                 sweeps, flow.ctypes.data)
    if rc != 0:
        raise ValueError(f"oracle estimate_raster failed: {rc}")
    return flow


def estimate_many(pairs, search_size, block_size, sweeps=2, threads=1):
    lib = load()
    n = len(pairs)
    h, w = pairs[0][0].shape
    rc, sh = plan_shape(w, h, block_size)
    if rc != 0:
        raise ValueError(f"oracle plan_shape failed: {rc}")
    a = [np.ascontiguousarray(p[0], np.uint8) for p in pairs]
    b = [np.ascontiguousarray(p[1], np.uint8) for p in pairs]
    out = [np.empty((sh["padded_height"], sh["padded_width"], 2), np.float32) for _ in range(n)]
    P = C.c_void_p * n
    st = OrcStats()
    rc = lib.orc_estimate_many(n, P(*[x.ctypes.data for x in a]), P(*[x.ctypes.data for x in b]), w, h, w,
                               len(block_size), _ia(search_size), _ia(block_size), sweeps,
                               P(*[x.ctypes.data for x in out]), int(threads), C.byref(st))
    if rc != 0:
        raise ValueError(f"oracle estimate_many failed: {rc}")
    return out, stats_dict(st)


def search_level(im1, im2, block_size, search_size, flow):
    lib = load()
    im1 = np.ascontiguousarray(im1, np.uint8)
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im1.shape
    flow = np.ascontiguousarray(flow, np.float32).copy()
    st = OrcStats()
    lib.orc_search_level(im1.ctypes.data, im2.ctypes.data, w, h, block_size, search_size, flow.ctypes.data, C.byref(st), 0)
    return flow, stats_dict(st)


def search_level_raster(im1, im2, block_size, search_size, flow):
    """calcLevelBM with find_min_block (motion_framework.cpp:246-294): raster scan, L1-distance tie-break."""
    lib = load()
    im1 = np.ascontiguousarray(im1, np.uint8)
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im1.shape
    flow = np.ascontiguousarray(flow, np.float32).copy()
    lib.orc_search_level_raster(im1.ctypes.data, im2.ctypes.data, w, h, block_size, search_size, flow.ctypes.data)
    return flow


def compensate(im2, block_size, flow):
    """draw_MVimage (motion_framework.cpp:887-905) into a zero-filled frame."""
    lib = load()
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im2.shape
    flow = np.ascontiguousarray(flow, np.float32)
    out = np.zeros((h, w), np.uint8)
    lib.orc_compensate(im2.ctypes.data, w, h, block_size, flow.ctypes.data, out.ctypes.data)
    return out


def regularize_sweep(im1, im2, block_size, lam, mult, flow):
    lib = load()
    im1 = np.ascontiguousarray(im1, np.uint8)
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im1.shape
    flow = np.ascontiguousarray(flow, np.float32).copy()
    lib.orc_regularize_sweep(im1.ctypes.data, im2.ctypes.data, w, h, block_size, C.c_float(lam), mult, flow.ctypes.data, None, 0)
    return flow


def divide_blocks(flow, block_size):
    lib = load()
    flow = np.ascontiguousarray(flow, np.float32).copy()
    h, w, _ = flow.shape
    lib.orc_divide_blocks(w, h, block_size, flow.ctypes.data)
    return flow


def copy_mvs(coarse, coarse_block_size):
    lib = load()
    coarse = np.ascontiguousarray(coarse, np.float32)
    ch, cw, _ = coarse.shape
    fine = np.zeros((2 * ch, 2 * cw, 2), np.float32)
    lib.orc_copy_mvs(coarse.ctypes.data, cw, ch, coarse_block_size, fine.ctypes.data)
    return fine


def resize_linear(src, factor):
    """cv::resize(src, Size(), factor, factor, INTER_LINEAR) for 8-bit (main_class.cpp:32-33)."""
    lib = load()
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h * factor, w * factor), np.uint8)
    lib.orc_resize_linear(src.ctypes.data, w, h, int(factor), dst.ctypes.data)
    return dst


def strip_subsample(flow, pad_x, pad_y, factor):
    """main_class.cpp:58-70: padded dense field -> (h / factor) x (w / factor) x 2 sub-pixel field."""
    lib = load()
    flow = np.ascontiguousarray(flow, np.float32)
    ph, pw = flow.shape[:2]
    out = np.zeros(((ph - 2 * pad_y) // factor, (pw - 2 * pad_x) // factor, 2), np.float32)
    lib.orc_strip_subsample(flow.ctypes.data, pw, ph, int(pad_x), int(pad_y), int(factor), out.ctypes.data)
    return out


def flo_read(path):
    lib = load()
    w, h = C.c_int(0), C.c_int(0)
    rc = lib.orc_flo_read_header(str(path).encode(), C.byref(w), C.byref(h))
    if rc != 0:
        raise ValueError(f"orc_flo_read_header: {rc}")
    a = np.empty((h.value, w.value, 2), np.float32)
    rc = lib.orc_flo_read(str(path).encode(), a.ctypes.data, w.value, h.value)
    if rc != 0:
        raise ValueError(f"orc_flo_read: {rc}")
    return a


def flo_write(path, a):
    lib = load()
    a = np.ascontiguousarray(a, np.float32)
    h, w, _ = a.shape
    return lib.orc_flo_write(str(path).encode(), a.ctypes.data, w, h)


def aee(gt, flow):
    lib = load()
    gt = np.ascontiguousarray(gt, np.float32)
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = gt.shape
    return float(lib.orc_aee(gt.ctypes.data, flow.ctypes.data, w, h))


# ---- block-granular <-> dense helpers shared by the parity tests
def dense_to_blocks(flow, g):
    """Dense float field -> int16 (H/g, W/g, 2) sampling each block's top-left pixel."""
    return np.ascontiguousarray(np.rint(flow[::g, ::g, :]).astype(np.int16))


def blocks_to_dense(mv, g, h, w):
    """int16 block field -> dense float field with the MV only at block corners (rest zero), like level_flow."""
    out = np.zeros((h, w, 2), np.float32)
    out[::g, ::g, :] = mv.astype(np.float32)
    return out


# ---- the real reference compiled against oracle/cvshim (oracle/_ref)
def load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            return None
        lib = C.CDLL(REF_SO)
        lib.ref_mf_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.ref_flow_read.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        lib.ref_flow_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_flow_mse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_flow_mse.restype = C.c_double
        lib.ref_flow_color.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        lib.ref_find_min_block_level.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.ref_draw_mvimage.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _ref = lib
    return _ref


def ref_estimate(im1, im2, search_size, block_size):
    """Runs the reference's own MF (2 sweeps, hard-coded) -> (flow, shape4, t_ctor, t_run).  None if _ref is absent."""
    lib = load_ref()
    if lib is None:
        return None
    im1 = np.ascontiguousarray(im1, np.uint8)
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im1.shape
    rc, sh = plan_shape(w, h, block_size)
    if rc != 0:
        raise ValueError("shape not runnable by the reference")
    flow = np.empty((sh["padded_height"], sh["padded_width"], 2), np.float32)
    dims = (C.c_int * 4)()
    tc, tr = C.c_double(0), C.c_double(0)
    rc = lib.ref_mf_run(im1.ctypes.data, im2.ctypes.data, w, h, _ia(search_size), _ia(block_size), len(block_size),
                        flow.ctypes.data, dims, C.byref(tc), C.byref(tr))
    if rc != 0:
        raise ValueError(f"ref_mf_run failed: {rc}")
    return flow, list(dims), tc.value, tr.value


def ref_flow_color(flow, maxmotion=-1.0):
    """The reference's own Flow::MotionToColor (oracle/_ref); None when oracle/_ref is not built."""
    lib = load_ref()
    if lib is None:
        return None
    flow = np.ascontiguousarray(flow, np.float32)
    h, w = flow.shape[:2]
    out = np.zeros((h, w, 3), np.uint8)
    rc = lib.ref_flow_color(flow.ctypes.data, w, h, C.c_float(maxmotion), out.ctypes.data)
    if rc != 0:
        raise RuntimeError("ref_flow_color failed")
    return out


def ref_search_level_raster(im1, im2, block_size, search_size, flow):
    """The reference's own find_min_block (:246-294) driven block by block like calcLevelBM; None without oracle/_ref."""
    lib = load_ref()
    if lib is None:
        return None
    im1 = np.ascontiguousarray(im1, np.uint8)
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im1.shape
    flow = np.ascontiguousarray(flow, np.float32).copy()
    rc = lib.ref_find_min_block_level(im1.ctypes.data, im2.ctypes.data, w, h, block_size, search_size, flow.ctypes.data)
    if rc != 0:
        raise ValueError("ref_find_min_block_level: the size must be a multiple of the block size")
    return flow


def ref_compensate(im1, im2, block_size, flow):
    """The reference's own draw_MVimage (:887-905) into a zero-filled frame; None without oracle/_ref."""
    lib = load_ref()
    if lib is None:
        return None
    im1 = np.ascontiguousarray(im1, np.uint8)
    im2 = np.ascontiguousarray(im2, np.uint8)
    h, w = im2.shape
    flow = np.ascontiguousarray(flow, np.float32)
    out = np.zeros((h, w), np.uint8)
    rc = lib.ref_draw_mvimage(im1.ctypes.data, im2.ctypes.data, w, h, block_size, flow.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise ValueError("ref_draw_mvimage: the size must be a multiple of the block size")
    return out
