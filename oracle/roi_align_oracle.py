"""oracle/roi_align_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/ctypes front-end of oracle/roi_align_oracle.c (the plain-C restatement of
mmcv-1.4.7/mmcv/ops/csrc/pytorch/cpu/roi_align.cpp).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; gpt4roi_b200/ must never do so.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'liboracle.so')
NCHW, NHWC = 0, 1

_lib = None


def build(force=False):
    src = os.path.join(HERE, 'roi_align_oracle.c')
    if force or not os.path.isfile(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', HERE, '-s', '-B', 'liboracle.so'])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return 'f32', ctypes.c_float
    if dtype == np.float64:
        return 'f64', ctypes.c_double
    raise TypeError('oracle supports float32/float64, got %s' % dtype)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleError(RuntimeError):
    pass


def roi_align_forward(inp, rois, output_size, spatial_scale=1.0, sampling_ratio=0,
                      pool_mode='avg', aligned=True, in_layout=NCHW, out_layout=NCHW):
    """Reference semantics of mmcv.ops.roi_align forward on CPU.

    inp: [N,C,H,W] (NCHW) or [N,H,W,C] (NHWC); rois [K,5] same dtype.
    Returns (output, argmax_y, argmax_x); argmax_* are None for 'avg'.
    """
    inp = np.ascontiguousarray(inp)
    sfx, cty = _suffix(inp.dtype)
    rois = np.ascontiguousarray(rois, dtype=inp.dtype)
    if in_layout == NCHW:
        n, c, h, w = inp.shape
    else:
        n, h, w, c = inp.shape
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    k = rois.shape[0]
    shape = (k, c, ph, pw) if out_layout == NCHW else (k, ph, pw, c)
    out = np.zeros(shape, dtype=inp.dtype)
    mode = {'max': 0, 'avg': 1}[pool_mode]
    if mode == 0:
        ay, ax = np.zeros(shape, dtype=inp.dtype), np.zeros(shape, dtype=inp.dtype)
    else:
        ay = ax = np.zeros(1, dtype=inp.dtype)
    fn = getattr(lib(), 'roi_oracle_forward_' + sfx)
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 6 + [cty] + [ctypes.c_int] * 5
    rc = fn(_ptr(inp), _ptr(rois), _ptr(out), _ptr(ay), _ptr(ax), k, c, h, w, ph, pw,
            float(np.float32(spatial_scale)), int(sampling_ratio), mode, int(bool(aligned)),
            int(in_layout), int(out_layout))
    if rc == 2:
        raise OracleError('ROIs in ROIAlign cannot have non-negative size!')
    if rc:
        raise OracleError('oracle forward failed rc=%d' % rc)
    return (out, ay, ax) if mode == 0 else (out, None, None)


def roi_align_backward(grad_out, rois, input_shape, spatial_scale=1.0, sampling_ratio=0,
                       pool_mode='avg', aligned=True, argmax_y=None, argmax_x=None,
                       in_layout=NCHW, out_layout=NCHW):
    """Reference semantics of mmcv roi_align backward (sequential accumulation)."""
    grad_out = np.ascontiguousarray(grad_out)
    sfx, cty = _suffix(grad_out.dtype)
    rois = np.ascontiguousarray(rois, dtype=grad_out.dtype)
    if in_layout == NCHW:
        n, c, h, w = input_shape
    else:
        n, h, w, c = input_shape
    if out_layout == NCHW:
        k, _, ph, pw = grad_out.shape
    else:
        k, ph, pw, _ = grad_out.shape
    gin = np.zeros(input_shape, dtype=grad_out.dtype)
    mode = {'max': 0, 'avg': 1}[pool_mode]
    dummy = np.zeros(1, dtype=grad_out.dtype)
    ay = np.ascontiguousarray(argmax_y) if mode == 0 else dummy
    ax = np.ascontiguousarray(argmax_x) if mode == 0 else dummy
    fn = getattr(lib(), 'roi_oracle_backward_' + sfx)
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 6 + [cty] + [ctypes.c_int] * 5
    rc = fn(_ptr(grad_out), _ptr(rois), _ptr(ay), _ptr(ax), _ptr(gin), k, c, h, w, ph, pw,
            float(np.float32(spatial_scale)), int(sampling_ratio), mode, int(bool(aligned)),
            int(in_layout), int(out_layout))
    if rc == 2:
        raise OracleError('ROIs in ROIAlign do not have non-negative size!')
    if rc:
        raise OracleError('oracle backward failed rc=%d' % rc)
    return gin


class _PreCalc32(ctypes.Structure):
    _fields_ = [('pos', ctypes.c_int * 4), ('w', ctypes.c_float * 4),
                ('y', ctypes.c_float), ('x', ctypes.c_float)]


def sample_table(roi5, height, width, output_size, spatial_scale, sampling_ratio, aligned=True):
    """Per-RoI (pos1..4, w1..4, y, x) table, fp32 -- the PreCalc of cpu/roi_align.cpp:23-108.

    Returns (pos int32 [PH,PW,gh,gw,4], w float32 [PH,PW,gh,gw,4], yx float32 [PH,PW,gh,gw,2]).
    """
    roi5 = np.ascontiguousarray(roi5, dtype=np.float32)
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    gh, gw = ctypes.c_int(), ctypes.c_int()
    tab = ctypes.POINTER(_PreCalc32)()
    fn = lib().roi_oracle_sample_table_f32
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_float] + \
        [ctypes.c_int] * 2 + [ctypes.c_void_p] * 3
    rc = fn(_ptr(roi5), height, width, ph, pw, float(np.float32(spatial_scale)),
            int(sampling_ratio), int(bool(aligned)), ctypes.byref(gh), ctypes.byref(gw),
            ctypes.byref(tab))
    if rc:
        raise OracleError('sample_table rc=%d' % rc)
    n = ph * pw * gh.value * gw.value
    pos = np.zeros((n, 4), np.int32)
    wts = np.zeros((n, 4), np.float32)
    yx = np.zeros((n, 2), np.float32)
    for i in range(n):
        pos[i] = tab[i].pos[:]
        wts[i] = tab[i].w[:]
        yx[i] = (tab[i].y, tab[i].x)
    free = lib().roi_oracle_free_f32
    free.argtypes = [ctypes.c_void_p]
    free.restype = None
    free(tab)
    shp = (ph, pw, gh.value, gw.value)
    return pos.reshape(shp + (4,)), wts.reshape(shp + (4,)), yx.reshape(shp + (2,))


def splice(input_ids, embed_table, image_rows, region_rows, region_offsets, P,
           im_patch, im_start, im_end, bbox_tok):
    """Region-token splice (gpt4roi/models/spi_llava.py:99-196) on 16-bit rows.

    input_ids int64 [B,L]; embed_table uint16 [V,D]; image_rows uint16 [B,P,D];
    region_rows uint16 [K,D] or None; region_offsets int32 [B+1] or None.
    Raises ValueError with the reference's messages on malformed spans.
    """
    ids = np.ascontiguousarray(input_ids, dtype=np.int64)
    emb = np.ascontiguousarray(embed_table)
    assert emb.dtype.itemsize == 2
    b, l = ids.shape
    d = emb.shape[1]
    img = np.ascontiguousarray(image_rows)
    out = np.zeros((b, l, d), dtype=emb.dtype)
    if region_offsets is not None:
        ro = np.ascontiguousarray(region_offsets, dtype=np.int32)
        rr = np.ascontiguousarray(region_rows) if region_rows is not None and len(region_rows) \
            else np.zeros((1, d), emb.dtype)
        rop, rrp = _ptr(ro), _ptr(rr)
    else:
        rop = rrp = None
    fn = lib().splice_oracle
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 3 + [ctypes.c_size_t] + \
        [ctypes.c_int64] * 4
    rc = fn(_ptr(ids), _ptr(emb), _ptr(img), rrp, rop, _ptr(out), b, l, int(P), d * 2,
            int(im_patch), int(im_start), int(im_end), int(bbox_tok))
    if rc == 1:
        raise ValueError('The number of image start tokens and image end tokens should be the same.')
    if rc == 2:
        raise ValueError('The image end token should follow the image start token.')
    if rc == 3:
        raise ValueError('number of <bbox> tokens does not match the number of boxes')
    if rc == 4:
        raise ValueError('more than one <im_start> per sample is not supported')
    return out
