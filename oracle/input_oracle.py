"""oracle/input_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, integer-exact) of the image side of the reference's input pipeline
(gpt4roi/datasets/coco_det.py:55-83), i.e. of the mmdet / mmcv transforms it lists:

    Resize(img_scale=(S,S), keep_ratio=False)   mmdet/datasets/pipelines/transforms.py:209-243 -> mmcv.imresize ->
                                                cv2.resize(..., interpolation=cv2.INTER_LINEAR) on uint8 HWC
    RandomShift(0.5, 32)                        transforms.py:505-562  (zero-filled shift of the resized uint8 image)
    FilterAnnotations((2,2))                    mmdet/datasets/pipelines/loading.py:578-596
    RandomFlip(0.5)                             transforms.py:388-420 (horizontal)
    Normalize(mean, std, to_rgb=True)           mmcv/image/photometric.py imnormalize_: BGR->RGB, f32(x - mean32), then x (1/std) in fp64 -> fp32
    Pad(size_divisor)                           zero padding of the normalised image (no-op for S x S, S % divisor == 0)
    boxes / S                                   coco_det.py:154

cv2's 8-bit bilinear resize is fixed-point (third-party: OpenCV imgproc/src/resize.cpp, INTER_RESIZE_COEF_BITS = 11):
    fx = float((dx + 0.5) * (1. / (dst_w / src_w)) - 0.5); sx = floor(fx); fx -= sx; columns clamp (sx < 0 -> 0, fx = 0; sx >= w-1 ->
    w-1, fx = 0), rows keep fy and clip the two row indices to [0, h-1]
    alpha = (round((1 - fx) * 2048), round(fx * 2048))  as int16; same for rows (beta)
    row pass   : T[y][dx] = S[y][sx] * alpha0 + S[y][sx+1] * alpha1                       (int32)
    column pass: dst = (((beta0 * (T0 >> 4)) >> 16) + ((beta1 * (T1 >> 4)) >> 16) + 2) >> 2
Pinned by tests/test_input_pipeline_cpu.py against cv2.resize itself (bit-exact on random images and sizes) and against
the reference's own transform classes run under tests/golden/ref_shims.py (tests/golden/input_pipeline_ref.npz)."""
import numpy as np

CLIP_MEAN = (0.48145466 * 255, 0.4578275 * 255, 0.40821073 * 255)   # coco_det.py:56-58
CLIP_STD = (0.26862954 * 255, 0.26130258 * 255, 0.27577711 * 255)


def _axis_table(src, dst, rows=False):
    scale = 1.0 / (float(dst) / float(src))                # double, as cv2: inv_scale_x = (double)dsize.width / ssize.width; scale_x = 1. / inv_scale_x
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)      # fx = (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not rows:   # columns: the fraction is dropped at the borders (resize.cpp: "fx = 0, sx = 0" / "sx = width-1")
        lo = s < 0
        s[lo], f[lo] = 0, 0.0
        hi = s >= src - 1
        s[hi], f[hi] = src - 1, 0.0
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int32)     # saturate_cast<short>(cbuf * 2048): cvRound
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int32)
    # rows: the fraction is KEPT and only the two row indices are clipped (srows[k] = clip(sy + k, 0, height - 1))
    s0 = np.clip(s, 0, src - 1)
    s1 = np.clip(s + 1, 0, src - 1)
    return s0, s1, a0, a1


def resize_linear_u8(img, dst_h, dst_w):
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC, integer-exact."""
    h, w = img.shape[:2]
    sx0, sx1, ax0, ax1 = _axis_table(w, dst_w)
    sy0, sy1, by0, by1 = _axis_table(h, dst_h, rows=True)
    src = img.astype(np.int32)
    t = src[:, sx0] * ax0[None, :, None] + src[:, sx1] * ax1[None, :, None]          # [h, dst_w, c] int32
    t0, t1 = t[sy0], t[sy1]
    out = (((by0[:, None, None] * (t0 >> 4)) >> 16) + ((by1[:, None, None] * (t1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def shift_u8(img, shift_x, shift_y):
    """RandomShift's image part (transforms.py:553-561)."""
    out = np.zeros_like(img)
    h, w = img.shape[:2]
    nx, ox, ny, oy = max(0, shift_x), max(0, -shift_x), max(0, shift_y), max(0, -shift_y)
    nh, nw = h - abs(shift_y), w - abs(shift_x)
    out[ny:ny + nh, nx:nx + nw] = img[oy:oy + nh, ox:ox + nw]
    return out


def normalize(img_u8, mean=CLIP_MEAN, std=CLIP_STD, to_rgb=True):
    """mmdet Normalize -> mmcv.imnormalize_: mean / std are stored as float32 (transforms.py Normalize.__init__), then
    `cv2.subtract(img32, float64(mean32))` (a float32 subtraction) and `cv2.multiply(img32, 1 / float64(std32))` -- cv2
    multiplies a float32 image by a double scalar in DOUBLE and rounds once (checked against cv2 for all 256 x 3 inputs)."""
    x = img_u8.astype(np.float32)
    if to_rgb:
        x = x[..., ::-1]
    m32 = np.asarray(mean, np.float32)
    sinv = 1.0 / np.asarray(std, np.float32).astype(np.float64)
    return ((x - m32).astype(np.float64) * sinv).astype(np.float32)


def transform_boxes(boxes, src_h, src_w, S, shift=(0, 0), flip=False, shift_filter_px=1, min_wh=(2.0, 2.0)):
    """Box side of Resize -> RandomShift -> FilterAnnotations -> RandomFlip -> /S in float32, as the reference computes it.
    Returns (normalised boxes [K',4], keep indices into the input, shift actually applied): RandomShift leaves image and
    boxes untouched when no box would survive it (transforms.py:541-544); FilterAnnotations returning "no box left"
    (the dataset then draws another sample) is reported as an empty result."""
    b = np.asarray(boxes, np.float32).reshape(-1, 4)
    idx = np.arange(len(b))
    sf = np.array([S / src_w, S / src_h, S / src_w, S / src_h], dtype=np.float32)     # transforms.py:229-230
    b = b * sf
    b[:, 0::2] = np.clip(b[:, 0::2], 0, S)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, S)
    applied = (0, 0)
    if shift != (0, 0) or shift is None:
        sb = b.copy()
        sb[:, 0::2] += shift[0]
        sb[:, 1::2] += shift[1]
        sb[:, 0::2] = np.clip(sb[:, 0::2], 0, S)
        sb[:, 1::2] = np.clip(sb[:, 1::2], 0, S)
        valid = ((sb[:, 2] - sb[:, 0]) > shift_filter_px) & ((sb[:, 3] - sb[:, 1]) > shift_filter_px)
        if valid.any():
            b, idx, applied = sb[valid], idx[valid], (int(shift[0]), int(shift[1]))
    if len(b):
        keep = ((b[:, 2] - b[:, 0]) > min_wh[0]) & ((b[:, 3] - b[:, 1]) > min_wh[1])
        b, idx = b[keep], idx[keep]
    if flip and len(b):
        f = b.copy()
        f[:, 0] = S - b[:, 2]
        f[:, 2] = S - b[:, 0]
        b = f
    return (b / np.float32(S)).astype(np.float32), idx, applied


def preprocess_image(img_bgr_u8, S, shift=(0, 0), flip=False, mean=CLIP_MEAN, std=CLIP_STD):
    """uint8 HWC BGR image -> float32 CHW normalised RGB [3,S,S] (Resize -> RandomShift -> RandomFlip -> Normalize -> Pad
    -> DefaultFormatBundle's HWC->CHW)."""
    x = resize_linear_u8(img_bgr_u8, S, S)
    if shift != (0, 0):
        x = shift_u8(x, shift[0], shift[1])
    if flip:
        x = x[:, ::-1]
    return np.ascontiguousarray(normalize(x, mean, std).transpose(2, 0, 1))
