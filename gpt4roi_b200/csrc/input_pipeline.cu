// input_pipeline.cu -- image side of the reference's input pipeline as ONE kernel.
//
// Replaces, for a whole batch of decoded uint8 HWC BGR images of arbitrary sizes (reference: gpt4roi/datasets/
// coco_det.py:60-71, executed per sample by python dataloader workers):
//   Resize(img_scale=(S,S), keep_ratio=False)  mmdet/datasets/pipelines/transforms.py:209-243 -> mmcv.imresize ->
//                                              cv2.resize(INTER_LINEAR) on uint8
//   RandomShift                                transforms.py:553-561 (zero-filled shift of the resized image)
//   RandomFlip (horizontal)                    transforms.py:422-470
//   Normalize(mean, std, to_rgb=True)          mmcv imnormalize_: BGR->RGB, f32(x - mean32), x (1 / f64(std32)) in double
//   Pad(size_divisor = S) + DefaultFormatBundle  (no-op + HWC -> CHW)
// Output [B,3,S,S] fp32 (bit-identical to the reference pipeline) or bf16 (one more rounding; the engine's input dtype).
//
// Integer-exact restatement of OpenCV's 8-bit bilinear resize (imgproc/src/resize.cpp, INTER_RESIZE_COEF_BITS = 11):
//   f = float((d + 0.5) * (1. / (dst / src)) - 0.5); s = floor(f); f -= s;
//   columns: s < 0 -> (0, f = 0); s >= w-1 -> (w-1, f = 0); rows keep f and clip the two row indices;
//   a = (rint((1-f)*2048), rint(f*2048));  T = S[s]*a0 + S[s+1]*a1;  dst = (((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2
// The double / float expressions use explicit round-to-nearest intrinsics so that nvcc cannot contract them into FMAs.
// One thread per output pixel; the 4 x 3 source bytes of a pixel come from L2 (a source image is a few hundred KB).
#include "common.cuh"

namespace g4r {

struct PrepParams {
  const unsigned char* src;      // packed images, image b at src + offs[b], HWC, 3 channels (BGR)
  const long long* offs;
  const int* hw;                 // [B][2] source (h, w)
  const int* shift;              // [B][2] (shift_x, shift_y) applied to the resized image, or null
  const int* flip;               // [B] horizontal flip flag, or null
  void* out;
  int B, S, to_rgb, out_bf16;
  float mean[3];
  double stdinv[3];
};

__device__ __forceinline__ void axis_entry_u8(int d, int src, int dst, bool rows, int& s0, int& s1, int& a0, int& a1) {
  const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dst, (double)src));
  const float fx0 = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
  int s = (int)floorf(fx0);
  float f = __fsub_rn(fx0, (float)s);
  if (!rows) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
  }
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  s0 = min(max(s, 0), src - 1);
  s1 = min(max(s + 1, 0), src - 1);
}

__global__ void __launch_bounds__(256)
preprocess_images(const PrepParams p) {
  const int b = blockIdx.y;
  const int S = p.S;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= S * S) return;
  const int y = pix / S, x = pix - y * S;
  const int h = p.hw[2 * b], w = p.hw[2 * b + 1];
  const int shx = p.shift ? p.shift[2 * b] : 0, shy = p.shift ? p.shift[2 * b + 1] : 0;
  const int xf = (p.flip && p.flip[b]) ? S - 1 - x : x;
  const int rx = xf - shx, ry = y - shy;          // position in the resized (unshifted, unflipped) image
  int v[3] = {0, 0, 0};
  if (rx >= 0 && rx < S && ry >= 0 && ry < S) {
    int x0, x1, ax0, ax1, y0, y1, by0, by1;
    axis_entry_u8(rx, w, S, false, x0, x1, ax0, ax1);
    axis_entry_u8(ry, h, S, true, y0, y1, by0, by1);
    const unsigned char* img = p.src + p.offs[b];
    const unsigned char* r0 = img + (size_t)y0 * w * 3;
    const unsigned char* r1 = img + (size_t)y1 * w * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int t0 = (int)r0[x0 * 3 + c] * ax0 + (int)r0[x1 * 3 + c] * ax1;
      const int t1 = (int)r1[x0 * 3 + c] * ax0 + (int)r1[x1 * 3 + c] * ax1;
      v[c] = (((by0 * (t0 >> 4)) >> 16) + ((by1 * (t1 >> 4)) >> 16) + 2) >> 2;
    }
  }
  const size_t plane = (size_t)S * S;
#pragma unroll
  for (int co = 0; co < 3; co++) {
    const int ci = p.to_rgb ? 2 - co : co;
    const float d = __fsub_rn((float)v[ci], p.mean[co]);
    const float o = (float)__dmul_rn((double)d, p.stdinv[co]);
    const size_t idx = ((size_t)b * 3 + co) * plane + pix;
    if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[idx] = __float2bfloat16_rn(o);
    else reinterpret_cast<float*>(p.out)[idx] = o;
  }
}

}  // namespace g4r

using namespace g4r;

extern "C" int g4r_preprocess_images(const void* src_packed, const long long* offsets, const int* src_hw, const int* shift_xy,
                                     const int* flip, void* out, int B, int S, const float* mean3, const float* std3,
                                     int to_rgb, int out_dtype, void* stream) {
  G4R_REQUIRE(src_packed && offsets && src_hw && out && mean3 && std3, "preprocess_images: null argument");
  G4R_REQUIRE(B > 0 && B <= 65535 && S > 0 && S <= 4096, "preprocess_images: bad sizes B=%d S=%d", B, S);
  G4R_REQUIRE(out_dtype == G4R_F32 || out_dtype == G4R_BF16, "preprocess_images: out_dtype must be fp32 or bf16");
  PrepParams p{};
  p.src = (const unsigned char*)src_packed; p.offs = offsets; p.hw = src_hw; p.shift = shift_xy; p.flip = flip; p.out = out;
  p.B = B; p.S = S; p.to_rgb = to_rgb; p.out_bf16 = out_dtype == G4R_BF16;
  for (int c = 0; c < 3; c++) {
    p.mean[c] = mean3[c];                       // float32, as mmdet's Normalize stores it
    p.stdinv[c] = 1.0 / (double)std3[c];        // 1 / float64(std32), as mmcv.imnormalize_ computes it
  }
  dim3 grid((unsigned)((S * S + 255) / 256), (unsigned)B);
  preprocess_images<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  G4R_LAUNCH_CHECK("preprocess_images");
  return G4R_OK;
}
