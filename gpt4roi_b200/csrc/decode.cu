// decode.cu -- KV-cache append and single-query attention for the decode loop behind generate()
// (SURVEY.md 8(f1)): after the region-token prefill, every further step is plain LLaMA with one new
// token per sample (the vision/SPI branch is skipped, gpt4roi/models/spi_llava.py:47-48;
// llava/model/llava.py:263-283 prepare_inputs_for_generation; gpt4roi/app.py:293-300).
// Both kernels are HBM-bound: the step streams the whole KV cache of the batch once.
#include "common.cuh"

namespace g4r {

// rows [B*Ln, 3*HD] packed (q|k|v) -> caches [B, Lmax, HD] at positions pos0 .. pos0+Ln-1
__global__ void __launch_bounds__(256)
kv_append_bf16(const __nv_bfloat16* __restrict__ qkv, long long ld, __nv_bfloat16* __restrict__ kc,
               __nv_bfloat16* __restrict__ vc, int B, int Ln, int pos0, const int* __restrict__ pos_dev, int Lmax,
               int HD) {
  if (pos_dev) pos0 = *pos_dev;
  const int nvec = HD >> 3;
  const long long total = (long long)B * Ln * nvec * 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int which = (int)((i / nvec) & 1);  // 0 = k, 1 = v
    const long long r = i / (2LL * nvec);     // row in [0, B*Ln)
    const int t = (int)(r % Ln), b = (int)(r / Ln);
    const uint4 val = *reinterpret_cast<const uint4*>(qkv + r * ld + (long long)(1 + which) * HD + v * 8);
    __nv_bfloat16* dst = (which ? vc : kc) + ((long long)b * Lmax + pos0 + t) * HD + v * 8;
    *reinterpret_cast<uint4*>(dst) = val;
  }
}

// One CTA per (batch, head): out = softmax(q . K^T * scale) V over the first kv_len cached positions.
// 128 threads; phase 1: thread <-> key (fp32 dot over D), block max / sum; phase 2: thread <-> pair of
// output dims, coalesced reads of V rows.
template <int D>
__global__ void __launch_bounds__(128)
decode_attention_bf16(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ kc,
                      const __nv_bfloat16* __restrict__ vc, __nv_bfloat16* __restrict__ out, long long ldo,
                      int H, int kv_len, const int* __restrict__ pos_dev, int Lmax, float scale) {
  if (pos_dev) kv_len = *pos_dev + 1;
  extern __shared__ float sm[];  // [D] q | [kv_len] scores
  float* sq = sm;
  float* sc = sm + D;
  __shared__ float red[4];
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const int HD = H * D;
  const __nv_bfloat16* qrow = q + (long long)b * ldq + (long long)h * D;
  for (int i = tid; i < D; i += 128) sq[i] = __bfloat162float(qrow[i]) * scale;
  __syncthreads();
  const __nv_bfloat16* kb = kc + (long long)b * Lmax * HD + (long long)h * D;
  float mx = -INFINITY;
  for (int key = tid; key < kv_len; key += 128) {
    const uint4* kr = reinterpret_cast<const uint4*>(kb + (long long)key * HD);
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < D / 8; v++) {
      const uint4 raw = kr[v];
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float2 f = __bfloat1622float2(hh[j]);
        acc += f.x * sq[v * 8 + 2 * j] + f.y * sq[v * 8 + 2 * j + 1];
      }
    }
    sc[key] = acc;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int key = tid; key < kv_len; key += 128) {
    const float p = __expf(sc[key] - mx);
    sc[key] = p;
    sum += p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
  // phase 2: probabilities are cast to bf16 before the PV product like the reference (P.to(q.dtype))
  const __nv_bfloat16* vb = vc + (long long)b * Lmax * HD + (long long)h * D;
  constexpr int PAIRS = D / 2;  // 64 (D=128) or 32 (D=64) dim pairs; 128 threads split the keys 128/PAIRS ways
  const int pair = tid % PAIRS, part = tid / PAIRS, parts = 128 / PAIRS;
  float ax = 0.f, ay = 0.f;
  for (int key = part; key < kv_len; key += parts) {
    const float p = __bfloat162float(__float2bfloat16_rn(sc[key] * inv));
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vb + (long long)key * HD + 2 * pair));
    ax += p * f.x;
    ay += p * f.y;
  }
  __syncthreads();
  float* acc2 = sm;  // reuse: [parts][D]
  acc2[part * D + 2 * pair] = ax;
  acc2[part * D + 2 * pair + 1] = ay;
  __syncthreads();
  if (tid < PAIRS) {
    float x = 0.f, y = 0.f;
    for (int pp = 0; pp < parts; pp++) { x += acc2[pp * D + 2 * tid]; y += acc2[pp * D + 2 * tid + 1]; }
    *reinterpret_cast<__nv_bfloat162*>(out + (long long)b * ldo + (long long)h * D + 2 * tid) = __floats2bfloat162_rn(x, y);
  }
}

}  // namespace g4r

using namespace g4r;

extern "C" int g4r_kv_append_bf16(const void* qkv, long long ld, void* kcache, void* vcache, int B, int Ln,
                                  int pos0, const int* pos_dev, int Lmax, int HD, void* stream) {
  G4R_REQUIRE(qkv && kcache && vcache && B > 0 && Ln > 0 && pos0 >= 0 && pos0 + Ln <= Lmax && HD % 8 == 0 && ld % 8 == 0,
              "kv_append: bad arguments (pos0=%d Ln=%d Lmax=%d)", pos0, Ln, Lmax);
  const long long total = (long long)B * Ln * (HD / 8) * 2;
  long long grid = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (grid > cap) grid = cap;
  kv_append_bf16<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, ld, (__nv_bfloat16*)kcache,
                                                                  (__nv_bfloat16*)vcache, B, Ln, pos0, pos_dev, Lmax, HD);
  G4R_LAUNCH_CHECK("kv_append");
  return G4R_OK;
}

extern "C" int g4r_decode_attention_bf16(const void* q, long long ldq, const void* kcache, const void* vcache,
                                         void* out, long long ldo, int B, int H, int head_dim, int kv_len,
                                         const int* pos_dev, int Lmax, float scale, void* stream) {
  G4R_REQUIRE(q && kcache && vcache && out && B > 0 && H > 0 && kv_len > 0 && kv_len <= Lmax, "decode_attention: bad arguments");
  G4R_REQUIRE(head_dim == 64 || head_dim == 128, "decode_attention: head_dim %d", head_dim);
  const int smem = (head_dim + (kv_len > 2 * head_dim ? kv_len : 2 * head_dim) + 8) * 4 + (128 / (head_dim / 2)) * head_dim * 4;
  G4R_REQUIRE(smem <= 200 * 1024, "decode_attention: kv_len %d too long for the shared-memory score buffer", kv_len);
  dim3 grid(H, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (head_dim == 128) {
    static bool set = false;
    if (!set) { G4R_CUDA(cudaFuncSetAttribute(decode_attention_bf16<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); set = true; }
    decode_attention_bf16<128><<<grid, 128, smem, st>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)kcache,
                                                         (const __nv_bfloat16*)vcache, (__nv_bfloat16*)out, ldo, H, kv_len, pos_dev, Lmax, scale);
  } else {
    static bool set = false;
    if (!set) { G4R_CUDA(cudaFuncSetAttribute(decode_attention_bf16<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); set = true; }
    decode_attention_bf16<64><<<grid, 128, smem, st>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)kcache,
                                                        (const __nv_bfloat16*)vcache, (__nv_bfloat16*)out, ldo, H, kv_len, pos_dev, Lmax, scale);
  }
  G4R_LAUNCH_CHECK("decode_attention");
  return G4R_OK;
}
