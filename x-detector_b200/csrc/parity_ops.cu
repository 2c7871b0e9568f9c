// fp32-accurate ("bf16x3") PARITY MODE helpers.
//
// The throughput path runs the convolutions on bf16 operands; against the reference's fp32 graph that
// bounds agreement at ~1e-2 per stage.  north_star asks for fp32 box/score deltas within 1e-4, so the same
// tcgen05 kernel is also driven in a split-operand mode: every fp32 value v is written as three bf16 pieces
//     hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid)          (v == hi + mid + lo up to 2^-24 |v|)
// and a convolution over C channels becomes ONE convolution over 6*C channels whose operand pairs are the six
// largest cross products (all others are below 2^-24 of the result):
//     activations  [ mid | lo | hi | mid | hi | hi ]      (channel blocks of C)
//     weights      [ mid | hi | lo | hi  | mid| hi ]
// The products are exact in the fp32 accumulator, so the result carries fp32-level error (measured against
// the fp32 CPU oracle in tests/test_parity_mode_gpu.py).  No new tensor-core code: xdet_conv2d_bf16 runs
// unchanged with fp32 output; the kernels below are the elementwise glue that stays in fp32 between layers.
// Nothing here is on the throughput path (bench.py never enables it).
#include <cuda_bf16.h>

#include <cfloat>

#include "common.cuh"

namespace xdet {
namespace {

// src element (n,y,x,c) at src + n*sn + y*sy + x*sx + c*sc (fp32, any layout); dst [N,H,W,out_cs] bf16 with
// the six blocks of C channels described above and a zero tail up to out_cs.
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, long long sn, long long sy,
                                                     long long sx, long long sc, int H, int W, int C,
                                                     __nv_bfloat16* __restrict__ dst, int out_cs, long long total) {
  const long long step = (long long)gridDim.x * blockDim.x;
  const int tail = out_cs - 6 * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int c = (int)(e % C);
    const long long pix = e / C;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    const float v = __ldg(src + n * sn + y * sy + x * sx + c * sc);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = __fsub_rn(v, __bfloat162float(hi));  // exact
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const float r2 = __fsub_rn(r1, __bfloat162float(mid));  // exact
    const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
    __nv_bfloat16* d = dst + pix * out_cs + c;
    d[0] = mid;
    d[C] = lo;
    d[2 * C] = hi;
    d[3 * C] = mid;
    d[4 * C] = hi;
    d[5 * C] = hi;
    if (c == 0)
      for (int t = 0; t < tail; ++t) d[6 * C + t] = __float2bfloat16_rn(0.f);
  }
}

// v = x (+ residual) (ReLU if relu) ; out = v (optional) ; out2 = ReLU(v*scale2[c] + bias2[c]) (optional).
// The fp32 form of the conv epilogue's residual / second-output stages (conv_gemm.cu) and of affine_relu.
__global__ void __launch_bounds__(256) f32_post_kernel(const float* __restrict__ x, const float* __restrict__ residual,
                                                       int relu, float* __restrict__ out,
                                                       const float* __restrict__ scale2,
                                                       const float* __restrict__ bias2, int relu2,
                                                       float* __restrict__ out2, int C, long long total) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int c = (int)(e % C);
    float v = x[e];
    if (residual) v = __fadd_rn(v, residual[e]);
    if (relu) v = fmaxf(v, 0.f);
    if (out) out[e] = v;
    if (out2) {
      float w = __fadd_rn(__fmul_rn(v, scale2[c]), bias2[c]);
      if (relu2) w = fmaxf(w, 0.f);
      out2[e] = w;
    }
  }
}

// tf.layers.max_pooling2d(3, 2, 'SAME') on NHWC fp32 (+ residual), optional second output relu(y*scale2+bias2).
__global__ void __launch_bounds__(256) maxpool3x3s2_f32_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                               float* __restrict__ dst2,
                                                               const float* __restrict__ scale2,
                                                               const float* __restrict__ bias2,
                                                               const float* __restrict__ residual, int N, int H, int W,
                                                               int C, int Ho, int Wo, int pad_top, int pad_left,
                                                               long long total) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int c = (int)(e % C);
    const long long pix = e / C;
    const int xo = (int)(pix % Wo);
    const int yo = (int)((pix / Wo) % Ho);
    const long long n = pix / ((long long)Wo * Ho);
    float m = -FLT_MAX;
    for (int kh = 0; kh < 3; ++kh) {
      const int yi = yo * 2 + kh - pad_top;
      if (yi < 0 || yi >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int xi = xo * 2 + kw - pad_left;
        if (xi < 0 || xi >= W) continue;
        m = fmaxf(m, __ldg(src + ((n * H + yi) * W + xi) * C + c));
      }
    }
    if (residual) m = __fadd_rn(m, residual[e]);
    dst[e] = m;
    if (dst2) dst2[e] = fmaxf(__fadd_rn(__fmul_rn(m, scale2[c]), bias2[c]), 0.f);
  }
}

// Depthwise 3x3 'SAME' stride-1 convolution in fp32 (depth multiplier 1, dilation 1 or 2), taps summed in
// (kh, kw) order; relu_in applies ReLU while loading.
__global__ void __launch_bounds__(256) depthwise3x3_f32_kernel(const float* __restrict__ src,
                                                               const float* __restrict__ w9c, float* __restrict__ dst,
                                                               int N, int H, int W, int C, int dil, int relu_in,
                                                               long long total) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int c = (int)(e % C);
    const long long pix = e / C;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    float acc = 0.f;
    for (int kh = 0; kh < 3; ++kh) {
      const int yi = y + (kh - 1) * dil;
      if (yi < 0 || yi >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int xi = x + (kw - 1) * dil;
        if (xi < 0 || xi >= W) continue;
        float v = __ldg(src + ((n * H + yi) * W + xi) * C + c);
        if (relu_in) v = fmaxf(v, 0.f);
        acc = __fmaf_rn(v, __ldg(w9c + (kh * 3 + kw) * C + c), acc);
      }
    }
    dst[e] = acc;
  }
}

unsigned grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace
}  // namespace xdet

using namespace xdet;

extern "C" int xdet_split3_bf16(const float* d_src, long long sn, long long sy, long long sx, long long sc, int N,
                                int H, int W, int C, void* d_dst, int out_cs, void* stream) {
  if (N < 0 || H < 0 || W < 0 || C <= 0) return fail(XDET_EINVAL, "split3: bad shape");
  if (out_cs < 6 * C || out_cs % 8 != 0) return fail(XDET_EINVAL, "split3: out_cs (%d) must be >= 6*C and % 8 == 0", out_cs);
  const long long total = (long long)N * H * W * C;
  if (total == 0) return XDET_OK;
  split3_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(d_src, sn, sy, sx, sc, H, W, C,
                                                                   reinterpret_cast<__nv_bfloat16*>(d_dst), out_cs, total);
  return after_launch("split3_kernel");
}

extern "C" int xdet_f32_post(const float* d_x, const float* d_residual, int relu, float* d_out, const float* d_scale2,
                             const float* d_bias2, int relu2, float* d_out2, long long rows, int C, void* stream) {
  if (rows < 0 || C <= 0) return fail(XDET_EINVAL, "f32_post: bad shape");
  if (d_out2 && (!d_scale2 || !d_bias2)) return fail(XDET_EINVAL, "f32_post: out2 needs scale2 and bias2");
  const long long total = rows * C;
  if (total == 0) return XDET_OK;
  f32_post_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(d_x, d_residual, relu, d_out, d_scale2, d_bias2,
                                                                     relu2, d_out2, C, total);
  return after_launch("f32_post_kernel");
}

extern "C" int xdet_maxpool3x3s2_f32(const float* d_src, float* d_dst, float* d_dst2, const float* d_scale2,
                                     const float* d_bias2, const float* d_residual, int N, int H, int W, int C, int Ho,
                                     int Wo, int pad_top, int pad_left, void* stream) {
  if (N < 0 || H <= 0 || W <= 0 || C <= 0) return fail(XDET_EINVAL, "maxpool_f32: bad shape");
  const long long total = (long long)N * Ho * Wo * C;
  if (total == 0) return XDET_OK;
  maxpool3x3s2_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(d_src, d_dst, d_dst2, d_scale2, d_bias2,
                                                                             d_residual, N, H, W, C, Ho, Wo, pad_top,
                                                                             pad_left, total);
  return after_launch("maxpool3x3s2_f32_kernel");
}

extern "C" int xdet_depthwise3x3_f32(const float* d_src, const float* d_weights, float* d_dst, int N, int H, int W,
                                     int C, int dilation, int relu_in, void* stream) {
  if (N < 0 || H <= 0 || W <= 0 || C <= 0 || dilation < 1) return fail(XDET_EINVAL, "depthwise_f32: bad shape");
  const long long total = (long long)N * H * W * C;
  if (total == 0) return XDET_OK;
  depthwise3x3_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(d_src, d_weights, d_dst, N, H, W, C,
                                                                             dilation, relu_in, total);
  return after_launch("depthwise3x3_f32_kernel");
}
