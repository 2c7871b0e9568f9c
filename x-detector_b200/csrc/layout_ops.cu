// Bandwidth kernels around the tensor-core convolutions (all HBM-bound, no tensor cores):
//   * patch gather ("im2col") for the few STRIDED convolutions of the ResNet-v2 stem/stage heads
//     (7x7/s2 initial conv, 3x3/s2 and 1x1/s2 of the first block of block_layer2/3;
//     net/resnet_v2.py:62-100,320-343): writes [N,Ho,Wo,KH*KW*C] bf16 so that the strided conv
//     becomes a plain GEMM for conv_gemm.cu.  `fixed_padding` (resnet_v2.py:62-86) is the zero
//     border the gather produces on the fly.
//   * 3x3/s2 'SAME' max pooling (resnet_v2.py:326-328) with an optional fused second output
//     relu(x*scale+bias): the pre-activation BN+ReLU of the first bottleneck (resnet_v2.py:163-164).
//   * per-channel affine + ReLU (inference batch_norm_relu, resnet_v2.py:41-50) for the places where
//     one tensor feeds two different batch-norms.
//   * fp32 -> bf16 row repack with a padded row pitch (PsRoIAlign output -> dense-layer operand).
#include <cuda_bf16.h>

#include <cfloat>

#include "common.cuh"

namespace xdet {
namespace {

// src_kind: 0 = NHWC bf16 (pixel pitch in_cs), 1 = NCHW fp32
template <int SRC>
__global__ void __launch_bounds__(256) im2col_kernel(const void* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                     int N, int H, int W, int C, int in_cs, int Ho, int Wo, int KH,
                                                     int KW, int stride, int pad_top, int pad_left, int out_cs,
                                                     long long total) {
  const int K = KH * KW * C;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int k = (int)(e % out_cs);
    const long long pix = e / out_cs;
    float v = 0.f;
    if (k < K) {
      const int c = k % C, tap = k / C;
      const int kw = tap % KW, kh = tap / KW;
      const int xo = (int)(pix % Wo);
      const int yo = (int)((pix / Wo) % Ho);
      const int n = (int)(pix / ((long long)Wo * Ho));
      const int yi = yo * stride + kh - pad_top, xi = xo * stride + kw - pad_left;
      if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
        if (SRC == 0)
          v = __bfloat162float(
              reinterpret_cast<const __nv_bfloat16*>(src)[(((long long)n * H + yi) * W + xi) * in_cs + c]);
        else
          v = __ldg(reinterpret_cast<const float*>(src) + (((long long)n * C + c) * H + yi) * W + xi);
      }
    }
    dst[e] = __float2bfloat16_rn(v);
  }
}

// 16-byte vector version for NHWC bf16 with C % 8 == 0 and out_cs == KH*KW*C.
__global__ void __launch_bounds__(256) im2col_vec8_kernel(const __nv_bfloat16* __restrict__ src,
                                                          __nv_bfloat16* __restrict__ dst, int N, int H, int W, int C,
                                                          int in_cs, int Ho, int Wo, int KH, int KW, int stride,
                                                          int pad_top, int pad_left, long long total_vec) {
  const int C8 = C / 8, K8 = KH * KW * C8;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total_vec; e += step) {
    const int k8 = (int)(e % K8);
    const long long pix = e / K8;
    const int c8 = k8 % C8, tap = k8 / C8;
    const int kw = tap % KW, kh = tap / KW;
    const int xo = (int)(pix % Wo);
    const int yo = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((long long)Wo * Ho));
    const int yi = yo * stride + kh - pad_top, xi = xo * stride + kw - pad_left;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (yi >= 0 && yi < H && xi >= 0 && xi < W)
      v = __ldg(reinterpret_cast<const uint4*>(src + (((long long)n * H + yi) * W + xi) * in_cs) + c8);
    reinterpret_cast<uint4*>(dst)[e] = v;
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ src,
                                                           __nv_bfloat16* __restrict__ dst,
                                                           __nv_bfloat16* __restrict__ dst2,
                                                           const float* __restrict__ scale2,
                                                           const float* __restrict__ bias2,
                                                           const __nv_bfloat16* __restrict__ residual,
                                                           unsigned char* __restrict__ argmax, int N, int H,
                                                           int W, int C, int Ho, int Wo, int pad_top, int pad_left,
                                                           long long total_vec) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  const int C8 = C / 8;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total_vec; e += step) {
    const int c8 = (int)(e % C8);
    const long long pix = e / C8;
    const int xo = (int)(pix % Wo);
    const int yo = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((long long)Wo * Ho));
    // all nine window loads first (independent, in flight together), then the reduction
    uint4 win[9];
    bool ok[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yi = yo * 2 + kh - pad_top;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int xi = xo * 2 + kw - pad_left;
        ok[kh * 3 + kw] = yi >= 0 && yi < H && xi >= 0 && xi < W;
        win[kh * 3 + kw] = ok[kh * 3 + kw]
                               ? __ldg(reinterpret_cast<const uint4*>(src + (((long long)n * H + yi) * W + xi) * C) + c8)
                               : make_uint4(0, 0, 0, 0);
      }
    }
    float m[8];
    unsigned am[8];  // window position (kh*3 + kw) of the FIRST maximum: where the gradient goes
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m[j] = -FLT_MAX;
      am[j] = 0;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (!ok[k]) continue;
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&win[k]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        if (f.x > m[2 * q]) {
          m[2 * q] = f.x;
          am[2 * q] = k;
        }
        if (f.y > m[2 * q + 1]) {
          m[2 * q + 1] = f.y;
          am[2 * q + 1] = k;
        }
      }
    }
    if (argmax) {
      uint2 a;
      a.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
      a.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
      reinterpret_cast<uint2*>(argmax)[e] = a;
    }
    if (residual) {  // Xception entry flow: tf.add(max_pool(x), residual) (net/xception_body.py:283-289)
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(residual) + e);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        m[2 * q] += f.x;
        m[2 * q + 1] += f.y;
      }
    }
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) ho[q] = __floats2bfloat162_rn(m[2 * q], m[2 * q + 1]);
    reinterpret_cast<uint4*>(dst)[e] = o;
    if (dst2) {
      uint4 o2;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c8 * 8 + 2 * q;
        // the second output normalises the ROUNDED pooled value (what the unfused graph would read back)
        const float2 r = __bfloat1622float2(ho[q]);
        h2[q] = __floats2bfloat162_rn(fmaxf(fmaf(r.x, __ldg(scale2 + c), __ldg(bias2 + c)), 0.f),
                                      fmaxf(fmaf(r.y, __ldg(scale2 + c + 1), __ldg(bias2 + c + 1)), 0.f));
      }
      reinterpret_cast<uint4*>(dst2)[e] = o2;
    }
  }
}

// y = x*scale[c] + bias[c] (ReLU): a thread owns 8 channels (constants in registers) and strides over the rows;
// 256 threads = cgb channel groups x 256/cgb row lanes.
__global__ void __launch_bounds__(256) affine_relu_kernel(const __nv_bfloat16* __restrict__ src,
                                                          __nv_bfloat16* __restrict__ dst,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ bias, int C, int relu, int cgb,
                                                          long long rows) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  const int rl = 256 / cgb;
  const int cg = threadIdx.x % cgb, lane_row = threadIdx.x / cgb;
  const int c0 = (blockIdx.x * cgb + cg) * 8;
  if (c0 >= C) return;
  float sc[8], bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[c0 + j];
    bi[j] = bias[c0 + j];
  }
#pragma unroll 4
  for (long long r = (long long)blockIdx.y * rl + lane_row; r < rows; r += (long long)gridDim.y * rl) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + r * C + c0));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = __bfloat1622float2(h[q]);
      float a = fmaf(f.x, sc[2 * q], bi[2 * q]);
      float b = fmaf(f.y, sc[2 * q + 1], bi[2 * q + 1]);
      if (relu) {
        a = fmaxf(a, 0.f);
        b = fmaxf(b, 0.f);
      }
      ho[q] = __floats2bfloat162_rn(a, b);
    }
    *reinterpret_cast<uint4*>(dst + r * C + c0) = o;
  }
}

__global__ void __launch_bounds__(256) f32_to_bf16_rows_kernel(const float* __restrict__ src,
                                                               __nv_bfloat16* __restrict__ dst, long long rows,
                                                               int cols, int dst_pitch) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  const long long total = rows * dst_pitch;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int c = (int)(e % dst_pitch);
    const long long r = e / dst_pitch;
    dst[e] = __float2bfloat16_rn(c < cols ? __ldg(src + r * cols + c) : 0.f);
  }
}

// Depthwise 3x3 'SAME' stride-1 convolution (dilation 1 or 2), depth multiplier 1: the first half of
// tf.layers.separable_conv2d (net/xception_body.py:224-233,264-272,351-376).  NHWC bf16 in/out, fp32 taps and
// accumulation, optional ReLU on the input (the tf.nn.relu in front of relu_separable_bn_block, :223).
// HBM/L2-bandwidth work: one thread = ONE channel pair (4 bytes) of PX adjacent output pixels of a row, so a warp
// reads/writes 128 contiguous bytes per pixel; the whole 3 x (PX + 2*DIL) input window is fetched up front
// (30-36 independent 4-byte loads in flight per thread), then consumed from registers.
template <int PX, int DIL>
__global__ void __launch_bounds__(256, 3) depthwise3x3_kernel(const __nv_bfloat16* __restrict__ src,
                                                           const float* __restrict__ w /* [9][C] */,
                                                           __nv_bfloat16* __restrict__ dst, int N, int H, int W, int C,
                                                           int relu_in, long long total) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  constexpr int WIN = PX + 2 * DIL;
  const int C2 = C / 2;
  const int WX = (W + PX - 1) / PX;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int c2 = (int)(e % C2);
    long long t = e / C2;
    const int xg = (int)(t % WX);
    t /= WX;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    const int x0 = xg * PX;
    const uint32_t* base = reinterpret_cast<const uint32_t*>(src) + c2;
    uint32_t in[3][WIN];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yi = y + (kh - 1) * DIL;
      const bool row_ok = yi >= 0 && yi < H;
      const long long roff = ((long long)n * H + (row_ok ? yi : 0)) * W;
#pragma unroll
      for (int i = 0; i < WIN; ++i) {
        const int xi = x0 + i - DIL;
        in[kh][i] = (row_ok && xi >= 0 && xi < W) ? __ldg(base + (roff + xi) * C2) : 0u;
      }
    }
    float2 wt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wt[k] = __ldg(reinterpret_cast<const float2*>(w + (long long)k * C) + c2);
    float2 acc[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[p] = make_float2(0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int i = 0; i < WIN; ++i) {
        float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&in[kh][i]));
        if (relu_in) {
          v.x = fmaxf(v.x, 0.f);
          v.y = fmaxf(v.y, 0.f);
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int p = i - kw * DIL;  // output pixel fed by this input through tap kw
          if (p >= 0 && p < PX) {
            acc[p].x = fmaf(v.x, wt[kh * 3 + kw].x, acc[p].x);
            acc[p].y = fmaf(v.y, wt[kh * 3 + kw].y, acc[p].y);
          }
        }
      }
    }
    uint32_t* orow = reinterpret_cast<uint32_t*>(dst) + (((long long)n * H + y) * W + x0) * C2 + c2;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (x0 + p < W) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(acc[p].x, acc[p].y);
        orow[(long long)p * C2] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
  }
}

// Depthwise 3x3, dilation 1, "rolling rows" form: one thread = 4 channels (8 bytes) of PX adjacent output columns and
// walks DOWN a strip of YS rows with the 3 x (PX+2) input window in fp32 registers (rotated by renaming), so every
// input pixel is loaded and widened once per strip (0.06 load instructions per output element; the one-row-per-
// thread kernel needs 1.9).  What bounds such a kernel is memory-level parallelism: HBM latency x 6.5 TB/s needs
// ~35 KB in flight per SM.  The next D rows of every thread are therefore streamed by cp.async into a thread-private
// shared-memory ring (no registers held by loads in flight: D * 32 B * 512 threads = 64 KB per SM outstanding) and
// picked up with one 8-byte LDS per pixel.  Measured at 32x100x100x728: 1.7 TB/s (no prefetch) -> 2.9 TB/s (two rows
// ahead in registers) -> see profiles/ for this version.
// packed fp32 pairs (Blackwell FFMA2: two fp32 FMAs per issued instruction)
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int PX, int D>
__global__ void __launch_bounds__(256, 2) depthwise3x3_rows_kernel(const __nv_bfloat16* __restrict__ src,
                                                                   const float* __restrict__ w /* [9][C] */,
                                                                   __nv_bfloat16* __restrict__ dst, int N, int H, int W,
                                                                   int C, int relu_in, int YS, long long total) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  constexpr int WIN = PX + 2;
  constexpr int RS = D + 1;  // ring slots: D rows in flight + the one being consumed
  __shared__ uint2 ring[RS * WIN * 256];
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) e = total - 1;  // surplus threads of the last CTA redo its last item (identical values)
  const int C4 = C / 4;
  const int WX = (W + PX - 1) / PX;
  const int HY = (H + YS - 1) / YS;
  const int c4 = (int)(e % C4);
  long long t = e / C4;
  const int xg = (int)(t % WX);
  t /= WX;
  const int ys = (int)(t % HY);
  const int n = (int)(t / HY);
  const int x0 = xg * PX, y0 = ys * YS, y1 = min(y0 + YS, H);
  f32x2_t wt[9][2];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(w + (long long)k * C) + c4);
    wt[k][0] = pack2(v.x, v.y);
    wt[k][1] = pack2(v.z, v.w);
  }
  // Addressing is incremental: column offsets / validity are fixed for the strip, the row pointers advance by one
  // row pitch per step, the ring slots by one slot pitch (no multiplies, divisions or 64-bit products in the loop).
  const long long row_pitch = (long long)W * C4;  // uint2 units
  const uint2* img = reinterpret_cast<const uint2*>(src) + (long long)n * H * row_pitch + c4;
  int col_off[WIN], col_sz[WIN];
  bool col_ok[WIN];
#pragma unroll
  for (int i = 0; i < WIN; ++i) {
    const int xi = x0 + i - 1;
    col_ok[i] = xi >= 0 && xi < W;
    col_off[i] = (col_ok[i] ? xi : 0) * C4;
    col_sz[i] = col_ok[i] ? 8 : 0;  // cp.async source size: 0 = zero fill
  }
  const uint32_t ring_sa = (uint32_t)__cvta_generic_to_shared(ring) + threadIdx.x * 8u;
  constexpr uint32_t kSlotBytes = WIN * 256 * 8;
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  // queue input row yi (pointer `rowp`) into the ring slot at shared address `sa`; zero-filled where the row /
  // column is outside the image or below the strip's last halo row
  auto queue_row = [&](int yi, const uint2* rowp, uint32_t sa) {
    const bool row_ok = yi >= 0 && yi < H && yi <= y1;
    const uint2* rp = row_ok ? rowp : img;
    const int rmask = row_ok ? -1 : 0;
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa + (uint32_t)(i * 256 * 8)), "l"(rp + col_off[i]),
                   "r"(col_sz[i] & rmask)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto widen = [&](uint2 v, f32x2_t (&r)[2]) {
    if (relu_in) {  // on the packed pairs: one instruction per two values
      const __nv_bfloat162 a = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&v.x), zero2);
      const __nv_bfloat162 b2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&v.y), zero2);
      v.x = *reinterpret_cast<const uint32_t*>(&a);
      v.y = *reinterpret_cast<const uint32_t*>(&b2);
    }
    r[0] = pack2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
    r[1] = pack2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
  };
  auto take_row = [&](uint32_t sa, f32x2_t (&r)[WIN][2]) {  // oldest queued row -> fp32 registers
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      uint2 v;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sa + (uint32_t)(i * 256 * 8)));
      widen(v, r[i]);
    }
  };
  auto load_row = [&](int yi, f32x2_t (&r)[WIN][2]) {  // direct (the first two rows of the strip)
    const bool row_ok = yi >= 0 && yi < H;
    const uint2* rp = img + (long long)(row_ok ? yi : 0) * row_pitch;
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      const uint2 v = (row_ok && col_ok[i]) ? __ldg(rp + col_off[i]) : make_uint2(0u, 0u);
      widen(v, r[i]);
    }
  };
  uint2* orow = reinterpret_cast<uint2*>(dst) + ((long long)n * H + y0) * row_pitch + (long long)x0 * C4 + c4;
  bool out_ok[PX];
#pragma unroll
  for (int p = 0; p < PX; ++p) out_ok[p] = x0 + p < W;
  auto emit = [&](const f32x2_t (&ra)[WIN][2], const f32x2_t (&rb)[WIN][2], const f32x2_t (&rc)[WIN][2]) {
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      f32x2_t a2[2] = {0ull, 0ull};  // (+0, +0)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int h = 0; h < 2; ++h) a2[h] = ffma2(ra[p + kw][h], wt[kw][h], a2[h]);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int h = 0; h < 2; ++h) a2[h] = ffma2(rb[p + kw][h], wt[3 + kw][h], a2[h]);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int h = 0; h < 2; ++h) a2[h] = ffma2(rc[p + kw][h], wt[6 + kw][h], a2[h]);
      if (out_ok[p]) {
        float acc[4];
        unpack2(a2[0], acc[0], acc[1]);
        unpack2(a2[1], acc[2], acc[3]);
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
        orow[p * C4] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
    }
    orow += row_pitch;
  };
  // rows y0+1 .. y0+D go into slots 0 .. D-1; afterwards the slot consumed one step ago is refilled
  const uint2* qrow = img + (long long)(y0 + 1) * row_pitch;  // pointer of the next row to queue
  int qy = y0 + 1;
  uint32_t q_sa = ring_sa;                                      // its slot
#pragma unroll
  for (int j = 0; j < D; ++j) {
    queue_row(qy, qrow, q_sa);
    ++qy;
    qrow += row_pitch;
    q_sa += kSlotBytes;
  }
  // q_sa now points at slot D (= RS-1): the free one
  uint32_t t_sa = ring_sa;  // slot of the oldest queued row
  f32x2_t r0[WIN][2], r1[WIN][2], r2[WIN][2];
  load_row(y0 - 1, r0);
  load_row(y0, r1);
  int y = y0;
  auto step = [&](f32x2_t (&ra)[WIN][2], f32x2_t (&rb)[WIN][2], f32x2_t (&rc)[WIN][2]) {
    take_row(t_sa, rc);  // row y+1
    t_sa = (t_sa == ring_sa + (RS - 1) * kSlotBytes) ? ring_sa : t_sa + kSlotBytes;
    queue_row(qy, qrow, q_sa);
    ++qy;
    qrow += row_pitch;
    q_sa = (q_sa == ring_sa + (RS - 1) * kSlotBytes) ? ring_sa : q_sa + kSlotBytes;
    emit(ra, rb, rc);
    ++y;
  };
  while (y < y1) {  // three output rows per trip: the fp32 window rotates by renaming, not by moving registers
    step(r0, r1, r2);
    if (y >= y1) break;
    step(r1, r2, r0);
    if (y >= y1) break;
    step(r2, r0, r1);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// NCHW fp32 image -> row-padded NHWC bf16 with the channels zero-padded to `cs` (8): the input layout of the
// fold_w convolution (conv_gemm.cu) that runs the 7x7/s2 stem.  dst[n][y][pad_left + x][c], zeros elsewhere.
__global__ void __launch_bounds__(256) image_to_nhwc8_kernel(const float* __restrict__ src,
                                                             __nv_bfloat16* __restrict__ dst, int N, int C, int H,
                                                             int W, int Wp, int pad_left, long long total) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int xp = (int)(e % Wp);
    const long long row = e / Wp;  // n*H + y
    const int y = (int)(row % H);
    const int n = (int)(row / H);
    const int x = xp - pad_left;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = 0.f;
    if (x >= 0 && x < W) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < C) v[c] = __ldg(src + (((long long)n * C + c) * H + y) * W + x);
    }
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) ho[q] = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
    reinterpret_cast<uint4*>(dst)[e] = o;
  }
}

unsigned grid_for(long long total, int threads = 256) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace
}  // namespace xdet

using namespace xdet;

extern "C" int xdet_im2col_bf16(const void* d_src, int src_is_nchw_f32, void* d_dst, int N, int H, int W, int C,
                                int in_cs, int KH, int KW, int stride, int pad_top, int pad_left, int Ho, int Wo,
                                int out_cs, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || KH <= 0 || KW <= 0 || stride <= 0 || Ho <= 0 || Wo <= 0)
    return fail(XDET_EINVAL, "im2col: non-positive dimension");
  if (out_cs < KH * KW * C) return fail(XDET_EINVAL, "im2col: out_cs (%d) < KH*KW*C (%d)", out_cs, KH * KW * C);
  cudaStream_t st = (cudaStream_t)stream;
  const long long pixels = (long long)N * Ho * Wo;
  if (!src_is_nchw_f32 && C % 8 == 0 && in_cs % 8 == 0 && out_cs == KH * KW * C) {
    const long long tv = pixels * (out_cs / 8);
    im2col_vec8_kernel<<<grid_for(tv), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(d_src),
                                                     reinterpret_cast<__nv_bfloat16*>(d_dst), N, H, W, C, in_cs, Ho,
                                                     Wo, KH, KW, stride, pad_top, pad_left, tv);
    return after_launch("im2col_vec8_kernel");
  }
  const long long total = pixels * out_cs;
  if (src_is_nchw_f32)
    im2col_kernel<1><<<grid_for(total), 256, 0, st>>>(d_src, reinterpret_cast<__nv_bfloat16*>(d_dst), N, H, W, C,
                                                      in_cs, Ho, Wo, KH, KW, stride, pad_top, pad_left, out_cs, total);
  else
    im2col_kernel<0><<<grid_for(total), 256, 0, st>>>(d_src, reinterpret_cast<__nv_bfloat16*>(d_dst), N, H, W, C,
                                                      in_cs, Ho, Wo, KH, KW, stride, pad_top, pad_left, out_cs, total);
  return after_launch("im2col_kernel");
}

extern "C" int xdet_maxpool3x3s2_bf16(const void* d_src, void* d_dst, void* d_dst2, const float* d_scale2,
                                      const float* d_bias2, int N, int H, int W, int C, int Ho, int Wo, int pad_top,
                                      int pad_left, void* stream) {
  return xdet_maxpool3x3s2_add_bf16(d_src, d_dst, d_dst2, d_scale2, d_bias2, nullptr, N, H, W, C, Ho, Wo, pad_top,
                                    pad_left, stream);
}

extern "C" int xdet_maxpool3x3s2_argmax_bf16(const void* d_src, void* d_dst, void* d_argmax, int N, int H, int W, int C,
                                             int Ho, int Wo, int pad_top, int pad_left, void* stream) {
  if (C % 8 != 0) return fail(XDET_EINVAL, "maxpool: C (%d) must be a multiple of 8", C);
  const long long tv = (long long)N * Ho * Wo * (C / 8);
  if (tv <= 0) return XDET_OK;
  maxpool3x3s2_kernel<<<grid_for(tv), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_src), reinterpret_cast<__nv_bfloat16*>(d_dst), nullptr, nullptr, nullptr,
      nullptr, reinterpret_cast<unsigned char*>(d_argmax), N, H, W, C, Ho, Wo, pad_top, pad_left, tv);
  return after_launch("maxpool3x3s2_kernel");
}

extern "C" int xdet_maxpool3x3s2_add_bf16(const void* d_src, void* d_dst, void* d_dst2, const float* d_scale2,
                                          const float* d_bias2, const void* d_residual, int N, int H, int W, int C,
                                          int Ho, int Wo, int pad_top, int pad_left, void* stream) {
  if (C % 8 != 0) return fail(XDET_EINVAL, "maxpool: C (%d) must be a multiple of 8", C);
  if (d_dst2 && (!d_scale2 || !d_bias2)) return fail(XDET_EINVAL, "maxpool: second output needs scale2/bias2");
  const long long tv = (long long)N * Ho * Wo * (C / 8);
  if (tv <= 0) return XDET_OK;
  maxpool3x3s2_kernel<<<grid_for(tv), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_src), reinterpret_cast<__nv_bfloat16*>(d_dst),
      reinterpret_cast<__nv_bfloat16*>(d_dst2), d_scale2, d_bias2, reinterpret_cast<const __nv_bfloat16*>(d_residual),
      nullptr, N, H, W, C, Ho, Wo, pad_top, pad_left, tv);
  return after_launch("maxpool3x3s2_kernel");
}

extern "C" int xdet_affine_relu_bf16(const void* d_src, void* d_dst, const float* d_scale, const float* d_bias,
                                     long long pixels, int C, int relu, void* stream) {
  if (C % 8 != 0) return fail(XDET_EINVAL, "affine_relu: C (%d) must be a multiple of 8", C);
  if (pixels <= 0) return XDET_OK;
  int cgb = 1;
  while (cgb * 2 <= 32 && cgb * 2 <= C / 8) cgb *= 2;
  const int gx = (C / 8 + cgb - 1) / cgb, rl = 256 / cgb;
  long long slabs = (pixels + (long long)rl * 4 - 1) / ((long long)rl * 4);
  const long long cap = (long long)kNumSMs * 8 / gx < 1 ? 1 : (long long)kNumSMs * 8 / gx;
  if (slabs > cap) slabs = cap;
  affine_relu_kernel<<<dim3((unsigned)gx, (unsigned)slabs), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_src), reinterpret_cast<__nv_bfloat16*>(d_dst), d_scale, d_bias, C, relu, cgb,
      pixels);
  return after_launch("affine_relu_kernel");
}

extern "C" int xdet_f32_to_bf16_rows(const float* d_src, void* d_dst, long long rows, int cols, int dst_pitch,
                                     void* stream) {
  if (dst_pitch < cols) return fail(XDET_EINVAL, "f32_to_bf16_rows: dst_pitch < cols");
  if (rows <= 0) return XDET_OK;
  f32_to_bf16_rows_kernel<<<grid_for(rows * dst_pitch), 256, 0, (cudaStream_t)stream>>>(
      d_src, reinterpret_cast<__nv_bfloat16*>(d_dst), rows, cols, dst_pitch);
  return after_launch("f32_to_bf16_rows_kernel");
}

extern "C" int xdet_image_to_nhwc8_bf16(const float* d_src, void* d_dst, int N, int C, int H, int W, int Wp,
                                        int pad_left, void* stream) {
  if (N <= 0 || C <= 0 || C > 8 || H <= 0 || W <= 0) return fail(XDET_EINVAL, "image_to_nhwc8: need 1 <= C <= 8 and positive sizes");
  if (pad_left < 0 || Wp < W + pad_left) return fail(XDET_EINVAL, "image_to_nhwc8: Wp (%d) < W + pad_left", Wp);
  const long long total = (long long)N * H * Wp;
  image_to_nhwc8_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      d_src, reinterpret_cast<__nv_bfloat16*>(d_dst), N, C, H, W, Wp, pad_left, total);
  return after_launch("image_to_nhwc8_kernel");
}

static int g_dw_rows = 1;
extern "C" void xdet_set_depthwise_rows(int enabled) { g_dw_rows = enabled ? 1 : 0; }

extern "C" int xdet_depthwise3x3_bf16(const void* d_src, const float* d_weights, void* d_dst, int N, int H, int W, int C,
                                      int dilation, int relu_in, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0) return fail(XDET_EINVAL, "depthwise3x3: non-positive dimension");
  if (C % 8 != 0) return fail(XDET_EINVAL, "depthwise3x3: C (%d) must be a multiple of 8", C);
  if (dilation != 1 && dilation != 2) return fail(XDET_EINVAL, "depthwise3x3: dilation must be 1 or 2");
  constexpr int PX = 8;
  const long long total = (long long)N * H * ((W + PX - 1) / PX) * (C / 2);
  const unsigned grid = grid_for(total);
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(d_src);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(d_dst);
  if (dilation == 1 && g_dw_rows) {
    constexpr int PXR = 2;
    // strip height: as tall as keeps >= ~4 CTAs per SM in flight
    int YS = 16;
    while (YS > 4 && (long long)N * ((H + YS - 1) / YS) * ((W + PXR - 1) / PXR) * (C / 4) < 4ll * kNumSMs * 256) YS /= 2;
    const long long tot = (long long)N * ((H + YS - 1) / YS) * ((W + PXR - 1) / PXR) * (C / 4);
    depthwise3x3_rows_kernel<PXR, 4><<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, d_weights, dst, N, H, W, C, relu_in, YS, tot);
    return after_launch("depthwise3x3_rows_kernel");
  }
  if (dilation == 1)
    depthwise3x3_kernel<PX, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(src, d_weights, dst, N, H, W, C, relu_in, total);
  else
    depthwise3x3_kernel<PX, 2><<<grid, 256, 0, (cudaStream_t)stream>>>(src, d_weights, dst, N, H, W, C, relu_in, total);
  return after_launch("depthwise3x3_kernel");
}

// ---- input pipeline: light_head_preprocess_for_eval / _for_test (preprocessing/common_preprocessing.py:383-458) ------
// uint8 HWC image -> whitened fp32, bilinear WARP_RESIZE to the network input, written as one NCHW plane set:
//   tf.image.convert_image_dtype(uint8 -> float32) = cast * (1/255)   (:391)
//   * 2.                                                             (:391)
//   - [R,G,B mean / 127.5]                                           (tf_image_whitened, :136-152, :392)
//   tf.image.resize_images(BILINEAR, align_corners=False)            (:428-431): TF r1.6 ResizeBilinear (legacy
//     sampling, no half-pixel centres): in = out_index * (in_size / out_size), lower = (int)in,
//     upper = min(lower + 1, in_size - 1), lerp = in - lower;  top = tl + (tr - tl) * xl, bottom likewise,
//     value = top + (bottom - top) * yl -- all in fp32, one rounding per operation.
namespace xdet {
namespace {
__global__ void __launch_bounds__(256) preprocess_eval_kernel(const unsigned char* __restrict__ img, int H, int W,
                                                              int Ho, int Wo, float sy, float sx, float m0, float m1,
                                                              float m2, float* __restrict__ out /* [3,Ho,Wo] */) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Ho * Wo) return;
  const int x = e % Wo, y = e / Wo;
  const float in_y = __fmul_rn((float)y, sy), in_x = __fmul_rn((float)x, sx);
  const int y0 = (int)in_y, x0 = (int)in_x;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float yl = __fsub_rn(in_y, (float)y0), xl = __fsub_rn(in_x, (float)x0);
  const float mean[3] = {m0, m1, m2};
  const float k255 = 1.0f / 255.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    auto px = [&](int yy, int xx) {
      const float v = (float)img[((long long)yy * W + xx) * 3 + c];
      return __fsub_rn(__fmul_rn(__fmul_rn(v, k255), 2.0f), mean[c]);
    };
    const float tl = px(y0, x0), tr = px(y0, x1), bl = px(y1, x0), br = px(y1, x1);
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
    out[((long long)c * Ho + y) * Wo + x] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
  }
}
}  // namespace
}  // namespace xdet

extern "C" int xdet_preprocess_eval_u8(const unsigned char* d_image, int H, int W, int Ho, int Wo, const float* h_means3,
                                       float* d_out, void* stream) {
  if (H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return fail(XDET_EINVAL, "preprocess_eval: non-positive size");
  if (!h_means3) return fail(XDET_EINVAL, "preprocess_eval: means missing");
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;  // CalculateResizeScale, align_corners = false
  const int total = Ho * Wo;
  preprocess_eval_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_image, H, W, Ho, Wo, sy, sx, h_means3[0],
                                                                               h_means3[1], h_means3[2], d_out);
  return after_launch("preprocess_eval_kernel");
}
