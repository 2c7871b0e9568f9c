// Depthwise 3x3 weight gradient: the kernel the Xception TRAINING path adds to the convolution gradients
// (SURVEY 8 row a15 for the reference's own backbone).  GPU test: tests/test_depthwise_wgrad_gpu.py; the same source
// also runs under a host emulation (tests/test_staged_emulation.py).
//
// Weight gradient of tf.layers.separable_conv2d's depthwise stage (net/xception_body.py:220-234: 3x3, 'same',
// stride 1, depth multiplier 1, dilation 1 or 2, input ReLU'd by relu_separable_bn_block):
//     dW[kh][kw][c] = sum_{n,y,x} in(n, y + (kh-1)*d, x + (kw-1)*d, c) * dY(n, y, x, c)      in = relu(x) if relu_in
// The input gradient needs no kernel of its own: it is the forward kernel run on dY with the taps flipped
// (w'[kh][kw] = w[2-kh][2-kw]) and relu_in = 0, followed by the existing relu_bwd.
//
// HBM-bound: reads x and dY once (x's 8 neighbours come from L1/L2), writes 9*C floats.  Layout NHWC bf16, one
// thread per channel PAIR so that a warp covers 64 consecutive channels = 128 contiguous bytes per pixel; the 8 warps
// of a CTA take different pixels of the same 64-channel chunk, keep 9 fp32x2 partial sums in registers, fold them
// through shared memory and add ONE value per (tap, channel) and CTA to the fp32 gradient buffer (atomics: the
// last bits depend on CTA order, like the other weight gradients; DESIGN 7 "determinism").
#ifndef XDET_EMULATE_ON_CPU   // tests/staged/emulate_depthwise_wgrad.cc supplies host stand-ins instead
#include <cuda_bf16.h>

#include "common.cuh"
#endif

namespace xdet {
namespace {

constexpr int kWgThreads = 256;
constexpr int kWgWarps = kWgThreads / 32;

template <int DIL>
__global__ void __launch_bounds__(kWgThreads) depthwise3x3_wgrad_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw /* [9, C] */,
    int N, int H, int W, int C, int relu_in, int pixels_per_cta) {
  __shared__ float part[kWgWarps][9][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;          // this thread's channel pair
  const bool live = c < C;                           // C % 8 == 0, so a pair never straddles the end
  const long long total = (long long)N * H * W;
  const long long p_begin = (long long)blockIdx.y * pixels_per_cta;
  const long long p_end = p_begin + pixels_per_cta < total ? p_begin + pixels_per_cta : total;
  float2 acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = make_float2(0.f, 0.f);
  if (live) {
    for (long long p = p_begin + warp; p < p_end; p += kWgWarps) {
      const int xo = (int)(p % W);
      const int yo = (int)((p / W) % H);
      const long long n = p / ((long long)W * H);
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dy + p * C + c));
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int yi = yo + (kh - 1) * DIL;
        if (yi < 0 || yi >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int xi = xo + (kw - 1) * DIL;
          if (xi < 0 || xi >= W) continue;
          float2 v = __bfloat1622float2(
              *reinterpret_cast<const __nv_bfloat162*>(x + ((n * H + yi) * W + xi) * C + c));
          if (relu_in) {
            v.x = fmaxf(v.x, 0.f);
            v.y = fmaxf(v.y, 0.f);
          }
          acc[kh * 3 + kw].x = __fmaf_rn(v.x, g.x, acc[kh * 3 + kw].x);
          acc[kh * 3 + kw].y = __fmaf_rn(v.y, g.y, acc[kh * 3 + kw].y);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    part[warp][k][lane * 2] = acc[k].x;
    part[warp][k][lane * 2 + 1] = acc[k].y;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * 64; e += kWgThreads) {
    const int k = e / 64, cc = e % 64;
    if (blockIdx.x * 64 + cc >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWgWarps; ++w) s += part[w][k][cc];
    atomicAdd(dw + (long long)k * C + blockIdx.x * 64 + cc, s);
  }
}

#ifndef XDET_EMULATE_ON_CPU
// Row-sliding form (the one the step runs): a warp owns one image row and 128 channels (a lane: 4 channels = 8 bytes),
// walks the row once and keeps the 3 x (2*DIL+1) window of x in registers -- per pixel it loads ONE new column (3 rows)
// and dY, where the form above loads all nine taps (and ran 60 dependent pixel iterations per warp: 163 us on the
// 30x30x728 middle-flow tensors, 40x their HBM time).  The register roles rotate at compile time (pixel loop unrolled by
// the ring length), so nothing is moved.  The CTA's warps (different rows, same channels) fold through shared memory
// into one 16-byte reduction per (tap, 4 channels).
constexpr int kRowWarps = 4;

template <int DIL>
__global__ void __launch_bounds__(kRowWarps * 32) depthwise3x3_wgrad_rows_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw /* [9, C] */,
    int rows /* N*H */, int H, int W, int C, int relu_in, int rows_per_warp) {
  constexpr int R = 2 * DIL + 1;
  __shared__ float4 part[kRowWarps][9][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + lane * 4;
  const bool live = c < C;  // C % 8 == 0: four channels never straddle the end
  float acc[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;

  // Every load of the unrolled body is unconditional (out-of-range taps read a valid address and are masked to zero
  // afterwards), so that the compiler can issue a whole body's loads -- R pixels x (3 rows of x + dY) -- back to back and
  // pay ONE memory round trip per body instead of one per load.
  auto cvt4 = [&](const uint2 u, bool keep, bool relu, float* v) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    v[0] = keep ? a.x : 0.f; v[1] = keep ? a.y : 0.f; v[2] = keep ? b.x : 0.f; v[3] = keep ? b.y : 0.f;
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
  };

  const long long r0 = ((long long)blockIdx.y * kRowWarps + warp) * rows_per_warp;
  const bool relu = relu_in != 0;
  if (live) {
    for (long long r = r0; r < r0 + rows_per_warp && r < rows; ++r) {
      const int y = (int)(r % H);
      const __nv_bfloat16* grow = dy + (r * W) * (long long)C + c;
      const __nv_bfloat16* xrow[3];
      bool rok[3];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int yi = y + (kh - 1) * DIL;
        rok[kh] = yi >= 0 && yi < H;
        xrow[kh] = rok[kh] ? x + ((r + (kh - 1) * DIL) * W) * (long long)C + c : grow;  // (a valid row either way)
      }
      float win[R][3][4];  // slot of column xi: (xi + DIL) mod R
#pragma unroll
      for (int sl = 0; sl < R; ++sl)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int j = 0; j < 4; ++j) win[sl][kh][j] = 0.f;
#pragma unroll
      for (int xi = 0; xi < DIL; ++xi) {  // columns 0 .. DIL-1 (the step of pixel 0 loads column DIL)
        const int xc = xi < W ? xi : W - 1;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
          cvt4(__ldg(reinterpret_cast<const uint2*>(xrow[kh] + (long long)xc * C)), rok[kh] && xi < W, relu,
               win[(xi + DIL) % R][kh]);
      }
      for (int base = 0; base < W; base += R) {
        uint2 raw[R][3], rawg[R];
#pragma unroll
        for (int sl = 0; sl < R; ++sl) {
          const int xo = base + sl, xn = xo + DIL;  // xn: the column entering the window at this pixel
          const int xnc = xn < W ? xn : W - 1, xoc = xo < W ? xo : W - 1;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) raw[sl][kh] = __ldg(reinterpret_cast<const uint2*>(xrow[kh] + (long long)xnc * C));
          rawg[sl] = __ldg(reinterpret_cast<const uint2*>(grow + (long long)xoc * C));
        }
#pragma unroll
        for (int sl = 0; sl < R; ++sl) {
          const int xo = base + sl, xn = xo + DIL;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) cvt4(raw[sl][kh], rok[kh] && xn < W, relu, win[(sl + 2 * DIL) % R][kh]);
          float g[4];
          cvt4(rawg[sl], xo < W, false, g);  // (pixels past the row end contribute nothing)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                acc[kh * 3 + kw][j] = __fmaf_rn(win[(sl + kw * DIL) % R][kh][j], g[j], acc[kh * 3 + kw][j]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) part[warp][k][lane] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * 32; e += kRowWarps * 32) {
    const int k = e >> 5, q = e & 31;
    const int cc = blockIdx.x * 128 + q * 4;
    if (cc >= C) continue;
    float4 s4 = part[0][k][q];
#pragma unroll
    for (int w = 1; w < kRowWarps; ++w) {
      const float4 t = part[w][k][q];
      s4.x += t.x; s4.y += t.y; s4.z += t.z; s4.w += t.w;
    }
    if (s4.x != 0.f || s4.y != 0.f || s4.z != 0.f || s4.w != 0.f)
      asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw + (long long)k * C + cc),
                   "f"(s4.x), "f"(s4.y), "f"(s4.z), "f"(s4.w)
                   : "memory");
  }
}
#endif  // XDET_EMULATE_ON_CPU

}  // namespace
}  // namespace xdet

using namespace xdet;

// Grid: blockIdx.x = 64-channel chunk, blockIdx.y = slab of `per` consecutive pixels.  Enough CTAs to fill the chip
// ~4x over, but at least 64 pixels per warp so that the fold + atomics amortise.  (Plain host arithmetic, kept apart
// from the launch so that the CPU emulation of tests/staged/ drives the kernel over exactly this decomposition.)
static void depthwise_wgrad_grid(long long total, int C, int num_sms, int* chunks, int* slabs, int* per) {
  *chunks = (C + 63) / 64;
  long long s = (4ll * num_sms + *chunks - 1) / *chunks;
  const long long max_slabs = (total + 64 * 8 - 1) / (64 * 8);
  if (s > max_slabs) s = max_slabs;
  if (s < 1) s = 1;
  if (s > 65535) s = 65535;
  *per = (int)((total + s - 1) / s);
  *slabs = (int)((total + *per - 1) / *per);
}

#ifndef XDET_EMULATE_ON_CPU
using namespace xdet;

// d_dw [9, C] fp32 is ACCUMULATED into (the caller zeroes the flat gradient buffer once per step).
extern "C" int xdet_depthwise3x3_wgrad_bf16(const void* d_x, const void* d_dy, float* d_dw, int N, int H, int W, int C,
                                            int dilation, int relu_in, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0) return fail(XDET_EINVAL, "depthwise3x3_wgrad: non-positive dimension");
  if (C % 8 != 0) return fail(XDET_EINVAL, "depthwise3x3_wgrad: C (%d) must be a multiple of 8", C);
  if (dilation != 1 && dilation != 2) return fail(XDET_EINVAL, "depthwise3x3_wgrad: dilation must be 1 or 2");
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(d_x);
  const __nv_bfloat16* dyb = reinterpret_cast<const __nv_bfloat16*>(d_dy);
  if (W > 2 * dilation && (reinterpret_cast<uintptr_t>(d_dw) & 15) == 0 && C % 4 == 0 &&
      (long long)N * H < (1ll << 31)) {
    // row-sliding kernel: up to ~4 CTAs of 4 row-warps per SM, whole rows per warp
    const int rows = N * H, groups = (C + 127) / 128;
    int rpw = (int)(((long long)rows * groups + 4ll * kNumSMs * kRowWarps - 1) / (4ll * kNumSMs * kRowWarps));
    if (rpw < 1) rpw = 1;
    const int slabs_r = (rows + kRowWarps * rpw - 1) / (kRowWarps * rpw);
    const dim3 grid_r((unsigned)groups, (unsigned)slabs_r);
    if (dilation == 1)
      depthwise3x3_wgrad_rows_kernel<1><<<grid_r, kRowWarps * 32, 0, (cudaStream_t)stream>>>(xb, dyb, d_dw, rows, H, W, C,
                                                                                             relu_in, rpw);
    else
      depthwise3x3_wgrad_rows_kernel<2><<<grid_r, kRowWarps * 32, 0, (cudaStream_t)stream>>>(xb, dyb, d_dw, rows, H, W, C,
                                                                                             relu_in, rpw);
    return after_launch("depthwise3x3_wgrad_rows_kernel");
  }
  int chunks, slabs, per;
  depthwise_wgrad_grid((long long)N * H * W, C, kNumSMs, &chunks, &slabs, &per);
  const dim3 grid((unsigned)chunks, (unsigned)slabs);
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(d_x);
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(d_dy);
  if (dilation == 1)
    depthwise3x3_wgrad_kernel<1><<<grid, kWgThreads, 0, (cudaStream_t)stream>>>(x, dy, d_dw, N, H, W, C, relu_in, per);
  else
    depthwise3x3_wgrad_kernel<2><<<grid, kWgThreads, 0, (cudaStream_t)stream>>>(x, dy, d_dw, N, H, W, C, relu_in, per);
  return after_launch("depthwise3x3_wgrad_kernel");
}
#endif  // XDET_EMULATE_ON_CPU
