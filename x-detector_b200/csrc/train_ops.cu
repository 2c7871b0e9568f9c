// Training-step kernels around the tensor-core convolutions (all bandwidth / latency work, no tensor cores):
//   * batch-norm in training mode (tf.layers.batch_normalization(training=True, fused=True),
//     net/resnet_v2.py:41-50): per-channel batch statistics, finalisation (+ moving-average update), and the
//     two-pass backward of BN+ReLU;
//   * max-pool backward, bias gradient (column sums), NCHW fp32 <-> NHWC bf16 repacks for the thin feature map;
//   * losses of light_head_rfcn_train.py:257-275,361-413: sparse softmax cross-entropy and modified_smooth_l1,
//     forward value and gradient in one pass;
//   * MomentumOptimizer step (light_head_rfcn_train.py:426-441) with the L2 term of :420 folded in, writing the
//     bf16 forward / input-gradient weight packs of the next step as it goes;
//   * target assignment: iou_matrix + do_dual_max_match + box encoding (preprocessing/anchor_manipulator.py:
//     40-94,118-171,337-392) and the fg/bg sampling with up-sampling of :394-432 and
//     light_head_rfcn_train.py:321-358, tf.random_shuffle replaced by injected key arrays.
#include <cuda_bf16.h>

#include <algorithm>
#include <cfloat>

#include "common.cuh"

namespace xdet {
namespace {

unsigned blocks_for(long long total, int threads = 256, int per_sm = 8) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ---- column sums of a [rows, cs] bf16 matrix: sum (and sum of squares) per channel --------------------------
// A block of 256 threads = CGB channel groups (8 channels = 16 bytes each) x RL = 256/CGB row lanes; a thread owns
// its 8 channels for the whole kernel (per-channel constants live in registers) and strides over the rows.
// CGB = min(32, C/8) rounded down to a power of two, so narrow tensors (C = 64) still use every thread.
struct ColMap {
  int cgb, rl, cg, lane_row, c0;
  __device__ __forceinline__ ColMap(int cgb_) {
    cgb = cgb_;
    rl = 256 / cgb;
    cg = threadIdx.x % cgb;
    lane_row = threadIdx.x / cgb;
    c0 = (blockIdx.x * cgb + cg) * 8;
  }
};
__device__ __forceinline__ void block_col_reduce(float (*s)[8], const ColMap& m, int C, const float* a, const float* q,
                                                 bool sq, float* sums) {
  // s: [256][8] floats of shared memory per quantity, laid out [row lane][channel group][8]
  __shared__ float s_a[256][8], s_q[256][8];
  (void)s;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_a[threadIdx.x][j] = a[j];
    if (sq) s_q[threadIdx.x][j] = q[j];
  }
  __syncthreads();
  const int nch = m.cgb * 8;  // channels of this block
  float t = 0.f, t2 = 0.f;
  int c = C;
  if ((int)threadIdx.x < nch) {
    c = blockIdx.x * nch + threadIdx.x;
    if (c < C) {
      for (int r = 0; r < m.rl; ++r) {
        t += (&s_a[0][0])[r * nch + threadIdx.x];  // = s_a[r*cgb + cg][j] with threadIdx.x = cg*8 + j
        if (sq) t2 += (&s_q[0][0])[r * nch + threadIdx.x];
      }
    }
  }
  // four neighbouring channels leave in ONE 16-byte reduction: the L2 atomic units see a quarter of the operations (with
  // hundreds of blocks per tensor they, not HBM, bounded the small layers)
  const float u1 = __shfl_down_sync(0xffffffffu, t, 1), u2 = __shfl_down_sync(0xffffffffu, t, 2),
              u3 = __shfl_down_sync(0xffffffffu, t, 3);
  const float v1 = __shfl_down_sync(0xffffffffu, t2, 1), v2 = __shfl_down_sync(0xffffffffu, t2, 2),
              v3 = __shfl_down_sync(0xffffffffu, t2, 3);
  if (c >= C) return;
  if ((reinterpret_cast<uintptr_t>(sums) & 15) == 0) {
    if ((threadIdx.x & 3) == 0) {  // (C is a multiple of 8: c + 3 < C)
      atomicAdd(reinterpret_cast<float4*>(sums + c), make_float4(t, u1, u2, u3));
      if (sq) atomicAdd(reinterpret_cast<float4*>(sums + C + c), make_float4(t2, v1, v2, v3));
    }
  } else {
    atomicAdd(sums + c, t);
    if (sq) atomicAdd(sums + C + c, t2);
  }
}

template <bool SQ>
__global__ void __launch_bounds__(256) col_stats_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C,
                                                        int cs, int cgb, float* __restrict__ sums /* [2][C] or [C] */) {
  const ColMap m(cgb);
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = q[j] = 0.f;
  if (m.c0 < C) {
#pragma unroll 4
    for (long long r = (long long)blockIdx.y * m.rl + m.lane_row; r < rows; r += (long long)gridDim.y * m.rl) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + r * cs + m.c0));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        a[2 * k] += f.x;
        a[2 * k + 1] += f.y;
        if (SQ) {
          q[2 * k] = fmaf(f.x, f.x, q[2 * k]);
          q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
        }
      }
    }
  }
  block_col_reduce(nullptr, m, C, a, q, SQ, sums);
}

__device__ __forceinline__ void bn_finalize_channel(double s0, double s1, int c, const float* gamma, const float* beta,
                                                    long long rows, float eps, float decay, float* moving_mean,
                                                    float* moving_var, float* scale, float* shift, float* mean_out,
                                                    float* invstd_out) {
  const double m = s0 / (double)rows;
  double var = s1 / (double)rows - m * m;  // biased (what normalises the batch)
  if (var < 0.0) var = 0.0;
  const double inv = 1.0 / sqrt(var + (double)eps);
  scale[c] = (float)((double)gamma[c] * inv);
  shift[c] = (float)((double)beta[c] - m * (double)gamma[c] * inv);
  mean_out[c] = (float)m;
  invstd_out[c] = (float)inv;
  if (moving_mean) {  // fused batch norm feeds the UNBIASED variance to the moving average
    const double unb = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    moving_mean[c] = moving_mean[c] * decay + (1.f - decay) * (float)m;
    moving_var[c] = moving_var[c] * decay + (1.f - decay) * (float)unb;
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, long long rows, int C, float eps, float decay,
                                   float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  bn_finalize_channel((double)sums[c], (double)sums[C + c], c, gamma, beta, rows, eps, decay, moving_mean, moving_var,
                      scale, shift, mean_out, invstd_out);
}

// Statistics AND bn_finalize in one launch: the block that takes the last ticket finds every partial sum in L2, turns
// them into scale / shift / mean / invstd (+ the moving averages) and leaves the scratch (sums, ticket) zeroed for the
// next batch-norm on this stream -- so a layer costs no fill and no second tiny launch.
__global__ void __launch_bounds__(256) bn_stats_finalize_kernel(
    const __nv_bfloat16* __restrict__ x, long long rows, int C, int cs, int cgb, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, float decay, float* __restrict__ moving_mean,
    float* __restrict__ moving_var, float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
    float* __restrict__ invstd_out, float* __restrict__ sums /* [2][C], zero */, unsigned* __restrict__ ticket) {
  const ColMap m(cgb);
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = q[j] = 0.f;
  if (m.c0 < C) {
#pragma unroll 4
    for (long long r = (long long)blockIdx.y * m.rl + m.lane_row; r < rows; r += (long long)gridDim.y * m.rl) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + r * cs + m.c0));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        a[2 * k] += f.x;
        a[2 * k + 1] += f.y;
        q[2 * k] = fmaf(f.x, f.x, q[2 * k]);
        q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
      }
    }
  }
  block_col_reduce(nullptr, m, C, a, q, true, sums);
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1u);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += 256) {
    // (one block serves every channel here: only the cancellation-prone E[x^2] - E[x]^2 is done in double)
    const double inv_rows = 1.0 / (double)rows;
    const double md = (double)__ldcg(sums + c) * inv_rows;
    const double vd = fmax((double)__ldcg(sums + C + c) * inv_rows - md * md, 0.0);
    const float mean = (float)md, var = (float)vd;
    const float inv = 1.f / sqrtf(var + eps);
    const float sc = gamma[c] * inv;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    mean_out[c] = mean;
    invstd_out[c] = inv;
    if (moving_mean) {  // fused batch norm feeds the UNBIASED variance to the moving average
      const float unb = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
      moving_mean[c] = moving_mean[c] * decay + (1.f - decay) * mean;
      moving_var[c] = moving_var[c] * decay + (1.f - decay) * unb;
    }
    sums[c] = 0.f;
    sums[C + c] = 0.f;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// bn_finalize + normalise (+ReLU) in one launch when the statistics already exist (the producing convolution's epilogue
// accumulated them): every thread derives scale / shift of its 8 channels from the sums; the first row of blocks also
// stores them (with mean / invstd: what the backward reads) and advances the moving averages.
__global__ void __launch_bounds__(256) bn_apply_finalize_kernel(
    const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long rows, int C, int cgb,
    const float* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    float decay, float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ scale_out,
    float* __restrict__ shift_out, float* __restrict__ mean_out, float* __restrict__ invstd_out, int relu) {
  // (the convolution that follows is launched with programmatic stream serialization: let its CTAs come up -- barriers,
  // TMEM, tensor maps -- while this kernel runs; it still waits for this grid's completion before it reads anything)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const ColMap m(cgb);
  // the first row lane of the block finalizes the block's channels once and shares them (every thread repeating the
  // divisions and square roots cost more than the normalisation itself on the small layers)
  __shared__ float s_sc[256], s_sh[256];
  if (m.lane_row == 0 && m.c0 < C) {
    const float inv_rows = 1.f / (float)rows;
    const bool writer = blockIdx.y == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // fp32 throughout: E[x^2] - E[x]^2 with the product exact inside the fma
      const int c = m.c0 + j;
      const float mean = sums[c] * inv_rows;
      const float var = fmaxf(fmaf(-mean, mean, sums[C + c] * inv_rows), 0.f);
      const float inv = 1.f / sqrtf(var + eps);
      const float scj = gamma[c] * inv, shj = beta[c] - mean * scj;
      s_sc[m.cg * 8 + j] = scj;
      s_sh[m.cg * 8 + j] = shj;
      if (writer) {
        scale_out[c] = scj;
        shift_out[c] = shj;
        mean_out[c] = mean;
        invstd_out[c] = inv;
        if (moving_mean) {  // fused batch norm feeds the UNBIASED variance to the moving average
          const float unb = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
          moving_mean[c] = moving_mean[c] * decay + (1.f - decay) * mean;
          moving_var[c] = moving_var[c] * decay + (1.f - decay) * unb;
        }
      }
    }
  }
  __syncthreads();
  if (m.c0 >= C) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = s_sc[m.cg * 8 + j];
    sh[j] = s_sh[m.cg * 8 + j];
  }
#pragma unroll 4
  for (long long r = (long long)blockIdx.y * m.rl + m.lane_row; r < rows; r += (long long)gridDim.y * m.rl) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + r * C + m.c0));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = __bfloat1622float2(h[q]);
      float a = fmaf(f.x, sc[2 * q], sh[2 * q]);
      float b = fmaf(f.y, sc[2 * q + 1], sh[2 * q + 1]);
      if (relu) {
        a = fmaxf(a, 0.f);
        b = fmaxf(b, 0.f);
      }
      ho[q] = __floats2bfloat162_rn(a, b);
    }
    *reinterpret_cast<uint4*>(y + r * C + m.c0) = o;
  }
}

// g = dy * [x*scale+shift > 0] (ReLU mask recomputed from the saved pre-BN tensor); sums: [0,C) = sum g,
// [C,2C) = sum g * xhat, xhat = (x - mean) * invstd.
__global__ void __launch_bounds__(256, 4) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy,
                                                               const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, long long rows, int C,
                                                               int relu, int cgb, float* __restrict__ sums) {
  const ColMap m(cgb);
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  if (m.c0 < C) {
    {
      // the loop keeps sum g and sum g*x; xhat = (x - mean)*invstd enters once per thread afterwards
      float sc[8], sh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j] = relu ? scale[m.c0 + j] : 0.f;
        sh[j] = relu ? shift[m.c0 + j] : 1.f;  // (no ReLU: the mask below is always on)
      }
#pragma unroll 4
      for (long long r = (long long)blockIdx.y * m.rl + m.lane_row; r < rows; r += (long long)gridDim.y * m.rl) {
        const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dy + r * C + m.c0));
        const uint4 ux = __ldg(reinterpret_cast<const uint4*>(x + r * C + m.c0));
        const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&ud);
        const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&ux);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 fd = __bfloat1622float2(hd[k]), fx = __bfloat1622float2(hx[k]);
          const float g0 = (fmaf(fx.x, sc[2 * k], sh[2 * k]) > 0.f) ? fd.x : 0.f;
          const float g1 = (fmaf(fx.y, sc[2 * k + 1], sh[2 * k + 1]) > 0.f) ? fd.y : 0.f;
          a[2 * k] += g0;
          a[2 * k + 1] += g1;
          b[2 * k] = fmaf(g0, fx.x, b[2 * k]);
          b[2 * k + 1] = fmaf(g1, fx.y, b[2 * k + 1]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (b[j] - mean[m.c0 + j] * a[j]) * invstd[m.c0 + j];
  }
  block_col_reduce(nullptr, m, C, a, b, true, sums);
}

// dx = scale * (g - sum_g/M - xhat * sum_gx/M) (+ add_in) = scale*g + B*x + D with per-channel constants
//   B = -scale*invstd*sum_gx/M,  D = scale*(mean*invstd*sum_gx - sum_g)/M   (scale = gamma * invstd).
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           const __nv_bfloat16* __restrict__ x,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ sums,
                                                           const __nv_bfloat16* __restrict__ add_in, long long rows,
                                                           int C, int relu, int cgb, __nv_bfloat16* __restrict__ dx) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // (see bn_apply_finalize_kernel)
  const ColMap m(cgb);
  if (m.c0 >= C) return;
  const float inv_m = 1.f / (float)rows;
  float sc[8], sh[8], B[8], D[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = m.c0 + j;
    sc[j] = scale[c];
    sh[j] = shift[c];
    const float is = invstd[c], s0 = sums[c], s1 = sums[C + c];
    B[j] = -sc[j] * is * s1 * inv_m;
    D[j] = sc[j] * (mean[c] * is * s1 - s0) * inv_m;
  }
#pragma unroll 4
  for (long long r = (long long)blockIdx.y * m.rl + m.lane_row; r < rows; r += (long long)gridDim.y * m.rl) {
    const long long off = r * C + m.c0;
    const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dy + off));
    const uint4 ux = __ldg(reinterpret_cast<const uint4*>(x + off));
    uint4 ua = make_uint4(0, 0, 0, 0);
    if (add_in) ua = __ldg(reinterpret_cast<const uint4*>(add_in + off));
    const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&ud);
    const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&ux);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fd = __bfloat1622float2(hd[k]), fx = __bfloat1622float2(hx[k]), fa = __bfloat1622float2(ha[k]);
      const float g0 = (!relu || fmaf(fx.x, sc[2 * k], sh[2 * k]) > 0.f) ? fd.x : 0.f;
      const float g1 = (!relu || fmaf(fx.y, sc[2 * k + 1], sh[2 * k + 1]) > 0.f) ? fd.y : 0.f;
      ho[k] = __floats2bfloat162_rn(fmaf(sc[2 * k], g0, fmaf(B[2 * k], fx.x, D[2 * k])) + fa.x,
                                    fmaf(sc[2 * k + 1], g1, fmaf(B[2 * k + 1], fx.y, D[2 * k + 1])) + fa.y);
    }
    *reinterpret_cast<uint4*>(dx + off) = o;
  }
}

// dx of tf.layers.max_pooling2d(3, 2, 'SAME'): every input pixel gathers dy from the (<= 4) windows that cover it
// and whose FIRST maximum it is -- the forward kernel recorded that position (kh*3 + kw) per output element.
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const unsigned char* __restrict__ argmax,
                                                          const __nv_bfloat16* __restrict__ dy,
                                                          __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C,
                                                          int Ho, int Wo, int pad_top, int pad_left, long long total) {
  const int C8 = C / 8;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int c8 = (int)(e % C8);
    long long t = e / C8;
    const int xi = (int)(t % W);
    t /= W;
    const int yi = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // windows (yo, xo) with yo*2 - pad_top <= yi <= yo*2 - pad_top + 2: at most two per axis.  All (<= 4) candidates
    // are loaded unconditionally from clamped positions and masked afterwards, so that the eight loads leave together
    // (a `continue` per window made every load wait for the one before it).
    const int yo_lo = max(0, (yi + pad_top - 1) / 2), yo_hi = min(Ho - 1, (yi + pad_top) / 2);
    const int xo_lo = max(0, (xi + pad_left - 1) / 2), xo_hi = min(Wo - 1, (xi + pad_left) / 2);
    uint2 am[4];
    uint4 dv[4];
    unsigned code[4];
    bool ok[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int yo = yo_lo + (w >> 1), xo = xo_lo + (w & 1);
      const int kh = yi - (yo * 2 - pad_top), kw = xi - (xo * 2 - pad_left);
      ok[w] = yo <= yo_hi && xo <= xo_hi && kh >= 0 && kh <= 2 && kw >= 0 && kw <= 2;
      code[w] = (unsigned)(kh * 3 + kw);
      const long long o = (((long long)n * Ho + min(yo, Ho - 1)) * Wo + min(xo, Wo - 1)) * C8 + c8;
      am[w] = __ldg(reinterpret_cast<const uint2*>(argmax) + o);
      dv[w] = __ldg(reinterpret_cast<const uint4*>(dy) + o);
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&dv[w]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(hd[q]);
        const unsigned wd = (q < 2) ? am[w].x : am[w].y;
        const unsigned a0 = (wd >> ((2 * q & 3) * 8)) & 255u, a1 = (wd >> (((2 * q + 1) & 3) * 8)) & 255u;
        if (ok[w] && a0 == code[w]) acc[2 * q] += f.x;
        if (ok[w] && a1 == code[w]) acc[2 * q + 1] += f.y;
      }
    }
    uint4 o4;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o4);
#pragma unroll
    for (int q = 0; q < 4; ++q) ho[q] = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
    reinterpret_cast<uint4*>(dx)[e] = o4;
  }
}

__global__ void relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, uint4* __restrict__ dx,
                                long long n8) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n8; e += step) {
    const uint4 d = __ldg(dy + e), v = __ldg(y + e);
    const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&d);
    const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&v);
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fd = __bfloat1622float2(hd[k]), fv = __bfloat1622float2(hv[k]);
      ho[k] = __floats2bfloat162_rn(fv.x > 0.f ? fd.x : 0.f, fv.y > 0.f ? fd.y : 0.f);
    }
    dx[e] = o;
  }
}

// [N,C,H,W] fp32 -> [N,H,W,C] bf16 and [N,H,W,C] bf16 -> relu(x*scale+shift) as [N,C,H,W] fp32 (thin feature map)
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int C,
                                             int cs, int HW, long long total) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int c = (int)(e % cs);
    const long long p = e / cs;  // n*HW + hw
    const long long n = p / HW, hw = p % HW;
    dst[e] = __float2bfloat16_rn(c < C ? __ldg(src + (n * C + c) * HW + hw) : 0.f);  // zero channel tail
  }
}
__global__ void affine_relu_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, const float* __restrict__ scale,
                                               const float* __restrict__ shift, float* __restrict__ dst, int N, int C,
                                               int cs, int HW, int relu, long long total) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const long long hw = e % HW;
    const long long t = e / HW;
    const int c = (int)(t % C);
    const long long n = t / C;
    float v = fmaf(__bfloat162float(src[(n * HW + hw) * cs + c]), __ldg(scale + c), __ldg(shift + c));
    if (relu) v = fmaxf(v, 0.f);
    dst[e] = v;
  }
}

// ---- losses ------------------------------------------------------------------------------------------------
// tf.nn.sparse_softmax_cross_entropy_with_logits per row; dlogits = row_w * (softmax - onehot).
__global__ void softmax_ce_kernel(const float* __restrict__ logits, int ld, int C, const int* __restrict__ labels,
                                  const float* __restrict__ row_w, float w_all, long long M, float* __restrict__ loss_row,
                                  float* __restrict__ dlogits, int dld) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const float* z = logits + r * ld;
  float mx = -FLT_MAX;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, z[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(z[c] - mx);
  const int y = labels[r];
  const float lse = logf(s) + mx;
  if (loss_row) loss_row[r] = lse - z[y];
  if (dlogits) {
    const float w = w_all * (row_w ? row_w[r] : 1.f);
    for (int c = 0; c < C; ++c) dlogits[r * dld + c] = w * (expf(z[c] - lse) - (c == y ? 1.f : 0.f));
  }
}

// modified_smooth_l1 (sigma = 1) summed over the 4 coordinates of a row, times row_w; dpred = w_all*row_w*clip(d,-1,1).
__global__ void smooth_l1_kernel(const float* __restrict__ pred, int ld, const float* __restrict__ target,
                                 const float* __restrict__ row_w, float w_all, long long M, float* __restrict__ loss_row,
                                 float* __restrict__ dpred, int dld) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const float rw = row_w ? row_w[r] : 1.f;
  float tot = 0.f;
  for (int j = 0; j < 4; ++j) {
    const float d = pred[r * ld + j] - target[r * 4 + j];
    const float ad = fabsf(d);
    tot += (ad < 1.f) ? 0.5f * d * d : ad - 0.5f;
    if (dpred) dpred[r * dld + j] = w_all * rw * ((ad < 1.f) ? d : (d > 0.f ? 1.f : -1.f));
  }
  if (loss_row) loss_row[r] = tot * rw;
}

// ---- optimizer -----------------------------------------------------------------------------------------------
// MomentumOptimizer: accum = momentum*accum + (g + wd*w); w -= lr*accum.  The master weights keep TF's layout
// [KH*KW][Cin][Cout]; dw arrives in the packed layout [Cout][KH*KW][cin_pad]; the bf16 packs of the next step
// (forward [Cout][tap][cin_pad] and input-gradient [Cin][flipped tap][cout_pad]) are rewritten on the fly.
// fold != 0: the stem layout, packed element kw*8 + ci of filter row kh.
__global__ void sgd_conv_kernel(const float* __restrict__ dw, float* __restrict__ w, float* __restrict__ mom,
                                __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ wd_pack, int Cout, int KH,
                                int KW, int Cin, int cin_pad, int cout_pad, int fold, float lr, float momentum, float wd,
                                float gscale, long long total) {
  const int taps = KH * KW;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int co = (int)(e % Cout);
    long long t = e / Cout;
    const int ci = (int)(t % Cin);
    const int tap = (int)(t / Cin);  // e indexes the master [tap][ci][co]
    long long pidx;
    if (fold) {
      const int kh = tap / KW, kw = tap - kh * KW;
      pidx = ((long long)co * KH + kh) * 64 + kw * 8 + ci;
    } else {
      pidx = ((long long)co * taps + tap) * cin_pad + ci;
    }
    const float wv = w[e];
    const float g = dw[pidx] * gscale + wd * wv;
    const float a = momentum * mom[e] + g;
    mom[e] = a;
    const float nw = wv - lr * a;
    w[e] = nw;
    const __nv_bfloat16 b = __float2bfloat16_rn(nw);
    wp[pidx] = b;
    if (wd_pack) wd_pack[((long long)ci * taps + (taps - 1 - tap)) * cout_pad + co] = b;
  }
}
// Tiled form for the regular layout: a block owns a 32 (ci) x 32 (co) tile of one tap and transposes through shared
// memory, so the packed arrays (ci fastest) and the TF-layout masters (co fastest) are BOTH accessed coalesced.
__global__ void __launch_bounds__(256) sgd_conv_tiled_kernel(const float* __restrict__ dw, float* __restrict__ w,
                                                             float* __restrict__ mom, __nv_bfloat16* __restrict__ wp,
                                                             __nv_bfloat16* __restrict__ wd_pack, int Cout, int taps,
                                                             int Cin, int cin_pad, int cout_pad, float lr, float momentum,
                                                             float wd, float gscale) {
  __shared__ float s_g[32][33];
  __shared__ __nv_bfloat16 s_w[32][34];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32, tap = blockIdx.z;
  for (int l = ty; l < 32; l += 8) {
    const int co = co0 + l, ci = ci0 + tx;
    s_g[l][tx] = (co < Cout && ci < Cin) ? dw[((long long)co * taps + tap) * cin_pad + ci] : 0.f;
  }
  __syncthreads();
  for (int l = ty; l < 32; l += 8) {
    const int ci = ci0 + l, co = co0 + tx;
    if (ci < Cin && co < Cout) {
      const long long e = ((long long)tap * Cin + ci) * Cout + co;
      const float wv = w[e];
      const float a = momentum * mom[e] + (s_g[tx][l] * gscale + wd * wv);
      mom[e] = a;
      const float nw = wv - lr * a;
      w[e] = nw;
      const __nv_bfloat16 b = __float2bfloat16_rn(nw);
      s_w[tx][l] = b;
      if (wd_pack) wd_pack[((long long)ci * taps + (taps - 1 - tap)) * cout_pad + co] = b;
    }
  }
  __syncthreads();
  for (int l = ty; l < 32; l += 8) {
    const int co = co0 + l, ci = ci0 + tx;
    if (co < Cout && ci < Cin) wp[((long long)co * taps + tap) * cin_pad + ci] = s_w[l][tx];
  }
}

__global__ void sgd_vec_kernel(const float* __restrict__ g, float* __restrict__ w, float* __restrict__ mom, long long n,
                               float lr, float momentum, float wd, float gscale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float wv = w[i];
  const float a = momentum * mom[i] + g[i] * gscale + wd * wv;
  mom[i] = a;
  w[i] = wv - lr * a;
}

// Every variable of the model in ONE launch: a block looks its item up in the table by its block index (the host filled
// first_block with the running block count) and then does what sgd_conv_tiled_kernel / sgd_vec_kernel do.
__global__ void __launch_bounds__(256) sgd_multi_kernel(const xdet_sgd_item* __restrict__ items, int n_items, float lr,
                                                        float momentum, float gscale) {
  __shared__ float s_g[32][33];
  __shared__ __nv_bfloat16 s_w[32][34];
  int lo = 0, hi = n_items - 1;  // last item with first_block <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&items[mid].first_block) <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const xdet_sgd_item it = items[lo];
  const int b = (int)blockIdx.x - it.first_block;
  float* __restrict__ w = it.w;
  float* __restrict__ mom = it.mom;
  const float* __restrict__ dw = it.dw;
  const float wd = it.wd;
  if (it.taps == 0) {  // a vector (bias / beta / gamma) of Cout elements
    const int i = b * 256 + threadIdx.x;
    if (i < it.Cout) {
      const float wv = w[i];
      const float a = momentum * mom[i] + dw[i] * gscale + wd * wv;
      mom[i] = a;
      w[i] = wv - lr * a;
    }
    return;
  }
  __nv_bfloat16* __restrict__ wp = reinterpret_cast<__nv_bfloat16*>(it.w_pack);
  __nv_bfloat16* __restrict__ wd_pack = reinterpret_cast<__nv_bfloat16*>(it.w_dgrad_pack);
  const int Cout = it.Cout, taps = it.taps, Cin = it.Cin, cin_pad = it.cin_pad, cout_pad = it.cout_pad;
  const int per_tap = it.tiles_ci * it.tiles_co;
  const int tap = b / per_tap, r = b - tap * per_tap;
  const int co0 = (r / it.tiles_ci) * 32, ci0 = (r % it.tiles_ci) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int l = ty; l < 32; l += 8) {
    const int co = co0 + l, ci = ci0 + tx;
    s_g[l][tx] = (co < Cout && ci < Cin) ? dw[((long long)co * taps + tap) * cin_pad + ci] : 0.f;
  }
  __syncthreads();
  for (int l = ty; l < 32; l += 8) {
    const int ci = ci0 + l, co = co0 + tx;
    if (ci < Cin && co < Cout) {
      const long long e = ((long long)tap * Cin + ci) * Cout + co;
      const float wv = w[e];
      const float a = momentum * mom[e] + (s_g[tx][l] * gscale + wd * wv);
      mom[e] = a;
      const float nw = wv - lr * a;
      w[e] = nw;
      const __nv_bfloat16 bv = __float2bfloat16_rn(nw);
      s_w[tx][l] = bv;
      if (wd_pack) wd_pack[((long long)ci * taps + (taps - 1 - tap)) * cout_pad + co] = bv;
    }
  }
  __syncthreads();
  for (int l = ty; l < 32; l += 8) {
    const int co = co0 + l, ci = ci0 + tx;
    if (co < Cout && ci < Cin) wp[((long long)co * taps + tap) * cin_pad + ci] = s_w[l][tx];
  }
}

// ---- target assignment ---------------------------------------------------------------------------------------
// boxes [N,A,4] (ymin,xmin,ymax,xmax); gt [N,G,4], gt_labels [N,G] (<= 0: padding, ignored).
__device__ __forceinline__ float iou_ref(const float4 g, const float4 b) {  // iou_matrix, :20-46
  const float iy0 = fmaxf(g.x, b.x), ix0 = fmaxf(g.y, b.y), iy1 = fminf(g.z, b.z), ix1 = fminf(g.w, b.w);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(iy1, iy0), 0.f), fmaxf(__fsub_rn(ix1, ix0), 0.f));
  const float ag = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
  const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float uni = __fsub_rn(__fadd_rn(ag, ab), inter);
  return uni == 0.f ? 0.f : __fdiv_rn(inter, uni);
}
__device__ __forceinline__ bool inside_ok(const float4 b, float border) {
  return b.x >= -border && b.y >= -border && b.z < 1.f + border && b.w < 1.f + border;
}
// pass 1: per ground truth the best box (first maximum): 64-bit atomicMax on (iou bits << 32 | ~index).  A warp first
// settles its own maximum (largest overlap, then smallest index) and sends ONE atomic per ground truth: every thread
// hammering the same G addresses made this the longest kernel of the target assignment.
__global__ void __launch_bounds__(256) match_gt_argmax_kernel(const float* __restrict__ boxes, long long box_img_stride,
                                                              const float* __restrict__ gt,
                                                              const int* __restrict__ gt_labels, int A, int G,
                                                              float border,
                                                              unsigned long long* __restrict__ best /* [N,G], zero */) {
  const int img = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = a < A;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) b = reinterpret_cast<const float4*>(boxes + (long long)img * box_img_stride)[a];
  const bool in = live && inside_ok(b, border);
  for (int g = 0; g < G; ++g) {
    if (gt_labels[img * G + g] <= 0) continue;  // (uniform over the block)
    const float ov = in ? iou_ref(reinterpret_cast<const float4*>(gt)[img * G + g], b) : 0.f;
    const unsigned bits = live ? __float_as_uint(ov) : 0u;  // overlaps are >= 0: their bit patterns order like the values
    const unsigned top = __reduce_max_sync(0xffffffffu, bits);
    const unsigned first = __reduce_min_sync(0xffffffffu, (live && bits == top) ? (unsigned)a : 0xFFFFFFFFu);
    if ((threadIdx.x & 31) == 0 && first != 0xFFFFFFFFu)
      atomicMax(best + img * G + g, ((unsigned long long)top << 32) | (0xFFFFFFFFu - first));
  }
}
// pass 2: do_dual_max_match + encode.  labels_out: matched class (> 0), 0 = background, -1 = ignore.
// ref_yxhw: optional [A,4] (cy,cx,h,w) reference boxes used for the encoding (the anchors' own centre form,
// encode_anchor :118-171); NULL = point2center of the box itself (ext_encode_rois :374-378).
__global__ void match_encode_kernel(const float* __restrict__ boxes, long long box_img_stride,
                                    const float* __restrict__ ref_yxhw, const float* __restrict__ gt,
                                    const int* __restrict__ gt_labels, const unsigned long long* __restrict__ best, int A,
                                    int G, float border, float high, float low, float p0, float p1, float p2, float p3,
                                    int* __restrict__ labels_out, float* __restrict__ targets_out,
                                    float* __restrict__ scores_out) {
  const int img = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const float4 b = reinterpret_cast<const float4*>(boxes + (long long)img * box_img_stride)[a];
  const bool in = inside_ok(b, border);
  float mv = -1.f, forced_val = -1.f;
  int a2g = -1, forced_g = -1;
  bool has = false;
  for (int g = 0; g < G; ++g) {
    if (gt_labels[img * G + g] <= 0) continue;  // tf.boolean_mask(_labels > 0): order of the valid ones is kept
    const float ov = in ? iou_ref(reinterpret_cast<const float4*>(gt)[img * G + g], b) : 0.f;
    if (ov > mv) {  // anchors_to_gt: first maximum
      mv = ov;
      a2g = g;
    }
    const unsigned bi = 0xFFFFFFFFu - (unsigned)(best[img * G + g] & 0xFFFFFFFFull);
    const bool picked = (bi == (unsigned)a);  // left_gt_to_anchors_mask[g, a]
    has = has || picked;
    const float lv = picked ? ov : 0.f;  // left_gt_to_anchors_scores
    if (lv > forced_val) {               // argmax over ALL valid g, first maximum
      forced_val = lv;
      forced_g = g;
    }
  }
  const long long o = (long long)img * A + a;
  if (a2g < 0) {  // no ground truth at all
    labels_out[o] = 0;
    scores_out[o] = 0.f;
    reinterpret_cast<float4*>(targets_out)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  int match;  // >= 0: gt index, -1 negative, -2 ignore
  if (has) {
    match = forced_g;
  } else {
    match = a2g;
    if (mv < low) match = -1;
    if (mv < high && mv >= low) match = -2;
  }
  const int gsel = has ? forced_g : a2g;
  scores_out[o] = in ? iou_ref(reinterpret_cast<const float4*>(gt)[img * G + gsel], b) : 0.f;
  labels_out[o] = match >= 0 ? gt_labels[img * G + match] : (match == -2 ? -1 : 0);
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (match >= 0) {
    const float4 gb = reinterpret_cast<const float4*>(gt)[img * G + match];
    float yref, xref, h, w;
    if (ref_yxhw) {
      const float4 r = reinterpret_cast<const float4*>(ref_yxhw)[a];
      yref = r.x; xref = r.y; h = r.z; w = r.w;
    } else {
      h = __fsub_rn(b.z, b.x);
      w = __fsub_rn(b.w, b.y);
      yref = __fadd_rn(b.x, __fmul_rn(h, 0.5f));
      xref = __fadd_rn(b.y, __fmul_rn(w, 0.5f));
    }
    const float gcy = __fmul_rn(__fadd_rn(gb.z, gb.x), 0.5f), gcx = __fmul_rn(__fadd_rn(gb.w, gb.y), 0.5f);
    const float gh = __fsub_rn(gb.z, gb.x), gw = __fsub_rn(gb.w, gb.y);
    t.x = __fdiv_rn(__fdiv_rn(__fsub_rn(gcy, yref), h), p0);
    t.y = __fdiv_rn(__fdiv_rn(__fsub_rn(gcx, xref), w), p1);
    t.z = __fdiv_rn(logf(__fdiv_rn(gh, h)), p2);
    t.w = __fdiv_rn(logf(__fdiv_rn(gw, w)), p3);
  }
  reinterpret_cast<float4*>(targets_out)[o] = t;
}

// ---- fg / bg sampling with up-sampling (anchor_manipulator.py:394-432, light_head_rfcn_train.py:321-358) --------
constexpr int kSampThreads = 1024;

// out[0..k) = indices of the k smallest keys[0..n) in ascending (key, index) order -- the first k entries of the
// stable argsort that stands in for tf.random_shuffle(tf.range(n)).  keys >= 0.  Whole CTA; sk = smem [P], P >= k.
__device__ void select_smallest(const float* __restrict__ keys, int n, int k, int* out, unsigned long long* sk, int P,
                                unsigned int* hist, unsigned long long* s_prefix, int* s_remaining, int* s_count) {
  const int tid = threadIdx.x;
  auto key64 = [&](int i) {  // larger = smaller (key, index)
    return ~(((unsigned long long)__float_as_uint(keys[i]) << 32) | (unsigned long long)(unsigned)i);
  };
  if (k <= 0) return;
  if (tid == 0) {
    *s_prefix = 0ull;
    *s_remaining = k;
    *s_count = 0;  // (doubles as the "resolved far enough" flag of the pass loop)
  }
  __syncthreads();
  // Radix select from the top byte down, but only until the candidates fit the sort buffer: after a pass, G keys lie
  // above the threshold bucket (all selected) and E inside it; once G + E <= P the bucket's keys are simply sorted
  // along.  With 150 000 anchors that is 2 passes over the keys instead of 8.  Four independent loads per thread and
  // iteration (clamped index, masked afterwards) keep the memory pipe busy.
  constexpr int U = 4;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = tid; i < 256; i += kSampThreads) hist[i] = 0;
    __syncthreads();
    const unsigned long long prefix = *s_prefix;
    for (int i0 = tid; i0 < n; i0 += U * kSampThreads) {
      unsigned long long kk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) kk[u] = key64(min(i0 + u * kSampThreads, n - 1));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool in = i0 + u * kSampThreads < n;
        const bool match = (pass == 0) || ((kk[u] >> (shift + 8)) == (prefix >> (shift + 8)));
        if (in && match) atomicAdd(&hist[(unsigned)(kk[u] >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (tid == 0) {
      int rem = *s_remaining;
      int b = 255;
      for (; b > 0; --b) {
        if ((int)hist[b] >= rem) break;
        rem -= (int)hist[b];
      }
      *s_prefix = prefix | ((unsigned long long)b << shift);
      *s_remaining = rem;
      if ((k - rem) + (int)hist[b] <= P) *s_count = 1;
    }
    __syncthreads();
    const int done = *s_count;
    __syncthreads();
    if (done) break;
  }
  const unsigned long long thr = *s_prefix;  // (low bits zero when the loop left early: the whole bucket qualifies)
  if (tid == 0) *s_count = 0;
  for (int i = tid; i < P; i += kSampThreads) sk[i] = 0ull;
  __syncthreads();
  for (int i0 = tid; i0 < n; i0 += U * kSampThreads) {
    unsigned long long kk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) kk[u] = key64(min(i0 + u * kSampThreads, n - 1));
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * kSampThreads < n && kk[u] >= thr) sk[atomicAdd(s_count, 1)] = kk[u];
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (P >> 1); t += kSampThreads) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = sk[lo], b = sk[hi];
        if ((a < b) == desc) {
          sk[lo] = b;
          sk[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += kSampThreads) out[j] = (int)((~sk[j]) & 0xFFFFFFFFull);
  __syncthreads();
}

// One CTA per group (an image for RoIs; the whole flattened batch for RPN anchors).
__global__ void __launch_bounds__(kSampThreads) sample_fg_bg_kernel(
    const int* __restrict__ labels, const float* __restrict__ scores /* or NULL */, float bg_low, int n, int exp_fg,
    int total, int P, const float* __restrict__ keys_fg, const float* __restrict__ keys_bg,
    const float* __restrict__ keys_up, int* __restrict__ ws /* [groups][2n] */, int* __restrict__ out /* [groups][total] */,
    int* __restrict__ counts /* [groups][3]: n_pos, n_neg, n_keep (may be NULL) */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* sk = reinterpret_cast<unsigned long long*>(smem_raw);  // [P]
  int* keep = reinterpret_cast<int*>(sk + P);                                // [total]
  int* sel = keep + total;                                                   // [total]
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_count;
  __shared__ int s_pos[65], s_neg[65];
  const int grp = blockIdx.x, tid = threadIdx.x;
  labels += (long long)grp * n;
  if (scores) scores += (long long)grp * n;
  keys_fg += (long long)grp * n;
  keys_bg += (long long)grp * n;
  keys_up += (long long)grp * total;
  int* pos_list = ws + (long long)grp * 2 * n;
  int* neg_list = pos_list + n;
  out += (long long)grp * total;

  // ordered compaction of positives / negatives: coalesced tiles of 1024 labels (the next tile's loads are issued before
  // the current one is placed), ranks inside a warp by ballot + popc, the 32 warp totals scanned by warp 0
  {
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    if (tid == 0) s_pos[64] = s_neg[64] = 0;  // running totals; [0,32) / [32,64): per-warp offsets of even / odd tiles
    int l_next = tid < n ? labels[tid] : -1;
    float s_next = (scores && tid < n) ? scores[tid] : 0.f;
    __syncthreads();
    int par = 0;
    for (int t0 = 0; t0 < n; t0 += kSampThreads, par ^= 32) {
      const int i = t0 + tid, l = l_next;
      const float sc = s_next;
      const int inx = i + kSampThreads;
      l_next = inx < n ? labels[inx] : -1;
      s_next = (scores && inx < n) ? scores[inx] : 0.f;
      const bool isp = l > 0, isn = (l == 0) && (!scores || sc > bg_low);
      const unsigned bp = __ballot_sync(0xffffffffu, isp), bn = __ballot_sync(0xffffffffu, isn);
      if (lane == 0) {
        s_pos[par + warp] = __popc(bp);
        s_neg[par + warp] = __popc(bn);
      }
      __syncthreads();
      if (warp == 0) {  // exclusive scan of the warp totals, offset by the running totals
        int vp = s_pos[par + lane], vn = s_neg[par + lane];
        int ip = vp, in_ = vn;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, ip, d), un = __shfl_up_sync(0xffffffffu, in_, d);
          if (lane >= d) {
            ip += up;
            in_ += un;
          }
        }
        const int bp0 = s_pos[64], bn0 = s_neg[64];
        s_pos[par + lane] = bp0 + ip - vp;
        s_neg[par + lane] = bn0 + in_ - vn;
        if (lane == 31) {
          s_pos[64] = bp0 + ip;
          s_neg[64] = bn0 + in_;
        }
      }
      __syncthreads();
      if (isp) pos_list[s_pos[par + warp] + __popc(bp & lt)] = i;
      if (isn) neg_list[s_neg[par + warp] + __popc(bn & lt)] = i;
      // (the next tile writes the OTHER half of the offset arrays; its first barrier orders it after these reads)
    }
    __syncthreads();
  }
  const int n_pos = s_pos[64], n_neg = s_neg[64];
  int n_fg;
  if (n_pos < exp_fg) {
    n_fg = n_pos;
    for (int j = tid; j < n_fg; j += kSampThreads) keep[j] = pos_list[j];
  } else {
    n_fg = exp_fg;
    select_smallest(keys_fg, n_pos, n_fg, sel, sk, P, hist, &s_prefix, &s_remaining, &s_count);
    for (int j = tid; j < n_fg; j += kSampThreads) keep[j] = pos_list[sel[j]];
  }
  __syncthreads();
  const int exp_bg = total - min(n_pos, exp_fg);
  int n_bg;
  if (n_neg < exp_bg) {
    n_bg = n_neg;
    for (int j = tid; j < n_bg; j += kSampThreads) keep[n_fg + j] = neg_list[j];
  } else {
    n_bg = exp_bg;
    select_smallest(keys_bg, n_neg, n_bg, sel, sk, P, hist, &s_prefix, &s_remaining, &s_count);
    for (int j = tid; j < n_bg; j += kSampThreads) keep[n_fg + j] = neg_list[sel[j]];
  }
  __syncthreads();
  const int n_keep = n_fg + n_bg;
  if (counts && tid == 0) {
    counts[grp * 3] = n_pos;
    counts[grp * 3 + 1] = n_neg;
    counts[grp * 3 + 2] = n_keep;
  }
  if (n_keep == 0) {
    for (int j = tid; j < total; j += kSampThreads) out[j] = 0;
    return;
  }
  if (n_keep >= total) {
    for (int j = tid; j < total; j += kSampThreads) out[j] = keep[j];
    return;
  }
  const int left = total - n_keep;
  const int rem = left % n_keep;
  const int tiled = n_keep * (left / n_keep + 1);
  select_smallest(keys_up, n_keep, rem, sel, sk, P, hist, &s_prefix, &s_remaining, &s_count);
  __syncthreads();
  for (int j = tid; j < total; j += kSampThreads) out[j] = keep[j < tiled ? (j % n_keep) : sel[j - tiled]];
}

}  // namespace
}  // namespace xdet

using namespace xdet;

// channel groups per block (power of two <= min(32, C/8)) and the launch grid of the column-owner kernels
static int col_cgb(int C) {
  int g = 1;
  while (g * 2 <= 32 && g * 2 <= C / 8) g *= 2;
  return g;
}
static dim3 col_grid(long long rows, int C, int cgb, int per_sm = 4, int min_rows = 4) {
  // one wave: at most per_sm blocks per SM over all channel-group columns, and at least min_rows rows per thread (a
  // tensor of 7200 rows x 256 channels still spreads over ~225 blocks)
  const int gx = (C / 8 + cgb - 1) / cgb;
  const int rl = 256 / cgb;
  long long slabs = (rows + (long long)rl * min_rows - 1) / ((long long)rl * min_rows);
  const long long cap = std::max(1LL, (long long)kNumSMs * per_sm / gx);
  if (slabs > cap) slabs = cap;
  if (slabs < 1) slabs = 1;
  return dim3((unsigned)gx, (unsigned)slabs);
}

extern "C" int xdet_col_stats_bf16(const void* d_x, long long rows, int C, int cs, int with_squares, float* d_sums,
                                   void* stream) {
  if (rows <= 0 || C <= 0) return XDET_OK;
  if (C % 8 || cs % 8 || cs < C) return fail(XDET_EINVAL, "col_stats: C and the row pitch must be multiples of 8");
  const int cgb = col_cgb(C);
  const dim3 grid = col_grid(rows, C, cgb);
  if (with_squares)
    col_stats_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(d_x), rows, C, cs, cgb, d_sums);
  else
    col_stats_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(d_x), rows, C, cs, cgb, d_sums);
  return after_launch("col_stats_kernel");
}

extern "C" int xdet_bn_finalize(const float* d_sums, const float* d_gamma, const float* d_beta, long long rows, int C,
                                float eps, float decay, float* d_moving_mean, float* d_moving_var, float* d_scale,
                                float* d_shift, float* d_mean, float* d_invstd, void* stream) {
  if (C <= 0 || rows <= 0) return fail(XDET_EINVAL, "bn_finalize: empty");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_sums, d_gamma, d_beta, rows, C, eps, decay,
                                                                        d_moving_mean, d_moving_var, d_scale, d_shift,
                                                                        d_mean, d_invstd);
  return after_launch("bn_finalize_kernel");
}

extern "C" size_t xdet_bn_train_scratch_bytes(int C) { return sizeof(float) * 2 * (size_t)(C > 0 ? C : 0) + 16; }

extern "C" int xdet_bn_train_stats_bf16(const void* d_x, long long rows, int C, int cs, const float* d_gamma,
                                        const float* d_beta, float eps, float decay, float* d_moving_mean,
                                        float* d_moving_var, float* d_scale, float* d_shift, float* d_mean,
                                        float* d_invstd, void* d_scratch, void* stream) {
  if (rows <= 0 || C <= 0) return fail(XDET_EINVAL, "bn_train_stats: empty");
  if (C % 8 || cs % 8 || cs < C) return fail(XDET_EINVAL, "bn_train_stats: C and the row pitch must be multiples of 8");
  if (!d_scratch) return fail(XDET_EINVAL, "bn_train_stats: no scratch");
  const int cgb = col_cgb(C);
  const dim3 grid = col_grid(rows, C, cgb, 6);
  float* sums = reinterpret_cast<float*>(d_scratch);
  bn_stats_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_x), rows, C, cs, cgb, d_gamma, d_beta, eps, decay, d_moving_mean,
      d_moving_var, d_scale, d_shift, d_mean, d_invstd, sums, reinterpret_cast<unsigned*>(sums + 2 * (size_t)C));
  return after_launch("bn_stats_finalize_kernel");
}

extern "C" int xdet_bn_train_apply_bf16(const void* d_x, void* d_y, long long rows, int C, const float* d_sums,
                                        const float* d_gamma, const float* d_beta, float eps, float decay,
                                        float* d_moving_mean, float* d_moving_var, float* d_scale, float* d_shift,
                                        float* d_mean, float* d_invstd, int relu, void* stream) {
  if (rows <= 0 || C <= 0) return fail(XDET_EINVAL, "bn_train_apply: empty");
  if (C % 8) return fail(XDET_EINVAL, "bn_train_apply: C must be a multiple of 8");
  const int cgb = col_cgb(C);
  const dim3 grid = col_grid(rows, C, cgb, 8, 4);
  bn_apply_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_x), reinterpret_cast<__nv_bfloat16*>(d_y), rows, C, cgb, d_sums, d_gamma,
      d_beta, eps, decay, d_moving_mean, d_moving_var, d_scale, d_shift, d_mean, d_invstd, relu);
  return after_launch("bn_apply_finalize_kernel");
}

extern "C" int xdet_bn_relu_bwd_bf16(const void* d_dy, const void* d_x, const float* d_scale, const float* d_shift,
                                     const float* d_mean, const float* d_invstd, long long rows, int C, int relu,
                                     const void* d_add_in, float* d_sums, void* d_dx, int sums_zeroed, void* stream) {
  if (rows <= 0 || C <= 0) return XDET_OK;
  if (C % 8) return fail(XDET_EINVAL, "bn_relu_bwd: C must be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  if (!sums_zeroed) XDET_TRY(check_cuda(cudaMemsetAsync(d_sums, 0, sizeof(float) * 2 * C, st), "memset(bn sums)"));
  const int cgb = col_cgb(C);
  const dim3 grid = col_grid(rows, C, cgb);
  bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(d_dy),
                                             reinterpret_cast<const __nv_bfloat16*>(d_x), d_scale, d_shift, d_mean,
                                             d_invstd, rows, C, relu, cgb, d_sums);
  XDET_TRY(after_launch("bn_bwd_reduce_kernel"));
  bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_dy), reinterpret_cast<const __nv_bfloat16*>(d_x), d_scale, d_shift, d_mean,
      d_invstd, d_sums, reinterpret_cast<const __nv_bfloat16*>(d_add_in), rows, C, relu, cgb,
      reinterpret_cast<__nv_bfloat16*>(d_dx));
  return after_launch("bn_bwd_apply_kernel");
}

extern "C" int xdet_relu_bwd_bf16(const void* d_dy, const void* d_y, void* d_dx, long long n, void* stream) {
  if (n <= 0) return XDET_OK;
  if (n % 8) return fail(XDET_EINVAL, "relu_bwd: n must be a multiple of 8");
  relu_bwd_kernel<<<blocks_for(n / 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(d_dy), reinterpret_cast<const uint4*>(d_y), reinterpret_cast<uint4*>(d_dx), n / 8);
  return after_launch("relu_bwd_kernel");
}

extern "C" int xdet_maxpool3x3s2_bwd_bf16(const void* d_argmax, const void* d_dy, void* d_dx, int N, int H, int W, int C,
                                          int Ho, int Wo, int pad_top, int pad_left, void* stream) {
  if (C % 8) return fail(XDET_EINVAL, "maxpool_bwd: C must be a multiple of 8");
  const long long total = (long long)N * H * W * (C / 8);
  if (total <= 0) return XDET_OK;
  maxpool_bwd_kernel<<<blocks_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned char*>(d_argmax), reinterpret_cast<const __nv_bfloat16*>(d_dy),
      reinterpret_cast<__nv_bfloat16*>(d_dx), N, H, W, C, Ho, Wo, pad_top, pad_left, total);
  return after_launch("maxpool_bwd_kernel");
}

extern "C" int xdet_nchw_f32_to_nhwc_bf16(const float* d_src, void* d_dst, int N, int C, int dst_cs, int HW,
                                          void* stream) {
  if (dst_cs < C) return fail(XDET_EINVAL, "nchw_f32_to_nhwc_bf16: dst_cs < C");
  const long long total = (long long)N * dst_cs * HW;
  if (total <= 0) return XDET_OK;
  nchw_f32_to_nhwc_bf16_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      d_src, reinterpret_cast<__nv_bfloat16*>(d_dst), N, C, dst_cs, HW, total);
  return after_launch("nchw_f32_to_nhwc_bf16_kernel");
}

extern "C" int xdet_affine_relu_to_nchw_f32(const void* d_src, const float* d_scale, const float* d_shift, float* d_dst,
                                            int N, int C, int src_cs, int HW, int relu, void* stream) {
  if (src_cs < C) return fail(XDET_EINVAL, "affine_relu_to_nchw_f32: src_cs < C");
  const long long total = (long long)N * C * HW;
  if (total <= 0) return XDET_OK;
  affine_relu_to_nchw_f32_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_src), d_scale, d_shift, d_dst, N, C, src_cs, HW, relu, total);
  return after_launch("affine_relu_to_nchw_f32_kernel");
}

extern "C" int xdet_softmax_ce(const float* d_logits, int ld, int C, const int* d_labels, const float* d_row_w,
                               float w_all, long long M, float* d_loss_row, float* d_dlogits, int dld, void* stream) {
  if (M <= 0) return XDET_OK;
  if (C <= 0 || ld < C || (d_dlogits && dld < C)) return fail(XDET_EINVAL, "softmax_ce: bad class count / pitch");
  softmax_ce_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_logits, ld, C, d_labels, d_row_w,
                                                                                  w_all, M, d_loss_row, d_dlogits, dld);
  return after_launch("softmax_ce_kernel");
}

extern "C" int xdet_smooth_l1(const float* d_pred, int ld, const float* d_target, const float* d_row_w, float w_all,
                              long long M, float* d_loss_row, float* d_dpred, int dld, void* stream) {
  if (M <= 0) return XDET_OK;
  if (ld < 4 || (d_dpred && dld < 4)) return fail(XDET_EINVAL, "smooth_l1: pitch < 4");
  smooth_l1_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_pred, ld, d_target, d_row_w, w_all, M,
                                                                                 d_loss_row, d_dpred, dld);
  return after_launch("smooth_l1_kernel");
}

extern "C" int xdet_sgd_momentum_conv(const float* d_dw, float* d_w, float* d_mom, void* d_w_pack, void* d_w_dgrad_pack,
                                      int Cout, int KH, int KW, int Cin, int pack_co_off, int pack_ci_off,
                                      int pack_cin_pad, int pack_cout_pad, int fold, float lr, float momentum, float wd,
                                      float grad_scale, void* stream) {
  const long long total = (long long)Cout * KH * KW * Cin;
  if (total <= 0) return XDET_OK;
  if (fold && (KW * 8 > 64 || Cin > 8)) return fail(XDET_EINVAL, "sgd_conv: bad fold geometry");
  const long long taps = (long long)KH * KW;
  // offsets select this variable's slice of a fused (concatenated) pack
  const float* dw = d_dw + ((long long)pack_co_off * taps) * (fold ? 64 : pack_cin_pad) + (fold ? 0 : pack_ci_off);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(d_w_pack) +
                      ((long long)pack_co_off * taps) * (fold ? 64 : pack_cin_pad) + (fold ? 0 : pack_ci_off);
  __nv_bfloat16* wdp = d_w_dgrad_pack ? reinterpret_cast<__nv_bfloat16*>(d_w_dgrad_pack) +
                                            ((long long)pack_ci_off * taps) * pack_cout_pad + pack_co_off
                                      : nullptr;
  if (!fold) {
    dim3 grid((unsigned)((Cin + 31) / 32), (unsigned)((Cout + 31) / 32), (unsigned)taps);
    sgd_conv_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dw, d_w, d_mom, wp, wdp, Cout, (int)taps, Cin,
                                                                 pack_cin_pad, pack_cout_pad, lr, momentum, wd, grad_scale);
    return after_launch("sgd_conv_tiled_kernel");
  }
  sgd_conv_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(dw, d_w, d_mom, wp, wdp, Cout, KH, KW, Cin,
                                                                      pack_cin_pad, pack_cout_pad, fold, lr, momentum, wd,
                                                                      grad_scale, total);
  return after_launch("sgd_conv_kernel");
}

extern "C" int xdet_sgd_momentum_vec(const float* d_g, float* d_w, float* d_mom, long long n, float lr, float momentum,
                                     float wd, float grad_scale, void* stream) {
  if (n <= 0) return XDET_OK;
  sgd_vec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_g, d_w, d_mom, n, lr, momentum, wd,
                                                                               grad_scale);
  return after_launch("sgd_vec_kernel");
}

extern "C" int xdet_sgd_momentum_multi(const xdet_sgd_item* d_items, int n_items, int total_blocks, float lr,
                                       float momentum, float grad_scale, void* stream) {
  if (n_items <= 0 || total_blocks <= 0) return XDET_OK;
  if (!d_items) return fail(XDET_EINVAL, "sgd_multi: no item table");
  sgd_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(d_items, n_items, lr, momentum, grad_scale);
  return after_launch("sgd_multi_kernel");
}

extern "C" size_t xdet_match_workspace_bytes(int N, int G) { return sizeof(unsigned long long) * (size_t)N * (size_t)G; }

extern "C" int xdet_match_encode(const float* d_boxes, long long box_img_stride, const float* d_ref_yxhw,
                                 const float* d_gt, const int* d_gt_labels, int N, int A, int G, float allowed_border,
                                 float high_thres, float low_thres, const float* prior_scaling4, int* d_labels,
                                 float* d_targets, float* d_scores, void* d_workspace, void* stream) {
  if (N <= 0 || A <= 0 || G <= 0) return fail(XDET_EINVAL, "match_encode: empty");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(d_workspace);
  XDET_TRY(check_cuda(cudaMemsetAsync(best, 0, sizeof(unsigned long long) * (size_t)N * G, st), "memset(match)"));
  dim3 grid((unsigned)((A + 255) / 256), (unsigned)N);
  if (box_img_stride != 0 && box_img_stride != (long long)A * 4)
    return fail(XDET_EINVAL, "match_encode: box_img_stride must be 0 or A*4");
  match_gt_argmax_kernel<<<grid, 256, 0, st>>>(d_boxes, box_img_stride, d_gt, d_gt_labels, A, G, allowed_border, best);
  XDET_TRY(after_launch("match_gt_argmax_kernel"));
  match_encode_kernel<<<grid, 256, 0, st>>>(d_boxes, box_img_stride, d_ref_yxhw, d_gt, d_gt_labels, best, A, G,
                                            allowed_border, high_thres, low_thres, prior_scaling4[0], prior_scaling4[1],
                                            prior_scaling4[2], prior_scaling4[3], d_labels, d_targets, d_scores);
  return after_launch("match_encode_kernel");
}

extern "C" int xdet_sample_fg_bg(const int* d_labels, const float* d_scores, float bg_low, int groups, int n, int exp_fg,
                                 int total, const float* d_keys_fg, const float* d_keys_bg, const float* d_keys_up,
                                 int* d_workspace /* [groups][2n] */, int* d_out, int* d_counts, void* stream) {
  if (groups <= 0 || n <= 0 || total <= 0) return fail(XDET_EINVAL, "sample_fg_bg: empty");
  if (total > 4096) return fail(XDET_EINVAL, "sample_fg_bg: at most 4096 samples per group");
  int P = 2;
  while (P < total) P <<= 1;
  const size_t smem = sizeof(unsigned long long) * P + sizeof(int) * 2 * (size_t)total;
  XDET_TRY(check_cuda(cudaFuncSetAttribute(sample_fg_bg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024),
                      "cudaFuncSetAttribute(sample)"));
  sample_fg_bg_kernel<<<groups, kSampThreads, smem, (cudaStream_t)stream>>>(d_labels, d_scores, bg_low, n, exp_fg, total, P,
                                                                           d_keys_fg, d_keys_bg, d_keys_up, d_workspace,
                                                                           d_out, d_counts);
  return after_launch("sample_fg_bg_kernel");
}
