// Weight gradient of the NHWC bf16 convolutions on the 5th-generation tensor cores (sm_100a).
//
//   dW[co, tap, ci] = sum_{n,y,x} dY[n,y,x,co] * X[n, y*s + kh*d - pad_top, x*s + kw*d - pad_left, ci]
//
// Stands in for the Conv2DBackpropFilter kernels TensorFlow's autodiff runs for every tf.layers.conv2d of the
// training graph (light_head_rfcn_train.py:426-441 `optimizer.minimize`; layers net/resnet_v2.py:89-100,
// net/xception_body.py:381-400,450-475).  A GEMM whose reduction axis is the PIXEL axis: M = Cout tile (128),
// N = Cin tile (<= 256), K = 128 output pixels (a BH x BW patch of one image, the forward tiling).  Both
// operands are "MN-major" for tcgen05: in NHWC memory the channels are contiguous and the pixels are the
// strided axis, so the very TMA boxes the forward kernel uses ({64 ch, BW, BH, 1}, 128-byte swizzle, shifted
// by the filter tap, stride-2 through element strides, zero-filled outside the image) land in shared memory as
// [128 pixels][64 channels] = the canonical MN-major SWIZZLE_128B layout
//     ((8,n),(8,k)):((1,LBO),(8,SBO))   (units of 16 bytes; n = 64-channel atoms, k = 8-pixel groups)
// with SBO = 1024 B (8 pixel rows) and LBO = 16 KB (next 64-channel box).  One tcgen05.mma consumes 16 pixels.
//
// Work item = (tap, Cout tile, Cin tile, pixel split); persistent CTAs; the fp32 accumulator leaves TMEM through
// tcgen05.ld and is added to dW with red.global.add.f32 (splits of the same tile meet there; the caller
// zero-fills dW).  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
//
// "f16x2" precision (xdet_conv2d_wgrad_f16x2, the fp32-accurate training mode): both operands arrive as two fp16
// planes (hi, lo*2^11, csrc/conv_gemm_f16x2.cu); a 16-pixel step is three products into two accumulators
// (acc0 += dYh*Xh, acc1 += dYh*Xl + dYl*Xh) and the epilogue adds  (acc0 + acc1*2^-11) * out_scale  to dW.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace xdet {
namespace {

constexpr int kWgThreads = 192;
constexpr int kPix = 128;                      // pixels (K) per stage
constexpr uint32_t kBoxBytes = kPix * 64 * 2;  // one {64 ch x 128 px} box

struct WgradArgs {
  int tiles_x, tiles_y, n_img, pix_tiles;  // pixel tiling of the OUTPUT (dY) grid
  int BW, BH;
  int taps_w, dil_h, dil_w, pad_top, pad_left, mul_x, mul_y;
  int Cout, cin_pad, taps;
  int co_tiles, ci_tiles, BN, a_boxes, b_boxes;
  int splits, items, stages, tmem_cols, overwrite;
  float* dw;
  long long ld_dw;  // taps * cin_pad
  int pair;         // f16x2 operands: every box is loaded twice (hi plane, lo plane)
  float out_scale;  // pair mode: factor applied to the result (undoes a power-of-two scaling of dY)
};

__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);       // start address  [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;  // LBO: next 64-element MN atom
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // SBO: next group of 8 K rows
  d |= static_cast<uint64_t>(1) << 46;                           // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                           // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ CUtensorMap map_dy_lo, const __grid_constant__ CUtensorMap map_x_lo,
                  const WgradArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_bytes = (uint32_t)p.a_boxes * kBoxBytes, b_bytes = (uint32_t)p.b_boxes * kBoxBytes;
  const uint32_t stage_bytes = (a_bytes + b_bytes) * (p.pair ? 2u : 1u);  // pair: [A_hi | B_hi | A_lo | B_lo]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_dy);
    ptx::prefetch_tmap(&map_x);
    if (p.pair) {
      ptx::prefetch_tmap(&map_dy_lo);
      ptx::prefetch_tmap(&map_x_lo);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      ptx::mbar_init(tmem_empty_bar, 4);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // item -> (unit = (tap, co tile, ci tile), split); pixel tiles [pt0, pt1)
  auto decode = [&](int item, int& tap, int& co0, int& ci0, int& pt0, int& pt1) {
    const int s = item % p.splits;
    int u = item / p.splits;
    ci0 = (u % p.ci_tiles) * p.BN;
    u /= p.ci_tiles;
    co0 = (u % p.co_tiles) * 128;
    tap = u / p.co_tiles;
    pt0 = (int)((long long)p.pix_tiles * s / p.splits);
    pt1 = (int)((long long)p.pix_tiles * (s + 1) / p.splits);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int tap, co0, ci0, pt0, pt1;
        decode(item, tap, co0, ci0, pt0, pt1);
        const int kh = tap / p.taps_w, kw = tap - kh * p.taps_w;
        for (int pt = pt0; pt < pt1; ++pt) {
          const int x0 = (pt % p.tiles_x) * p.BW;
          const int y0 = ((pt / p.tiles_x) % p.tiles_y) * p.BH;
          const int img = pt / (p.tiles_x * p.tiles_y);
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + (size_t)stage * stage_bytes;
          unsigned char* sb = sa + a_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          for (int a = 0; a < p.a_boxes; ++a)
            ptx::tma_load_4d(sa + (size_t)a * kBoxBytes, &map_dy, &full_bar[stage], co0 + 64 * a, x0, y0, img);
          const int xx = x0 * p.mul_x + kw * p.dil_w - p.pad_left, yy = y0 * p.mul_y + kh * p.dil_h - p.pad_top;
          for (int b = 0; b < p.b_boxes; ++b)
            ptx::tma_load_4d(sb + (size_t)b * kBoxBytes, &map_x, &full_bar[stage], ci0 + 64 * b, xx, yy, img);
          if (p.pair) {
            unsigned char* sal = sb + b_bytes;
            unsigned char* sbl = sal + a_bytes;
            for (int a = 0; a < p.a_boxes; ++a)
              ptx::tma_load_4d(sal + (size_t)a * kBoxBytes, &map_dy_lo, &full_bar[stage], co0 + 64 * a, x0, y0, img);
            for (int b = 0; b < p.b_boxes; ++b)
              ptx::tma_load_4d(sbl + (size_t)b * kBoxBytes, &map_x_lo, &full_bar[stage], ci0 + 64 * b, xx, yy, img);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // instruction descriptor: D = f32, A = B = bf16, BOTH MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
    // (pair mode: fp16 operands, formats 0)
    const uint32_t fmt = p.pair ? 0u : ((1u << 7) | (1u << 10));
    const uint32_t idesc = (1u << 4) | fmt | (1u << 15) | (1u << 16) |
                           (static_cast<uint32_t>(p.BN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0, aphase = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int tap, co0, ci0, pt0, pt1;
      decode(item, tap, co0, ci0, pt0, pt1);
      ptx::mbar_wait(tmem_empty_bar, aphase ^ 1);
      ptx::tc_fence_after();
      for (int pt = pt0; pt < pt1; ++pt) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < kPix / 16; ++k) {  // 16 pixels = two 8-row groups = 2048 B further down the tile
            const uint64_t da = make_smem_desc_mn_sw128(sa + k * 2048, kBoxBytes);
            const uint64_t db = make_smem_desc_mn_sw128(sb + k * 2048, kBoxBytes);
            const uint32_t acc = (pt > pt0 || k > 0) ? 1u : 0u;
            ptx::mma_bf16_ss(tmem_base, da, db, idesc, acc);
            if (p.pair) {
              const uint32_t sal = sb + b_bytes, sbl = sal + a_bytes;
              const uint64_t dal = make_smem_desc_mn_sw128(sal + k * 2048, kBoxBytes);
              const uint64_t dbl = make_smem_desc_mn_sw128(sbl + k * 2048, kBoxBytes);
              ptx::mma_bf16_ss(tmem_base + (uint32_t)p.BN, da, dbl, idesc, acc);
              ptx::mma_bf16_ss(tmem_base + (uint32_t)p.BN, dal, db, idesc, 1u);
            }
          }
          ptx::mma_commit(&empty_bar[stage]);
          if (pt == pt1 - 1) ptx::mma_commit(tmem_full_bar);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (pt1 > pt0) aphase ^= 1;
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int tap, co0, ci0, pt0, pt1;
      decode(item, tap, co0, ci0, pt0, pt1);
      if (pt1 <= pt0) continue;  // empty split (never with splits <= pix_tiles)
      ptx::mbar_wait(tmem_full_bar, aphase);
      aphase ^= 1;
      ptx::tc_fence_after();
      const int co = co0 + m;
      float* row = p.dw + (long long)co * p.ld_dw + (long long)tap * p.cin_pad + ci0;
      const int ncols = min(p.BN, p.cin_pad - ci0);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
        if (p.pair) {
          uint32_t r1[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(p.BN + c0), r1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            r[j] = __float_as_uint(__fmul_rn(__fmaf_rn(__uint_as_float(r1[j]), 1.f / 2048.f, __uint_as_float(r[j])), p.out_scale));
        }
        ptx::tmem_ld_wait();
        if (co < p.Cout) {  // ncols is a multiple of 64: whole 16-byte groups
          float4* dst = reinterpret_cast<float4*>(row + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                         __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            if (p.splits == 1 && p.overwrite) {
              dst[j] = v;  // the only contribution to this tile: plain vector store
            } else {
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v.x), "f"(v.y),
                           "f"(v.z), "f"(v.w)
                           : "memory");
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

PFN_cuTensorMapEncodeTiled_v12000 wg_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

int wg_encode(CUtensorMap* map, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box, const cuuint32_t* estr) {
  auto fn = wg_encode_fn();
  if (!fn) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return XDET_OK;
}

}  // namespace
}  // namespace xdet

using namespace xdet;

static int launch_wgrad(const void* d_x, const void* d_dy, const xdet_wgrad_desc* d, void* stream, bool pair,
                        long long x_plane, long long dy_plane, float out_scale) {
  if (!d || !d_x || !d_dy || !d->dw) return fail(XDET_EINVAL, "wgrad: null argument");
  if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->KH <= 0 || d->KW <= 0 || d->Hout <= 0 ||
      d->Wout <= 0)
    return fail(XDET_EINVAL, "wgrad: non-positive dimension");
  if (d->in_cs < d->Cin || d->in_cs % 8 || d->dy_cs < d->Cout || d->dy_cs % 8)
    return fail(XDET_EINVAL, "wgrad: channel pitches must cover the channels and be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(d_x) & 15) || (reinterpret_cast<uintptr_t>(d_dy) & 15))
    return fail(XDET_EINVAL, "wgrad: tensors must be 16-byte aligned");
  const int sh = d->stride_h <= 0 ? 1 : d->stride_h, sw = d->stride_w <= 0 ? 1 : d->stride_w;
  if (sh > 2 || sw > 2) return fail(XDET_EINVAL, "wgrad: strides 1 and 2 are supported");
  const int dil_h = d->dil_h <= 0 ? 1 : d->dil_h, dil_w = d->dil_w <= 0 ? 1 : d->dil_w;
  const bool fold = d->fold_w != 0;  // the stem: x is the row-padded NHWC8 image, a K chunk = one filter row
  if (fold && (d->KW * d->in_cs > 64 || dil_w != 1 || d->in_wp < (d->Wout - 1) * sw + 64 / d->in_cs))
    return fail(XDET_EINVAL, "wgrad: fold_w needs KW*in_cs <= 64, dil_w == 1 and in_wp >= (Wout-1)*stride_w + 64/in_cs");

  // 1x1 stride-1 convolutions flatten to one long pixel row (no ragged tiles)
  int N = d->N, H = d->H, W = d->W, Hout = d->Hout, Wout = d->Wout;
  if (!fold && d->KH == 1 && d->KW == 1 && sh == 1 && sw == 1 && d->pad_top == 0 && d->pad_left == 0 && Hout == H && Wout == W &&
      (long long)N * H * W < (1ll << 31)) {
    W = Wout = N * H * W;
    H = Hout = 1;
    N = 1;
  }
  WgradArgs a{};
  int BW = 8;
  while (BW < Wout && BW < kPix) BW <<= 1;
  a.BW = BW;
  a.BH = kPix / BW;
  if (a.BW * sw > 256 || a.BH * sh > 256) return fail(XDET_EINVAL, "wgrad: strided tile exceeds the TMA box limit");
  a.tiles_x = (Wout + a.BW - 1) / a.BW;
  a.tiles_y = (Hout + a.BH - 1) / a.BH;
  a.n_img = N;
  a.pix_tiles = a.tiles_x * a.tiles_y * N;
  a.taps_w = fold ? 1 : d->KW;
  a.taps = fold ? d->KH : d->KH * d->KW;
  a.dil_h = dil_h;
  a.dil_w = fold ? 0 : dil_w;
  a.pad_top = d->pad_top;
  a.pad_left = fold ? 0 : d->pad_left;
  a.mul_x = fold ? 1 : sw;
  a.mul_y = sh;
  a.Cout = d->Cout;
  a.cin_pad = fold ? 64 : (d->Cin + 63) / 64 * 64;
  a.co_tiles = (d->Cout + 127) / 128;
  a.BN = pair ? 64 : (a.cin_pad >= 256 ? 256 : (a.cin_pad >= 128 ? 128 : 64));  // pair: twice the boxes per stage
  a.pair = pair ? 1 : 0;
  a.out_scale = out_scale;
  a.ci_tiles = (a.cin_pad + a.BN - 1) / a.BN;
  a.a_boxes = 2;
  a.b_boxes = a.BN / 64;
  const int units = a.taps * a.co_tiles * a.ci_tiles;
  // pixel splits: the persistent grid runs ceil(items / SMs) rounds of items, an item costs its pixel tiles plus a few
  // tiles' worth of ramp and fp32 reductions (2 + BN/32) -- take the split count with the cheapest rounds x item product (the
  // smallest such count: every extra split adds a full tile of atomics).  E.g. 9 taps of a 64 -> 64 3x3: 16 splits =
  // 144 items in ONE round, where "one item per SM, rounded up" (17 splits, 153 items) ran two.
  int splits = d->splits;
  if (splits <= 0) {
    const int max_splits = std::max(1, std::min(a.pix_tiles / 4, 4 * kNumSMs));
    long long best_cost = -1;
    for (int sp = 1; sp <= max_splits; ++sp) {
      const long long rounds = ((long long)units * sp + kNumSMs - 1) / kNumSMs;
      const long long cost = rounds * ((a.pix_tiles + sp - 1) / sp + 2 + a.BN / 32);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        splits = sp;
      }
    }
  }
  if (splits > a.pix_tiles) splits = a.pix_tiles;
  if (splits < 1) splits = 1;
  a.overwrite = 0;  // dw is accumulated into (a caller-zeroed buffer or a partial sum)
  a.splits = splits;
  a.items = units * splits;
  const size_t stage_bytes = (size_t)(a.a_boxes + a.b_boxes) * kBoxBytes * (pair ? 2 : 1);
  int stages = (int)((227 * 1024 - 1024 - 256) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return fail(XDET_EINVAL, "wgrad: tile does not fit shared memory");
  a.stages = stages;
  a.tmem_cols = pair ? 2 * a.BN : (a.BN < 32 ? 32 : a.BN);  // powers of two; pair: acc0 | acc1
  a.dw = d->dw;
  a.ld_dw = (long long)a.taps * a.cin_pad;

  CUtensorMap map_dy[2], map_x[2];
  for (int pl = 0; pl < (pair ? 2 : 1); ++pl) {
    const unsigned char* dyb = reinterpret_cast<const unsigned char*>(d_dy) + (size_t)pl * dy_plane * 2;
    const unsigned char* xb = reinterpret_cast<const unsigned char*>(d_x) + (size_t)pl * x_plane * 2;
    {
      const cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)Wout, (cuuint64_t)Hout, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)d->dy_cs * 2, (cuuint64_t)d->dy_cs * 2 * Wout,
                                     (cuuint64_t)d->dy_cs * 2 * Wout * Hout};
      const cuuint32_t box[4] = {64, (cuuint32_t)a.BW, (cuuint32_t)a.BH, 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      XDET_TRY(wg_encode(&map_dy[pl], dyb, dims, strides, box, estr));
    }
    if (fold) {
      const cuuint64_t dims[4] = {64, (cuuint64_t)Wout, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)sw * d->in_cs * 2, (cuuint64_t)d->in_wp * d->in_cs * 2,
                                     (cuuint64_t)d->in_wp * d->in_cs * 2 * H};
      const cuuint32_t box[4] = {64, (cuuint32_t)a.BW, (cuuint32_t)(a.BH * sh), 1};
      const cuuint32_t estr[4] = {1, 1, (cuuint32_t)sh, 1};
      XDET_TRY(wg_encode(&map_x[pl], xb, dims, strides, box, estr));
    } else {
      const cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)d->in_cs * 2, (cuuint64_t)d->in_cs * 2 * W,
                                     (cuuint64_t)d->in_cs * 2 * W * H};
      const cuuint32_t box[4] = {64, (cuuint32_t)(a.BW * sw), (cuuint32_t)(a.BH * sh), 1};
      const cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
      XDET_TRY(wg_encode(&map_x[pl], xb, dims, strides, box, estr));
    }
  }
  if (!pair) {
    map_dy[1] = map_dy[0];
    map_x[1] = map_x[0];
  }
  const size_t smem = (size_t)stages * stage_bytes + 256 + 1024;
  XDET_TRY(check_cuda(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                      "cudaFuncSetAttribute(conv_wgrad)"));
  const int grid = a.items < kNumSMs ? a.items : kNumSMs;
  conv_wgrad_kernel<<<grid, kWgThreads, smem, (cudaStream_t)stream>>>(map_dy[0], map_x[0], map_dy[1], map_x[1], a);
  return after_launch("conv_wgrad_kernel");
}

extern "C" int xdet_conv2d_wgrad_bf16(const void* d_x, const void* d_dy, const xdet_wgrad_desc* d, void* stream) {
  return launch_wgrad(d_x, d_dy, d, stream, false, 0, 0, 1.f);
}

extern "C" int xdet_conv2d_wgrad_f16x2(const void* d_x_pair, long long x_plane, const void* d_dy_pair, long long dy_plane,
                                       const xdet_wgrad_desc* d, float out_scale, void* stream) {
  if (x_plane % 8 || dy_plane % 8) return fail(XDET_EINVAL, "wgrad_f16x2: operand planes must be 16-byte aligned");
  return launch_wgrad(d_x_pair, d_dy_pair, d, stream, true, x_plane, dy_plane, out_scale);
}
