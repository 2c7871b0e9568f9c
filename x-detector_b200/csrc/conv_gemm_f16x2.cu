// fp32-ACCURATE implicit-GEMM convolution / GEMM on the tcgen05 tensor cores ("f16x2" precision).
//
// The reference computes every tf.layers.conv2d / dense in fp32 (net/resnet_v2.py:89-100, net/xception_body.py:
// 224-233,381-400,450-475,540-558) and north_star asks for box / score deltas within 1e-4 of it.  bf16 operands
// cannot hold that (~1e-2 per stage), so this kernel is the precision the parity claim is benchmarked at.
//
//   every fp32 operand v is carried as TWO fp16 values      hi = fp16(v),  lo = fp16((v - hi) * 2^11)
//   (v = hi + lo*2^-11 to 2^-22 |v|; weights are pre-scaled per output channel by a power of two so that they sit in
//   fp16's normal range, the inverse scale is folded into the epilogue's `scale` vector), and a product is THREE
//   tensor-core products into TWO fp32 TMEM accumulators:
//       acc0 += A_hi * B_hi                     (magnitude 1)
//       acc1 += A_hi * B_lo + A_lo * B_hi       (magnitude 2^-11 of acc0, carried at 2^11 x its weight)
//   (the lo*lo term is below 2^-22 of the result and is dropped).  Every fp16 x fp16 product is exact in fp32.
//   The tensor core's accumulator TRUNCATES, so the error of acc0 grows with the number of accumulation steps:
//   the reduction is therefore cut into chunks of <= chunk_kb k-blocks (128 reduction elements by default, 8
//   steps); after each chunk the epilogue warps read both accumulators from TMEM and add  acc0 + acc1 * 2^-11  to
//   a running fp32 sum held in REGISTERS with round-to-nearest, while the MMA warp already works on the next chunk
//   in the other TMEM buffer.  One launch per layer, whatever its reduction length.
//
// Same structure as conv_gemm.cu (persistent CTAs, warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue;
// no im2col: a filter tap is the same 4-D TMA box at shifted coordinates, out-of-bounds zero fill = SAME padding),
// with these differences: 4 operand tiles per stage (A_hi, A_lo, B_hi, B_lo), the epilogue groups split the tile's
// COLUMNS (group g owns columns [g*BN/2, (g+1)*BN/2)), outputs are fp32 (any strides: NHWC for the next layer,
// NCHW for PsRoIAlign) written straight from registers, and the epilogue can also emit the f16x2 split of its
// outputs so that the next convolution needs no separate split pass.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cfloat>
#include <mutex>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace xdet {
namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kUmmaK = 16;
constexpr int kThreads = 320;     // TMA, MMA, 8 epilogue warps (2 column groups x 4 TMEM quadrants)
constexpr int kEpiThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kMaxBN = 128;
constexpr float kLoScale = 2048.f;            // 2^11
constexpr float kLoInv = 1.f / 2048.f;
constexpr size_t kSmemCapShared = 211 * 1024;
constexpr size_t kSmemCapAlone = 227 * 1024;

struct PairArgs {
  int tiles_x, tiles_y, n_tiles_n, total_tiles;
  int BW, BH;
  int Hout, Wout, Cout;
  int taps_w, dil_h, dil_w, pad_top, pad_left, mul_x, mul_y;
  int k_chunks_per_tap, num_k_blocks, chunk_kb;
  int BN, stages, tmem_cols;
  const float* scale;
  const float* bias;
  int relu;
  const float* residual;
  float* out;
  long long out_sn, out_sy, out_sx, out_sc;
  __half* out_pair;      // [2][N*Hout*Wout][pair_cs] or null
  long long pair_plane;  // elements between the hi and lo planes of out_pair / out2_pair
  int pair_cs;
  float* out2;
  const float* scale2;
  const float* bias2;
  __half* out2_pair;
  int m_tiles, total_pairs;  // cluster mode: tiles are dealt in pairs of M tiles that share one B tile
  int vec_ok;      // fp32 rows: 2 = 32-byte aligned (256-bit accesses), 1 = 16-byte aligned, 0 = scalar / strided
  int pair_vec32;  // plane rows are 32-byte aligned
};

struct TileCoord {
  int x0, y0, img, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const PairArgs& p, int t) {
  TileCoord c;
  const int nt = t % p.n_tiles_n;
  int mt = t / p.n_tiles_n;
  c.n0 = nt * p.BN;
  c.x0 = (mt % p.tiles_x) * p.BW;
  mt /= p.tiles_x;
  c.y0 = (mt % p.tiles_y) * p.BH;
  c.img = mt / p.tiles_y;
  return c;
}

// Cluster mode (2 CTAs): tile pair tp -> N tile tp % n_tiles_n and the M tiles 2*(tp / n_tiles_n) + {0, 1}; CTA `rank`
// takes the rank-th of them.  An odd M-tile count leaves a phantom second tile: that CTA recomputes the last real
// tile (identical values, benign duplicate stores) so that the pair stays in lockstep.
__device__ __forceinline__ int pair_tile(const PairArgs& p, int tp, int rank) {
  const int nt = tp % p.n_tiles_n;
  int mt = 2 * (tp / p.n_tiles_n) + rank;
  if (mt >= p.m_tiles) mt = p.m_tiles - 1;
  return mt * p.n_tiles_n + nt;
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// kind::f16 instruction descriptor with fp16 operands (a_format = b_format = 0), fp32 accumulator
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// v -> (hi, lo) of the f16x2 representation; saturating conversions keep out-of-range values finite
__device__ __forceinline__ void split_f16x2(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  const float r = __fsub_rn(v, __half2float(hi));  // exact for |v| <= 65504
  lo = __float2half_rn(fminf(fmaxf(__fmul_rn(r, kLoScale), -65504.f), 65504.f));
}

// 256-bit global accesses (sm_100: LDG/STG.256): a lane moves a whole 32-byte sector per instruction
__device__ __forceinline__ void st_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void st_v8(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_v8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// 16 consecutive channels of one pixel: fp32 row (vec: 2 = 32-byte, 1 = 16-byte aligned rows, 0 = scalar / strided)
__device__ __forceinline__ void store_row16(float* op, long long sc, const float* v, int nv, int vec) {
  if (vec == 2 && nv == 16) {
    st_v8(op, v);
    st_v8(op + 8, v + 8);
  } else if (vec >= 1 && nv == 16) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      reinterpret_cast<float4*>(op)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nv) op[(long long)j * sc] = v[j];
  }
}
// ... and their f16x2 planes (channels >= nv as zeros; units beyond the row's pitch are not written)
__device__ __forceinline__ void store_pair16(__half* hp, long long plane, const float* v, int nv, int units, int vec32) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __half h0, l0, h1, l1;
    split_f16x2(2 * j < nv ? v[2 * j] : 0.f, h0, l0);
    split_f16x2(2 * j + 1 < nv ? v[2 * j + 1] : 0.f, h1, l1);
    hi[j] = pack_h2(h0, h1);
    lo[j] = pack_h2(l0, l1);
  }
  __half* lp = hp + plane;
  if (vec32 && units == 2) {
    st_v8(hp, hi);
    st_v8(lp, lo);
  } else {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q < units) {
        reinterpret_cast<uint4*>(hp)[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
        reinterpret_cast<uint4*>(lp)[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
      }
    }
  }
}

// CG = columns per epilogue group (BN = 2*CG).  CL: the CTAs run as clusters of two that work on two M tiles of the same
// N tile and fill each other's B tiles -- each loads HALF of B_hi / B_lo and multicasts it to both (a quarter less
// operand traffic L2 -> SM, which is what bounds the 30^2 / 60^2 layers); a stage is free when BOTH have consumed it.
template <int CG, bool CL>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_f16x2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                       const PairArgs p) {
  constexpr int BN = 2 * CG;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr uint32_t a_bytes = kBM * kBK * 2;
  constexpr uint32_t b_bytes = (uint32_t)BN * kBK * 2;
  constexpr uint32_t b_pad = (b_bytes + 1023u) & ~1023u;
  constexpr uint32_t stage_bytes = 2 * a_bytes + 2 * b_pad;
  unsigned char* tiles = smem;
  float* sbuf = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes);  // [4][kMaxBN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbuf + 4 * kMaxBN);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL ? (int)ptx::cluster_ctarank() : 0;
  // the sequence of tiles of this CTA: first, stride, count
  const int t_first = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_stride = CL ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int t_count = CL ? p.total_pairs : p.total_tiles;
  auto tile_of = [&](int i) { return CL ? pair_tile(p, i, rank) : i; };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_a_hi);
    ptx::prefetch_tmap(&map_a_lo);
    ptx::prefetch_tmap(&map_b_hi);
    ptx::prefetch_tmap(&map_b_lo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], CL ? 2 : 1);  // cluster: the MMA warps of both CTAs release a stage
      }
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&tmem_full_bar[s], 1);
        ptx::mbar_init(&tmem_empty_bar[s], 8);  // one arrival per epilogue warp
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  if (CL) ptx::cluster_sync();  // the peer's barriers exist before anything is multicast to them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();

  const int nchunks = (p.num_k_blocks + p.chunk_kb - 1) / p.chunk_kb;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = t_first; ti < t_count; ti += t_stride) {
        const TileCoord tc = decode_tile(p, tile_of(ti));
        const int ax = tc.x0 * p.mul_x - p.pad_left, ay = tc.y0 * p.mul_y - p.pad_top;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          const int tap = kb / p.k_chunks_per_tap, cc = kb - tap * p.k_chunks_per_tap;
          const int kh = tap / p.taps_w, kw = tap - kh * p.taps_w;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* s_ahi = tiles + (size_t)stage * stage_bytes;
          unsigned char* s_alo = s_ahi + a_bytes;
          unsigned char* s_bhi = s_alo + a_bytes;
          unsigned char* s_blo = s_bhi + b_pad;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * a_bytes + 2 * b_bytes);
          const int cx = ax + kw * p.dil_w, cy = ay + kh * p.dil_h;
          ptx::tma_load_4d(s_ahi, &map_a_hi, &full_bar[stage], cc * kBK, cx, cy, tc.img);
          ptx::tma_load_4d(s_alo, &map_a_lo, &full_bar[stage], cc * kBK, cx, cy, tc.img);
          if (CL) {  // my half of the B rows, into both CTAs
            constexpr uint32_t half = (uint32_t)(BN / 2) * kBK * 2;
            ptx::tma_load_2d_multicast(s_bhi + rank * half, &map_b_hi, &full_bar[stage], kb * kBK, tc.n0 + rank * (BN / 2),
                                       (uint16_t)3);
            ptx::tma_load_2d_multicast(s_blo + rank * half, &map_b_lo, &full_bar[stage], kb * kBK, tc.n0 + rank * (BN / 2),
                                       (uint16_t)3);
          } else {
            ptx::tma_load_2d(s_bhi, &map_b_hi, &full_bar[stage], kb * kBK, tc.n0);
            ptx::tma_load_2d(s_blo, &map_b_lo, &full_bar[stage], kb * kBK, tc.n0);
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
    // ===== MMA issuer: one accumulator pair per CHUNK of the reduction =====
    constexpr uint32_t idesc = make_idesc_f16(kBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t g = 0;  // running chunk index: TMEM buffer g & 1, barrier parity (g >> 1) & 1
    for (int ti = t_first; ti < t_count; ti += t_stride) {
      for (int kb0 = 0; kb0 < p.num_k_blocks; kb0 += p.chunk_kb, ++g) {
        const int kb1 = min(kb0 + p.chunk_kb, p.num_k_blocks);
        const uint32_t buf = g & 1u;
        ptx::mbar_wait(&tmem_empty_bar[buf], ((g >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + buf * (2u * BN);
        const uint32_t d1 = d0 + (uint32_t)BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t s_ahi = ptx::smem_u32(tiles + (size_t)stage * stage_bytes);
            const uint32_t s_alo = s_ahi + a_bytes;
            const uint32_t s_bhi = s_alo + a_bytes;
            const uint32_t s_blo = s_bhi + b_pad;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t dah = ptx::make_smem_desc_sw128(s_ahi + k * kUmmaK * 2);
              const uint64_t dal = ptx::make_smem_desc_sw128(s_alo + k * kUmmaK * 2);
              const uint64_t dbh = ptx::make_smem_desc_sw128(s_bhi + k * kUmmaK * 2);
              const uint64_t dbl = ptx::make_smem_desc_sw128(s_blo + k * kUmmaK * 2);
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              ptx::mma_bf16_ss(d0, dah, dbh, idesc, acc);
              ptx::mma_bf16_ss(d1, dah, dbl, idesc, acc);
              ptx::mma_bf16_ss(d1, dal, dbh, idesc, 1u);
            }
            if (CL) ptx::mma_commit_multicast(&empty_bar[stage], (uint16_t)3);
            else ptx::mma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) ptx::mma_commit(&tmem_full_bar[buf]);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===== epilogue: warps 2..5 = column group 0, warps 6..9 = column group 1; TMEM quadrant = warp % 4 =====
    const int grp = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int m = quad * 32 + lane;  // row of the tile == TMEM lane
    const int e = threadIdx.x - 64;  // 0..255 among the epilogue threads
    uint32_t g = 0;
    float sum[CG];
    for (int ti = t_first; ti < t_count; ti += t_stride) {
      const TileCoord tc = decode_tile(p, tile_of(ti));
      // per-tile scale / bias vectors -> shared memory (every epilogue thread is past the previous tile's reads)
      ptx::named_bar_sync(1, kEpiThreads);
      for (int i = e; i < BN; i += kEpiThreads) {
        const int c = tc.n0 + i;
        const bool ok = c < p.Cout;
        sbuf[i] = (ok && p.scale) ? __ldg(p.scale + c) : 1.f;
        sbuf[kMaxBN + i] = (ok && p.bias) ? __ldg(p.bias + c) : 0.f;
        sbuf[2 * kMaxBN + i] = (ok && p.scale2) ? __ldg(p.scale2 + c) : 1.f;
        sbuf[3 * kMaxBN + i] = (ok && p.bias2) ? __ldg(p.bias2 + c) : 0.f;
      }
      ptx::named_bar_sync(1, kEpiThreads);

      // The residual does not depend on the accumulators: prefetch this thread's row of it BEFORE waiting for the
      // tile's last chunk, so that its DRAM latency hides behind the MMAs (8 epilogue warps alone cannot keep enough
      // loads in flight otherwise; holding the row in registers instead spilled and was slower).
      const int py = tc.y0 + m / p.BW, px = tc.x0 + m % p.BW;
      const bool row_ok = py < p.Hout && px < p.Wout;
      const long long pix_off = (long long)tc.img * p.out_sn + (long long)py * p.out_sy + (long long)px * p.out_sx;
      const bool res_pref = p.residual != nullptr && p.vec_ok == 2 && row_ok && tc.n0 + grp * CG + CG <= p.Cout;
      for (int ch = 0; ch < nchunks; ++ch, ++g) {
        if (ch == nchunks - 1 && res_pref) {  // pull this thread's residual row towards L2 / L1 (no registers held)
          const float* rp = p.residual + pix_off + (long long)(tc.n0 + grp * CG) * p.out_sc;
#pragma unroll
          for (int q = 0; q < CG / 32 + (CG % 32 ? 1 : 0); ++q)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rp + 32 * q));
        }
        const uint32_t buf = g & 1u;
        ptx::mbar_wait(&tmem_full_bar[buf], (g >> 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (2u * BN) + (uint32_t)(grp * CG);
#pragma unroll
        for (int u = 0; u < CG / 16; ++u) {
          uint32_t r0[16], r1[16];
          tmem_ld_32x16(t0 + (uint32_t)(u * 16), r0);
          tmem_ld_32x16(t0 + (uint32_t)(BN + u * 16), r1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float part = __fmaf_rn(__uint_as_float(r1[j]), kLoInv, __uint_as_float(r0[j]));
            sum[u * 16 + j] = (ch == 0) ? part : __fadd_rn(sum[u * 16 + j], part);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[buf]);
      }

      // ---- the tile's outputs, straight from registers ----
      if (row_ok) {
        const long long pix_lin = ((long long)tc.img * p.Hout + py) * p.Wout + px;
#pragma unroll
        for (int u = 0; u < CG / 16; ++u) {
          const int cl = grp * CG + u * 16;  // column inside the tile
          const int c0 = tc.n0 + cl;         // output channel
          if (c0 >= p.Cout) continue;
          const int nv = min(16, p.Cout - c0);
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __fmaf_rn(sum[u * 16 + j], sbuf[cl + j], sbuf[kMaxBN + cl + j]);
          if (p.residual) {
            const float* rp = p.residual + pix_off + (long long)c0 * p.out_sc;
            if (p.vec_ok == 2 && nv == 16) {
              float r[16];
              ld_v8(rp, r);
              ld_v8(rp + 8, r + 8);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __fadd_rn(v[j], r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < nv) v[j] = __fadd_rn(v[j], __ldg(rp + (long long)j * p.out_sc));
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          const int units = min(2, (p.pair_cs - c0) / 8);
          if (p.out) store_row16(p.out + pix_off + (long long)c0 * p.out_sc, p.out_sc, v, nv, p.vec_ok);
          if (p.out_pair) store_pair16(p.out_pair + pix_lin * p.pair_cs + c0, p.pair_plane, v, nv, units, p.pair_vec32);
          if (p.out2 || p.out2_pair) {
            float w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              w[j] = fmaxf(__fmaf_rn(v[j], sbuf[2 * kMaxBN + cl + j], sbuf[3 * kMaxBN + cl + j]), 0.f);
            if (p.out2) store_row16(p.out2 + pix_off + (long long)c0 * p.out_sc, p.out_sc, w, nv, p.vec_ok);
            if (p.out2_pair)
              store_pair16(p.out2_pair + pix_lin * p.pair_cs + c0, p.pair_plane, w, nv, units, p.pair_vec32);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  if (CL) ptx::cluster_sync();  // nobody leaves while the peer may still signal its barriers / fill its tiles
  else __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// fp32 (any layout) -> f16x2 planes [2][N][H][Wp][cs]: pixel (n,y,x) at ((n*H + y)*Wp + x + x_off)*cs.  Only the
// interior is written (the caller zero-fills padded layouts once).
__global__ void __launch_bounds__(256) split2_kernel(const float* __restrict__ src, long long sn, long long sy,
                                                     long long sx, long long sc, int H, int W, int C,
                                                     __half* __restrict__ dst, int cs, int Wp, int x_off,
                                                     long long plane, int relu, long long total) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  const long long step = (long long)gridDim.x * blockDim.x;
  const int c8 = cs / 8;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int cb = (int)(e % c8) * 8;
    const long long pix = e / c8;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    const float* s = src + n * sn + y * sy + x * sx + (long long)cb * sc;
    __half hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = (cb + j < C) ? __ldg(s + (long long)j * sc) : 0.f;
      if (relu) v = fmaxf(v, 0.f);
      split_f16x2(v, hi[j], lo[j]);
    }
    __half* d = dst + ((n * H + y) * Wp + x + x_off) * cs + cb;
    *reinterpret_cast<uint4*>(d) = make_uint4(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]), pack_h2(hi[4], hi[5]), pack_h2(hi[6], hi[7]));
    *reinterpret_cast<uint4*>(d + plane) = make_uint4(pack_h2(lo[0], lo[1]), pack_h2(lo[2], lo[3]), pack_h2(lo[4], lo[5]), pack_h2(lo[6], lo[7]));
  }
}

// tf.layers.max_pooling2d(3, 2, 'SAME') on NHWC fp32 (+ residual), 8 channels per thread (two float4 per tap):
//   v = max over the window (+ residual);  dst = v (fp32, optional), dst_pair = split(v) (optional);
//   w = ReLU(v*scale2 + bias2):  dst2 (fp32, optional), dst2_pair (optional).
__global__ void __launch_bounds__(256) maxpool3x3s2_f32x_kernel(
    const float* __restrict__ src, float* __restrict__ dst, __half* __restrict__ dst_pair, float* __restrict__ dst2,
    __half* __restrict__ dst2_pair, const float* __restrict__ scale2, const float* __restrict__ bias2,
    const float* __restrict__ residual, long long plane, int H, int W, int C, int Ho, int Wo, int pad_top, int pad_left,
    long long total) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  const long long step = (long long)gridDim.x * blockDim.x;
  const int c8 = C / 8;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
    const int cb = (int)(e % c8) * 8;
    const long long pix = e / c8;
    const int xo = (int)(pix % Wo);
    const int yo = (int)((pix / Wo) % Ho);
    const long long n = pix / ((long long)Wo * Ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -FLT_MAX;
    // all nine taps are loaded from clamped positions first and masked afterwards: a `continue` per tap made every load
    // wait for the maximum over the one before it (nine dependent round trips per output)
    float4 ta[9], tb[9];
    bool ok[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yi = yo * 2 + kh - pad_top;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int xi = xo * 2 + kw - pad_left;
        ok[kh * 3 + kw] = yi >= 0 && yi < H && xi >= 0 && xi < W;
        const int yc = min(max(yi, 0), H - 1), xc = min(max(xi, 0), W - 1);
        const float4* p = reinterpret_cast<const float4*>(src + ((n * H + yc) * W + xc) * C + cb);
        ta[kh * 3 + kw] = __ldg(p);
        tb[kh * 3 + kw] = __ldg(p + 1);
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (ok[t]) {
        m[0] = fmaxf(m[0], ta[t].x); m[1] = fmaxf(m[1], ta[t].y); m[2] = fmaxf(m[2], ta[t].z); m[3] = fmaxf(m[3], ta[t].w);
        m[4] = fmaxf(m[4], tb[t].x); m[5] = fmaxf(m[5], tb[t].y); m[6] = fmaxf(m[6], tb[t].z); m[7] = fmaxf(m[7], tb[t].w);
      }
    }
    const long long o = pix * C + cb;
    if (residual) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(residual + o)), b = __ldg(reinterpret_cast<const float4*>(residual + o) + 1);
      m[0] = __fadd_rn(m[0], a.x); m[1] = __fadd_rn(m[1], a.y); m[2] = __fadd_rn(m[2], a.z); m[3] = __fadd_rn(m[3], a.w);
      m[4] = __fadd_rn(m[4], b.x); m[5] = __fadd_rn(m[5], b.y); m[6] = __fadd_rn(m[6], b.z); m[7] = __fadd_rn(m[7], b.w);
    }
    auto put_pair = [&](__half* base, const float* v) {
      __half hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_f16x2(v[j], hi[j], lo[j]);
      *reinterpret_cast<uint4*>(base + o) = make_uint4(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]), pack_h2(hi[4], hi[5]), pack_h2(hi[6], hi[7]));
      *reinterpret_cast<uint4*>(base + plane + o) = make_uint4(pack_h2(lo[0], lo[1]), pack_h2(lo[2], lo[3]), pack_h2(lo[4], lo[5]), pack_h2(lo[6], lo[7]));
    };
    if (dst) {
      reinterpret_cast<float4*>(dst + o)[0] = make_float4(m[0], m[1], m[2], m[3]);
      reinterpret_cast<float4*>(dst + o)[1] = make_float4(m[4], m[5], m[6], m[7]);
    }
    if (dst_pair) put_pair(dst_pair, m);
    if (dst2 || dst2_pair) {
      float w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = fmaxf(__fadd_rn(__fmul_rn(m[j], __ldg(scale2 + cb + j)), __ldg(bias2 + cb + j)), 0.f);
      if (dst2) {
        reinterpret_cast<float4*>(dst2 + o)[0] = make_float4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<float4*>(dst2 + o)[1] = make_float4(w[4], w[5], w[6], w[7]);
      }
      if (dst2_pair) put_pair(dst2_pair, w);
    }
  }
}

// Depthwise 3x3 'SAME' stride-1 convolution in fp32 (depth multiplier 1, dilation D), the depthwise half of
// tf.layers.separable_conv2d (net/xception_body.py:224-233) for the f16x2 precision.  A thread owns 4 channels of one
// column and walks a vertical strip of YS output rows with a rolling window of 2D+1 input rows x 3 columns in
// registers: 3 float4 loads per output row instead of 9.  Taps are summed in (kh, kw) order with fmaf.  The result
// goes out as fp32 and / or as the f16x2 planes the pointwise convolution reads (its only consumer in XceptionBody).
template <int D>
__global__ void __launch_bounds__(256) depthwise3x3_f32x_kernel(const float* __restrict__ src, const float* __restrict__ w9c,
                                                                float* __restrict__ dst, __half* __restrict__ dst_pair,
                                                                long long plane, int N, int H, int W, int C, int relu_in,
                                                                int YS, long long total) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  constexpr int R = 2 * D + 1;
  const int C4 = C / 4;
  const int HY = (H + YS - 1) / YS;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(e % C4);
    long long t = e / C4;
    const int x = (int)(t % W);
    t /= W;
    const int ys = (int)(t % HY);
    const int n = (int)(t / HY);
    const int y0 = ys * YS, y1 = min(y0 + YS, H);
    float4 wt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wt[k] = __ldg(reinterpret_cast<const float4*>(w9c + (long long)k * C) + c4);
    const float4* img = reinterpret_cast<const float4*>(src) + (long long)n * H * W * C4 + c4;
    const bool okl = x - D >= 0, okr = x + D < W;
    auto load_row = [&](int yi, float4 (&r)[3]) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yi < 0 || yi >= H) {
        r[0] = r[1] = r[2] = z;
        return;
      }
      const float4* rp = img + (long long)yi * W * C4;
      r[0] = okl ? __ldg(rp + (long long)(x - D) * C4) : z;
      r[1] = __ldg(rp + (long long)x * C4);
      r[2] = okr ? __ldg(rp + (long long)(x + D) * C4) : z;
      if (relu_in) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          r[i] = make_float4(fmaxf(r[i].x, 0.f), fmaxf(r[i].y, 0.f), fmaxf(r[i].z, 0.f), fmaxf(r[i].w, 0.f));
      }
    };
    float4 win[R][3];  // input rows y-D .. y+D of the current output row y
#pragma unroll
    for (int i = 0; i < R - 1; ++i) load_row(y0 - D + i, win[i + 1]);  // slot i+1: shifted down at the top of the loop
    for (int y = y0; y < y1; ++y) {
#pragma unroll
      for (int i = 0; i < R - 1; ++i) {
        win[i][0] = win[i + 1][0];
        win[i][1] = win[i + 1][1];
        win[i][2] = win[i + 1][2];
      }
      load_row(y + D, win[R - 1]);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 v = win[kh * D][kw], k = wt[kh * 3 + kw];
          a.x = __fmaf_rn(v.x, k.x, a.x);
          a.y = __fmaf_rn(v.y, k.y, a.y);
          a.z = __fmaf_rn(v.z, k.z, a.z);
          a.w = __fmaf_rn(v.w, k.w, a.w);
        }
      }
      const long long o = (((long long)n * H + y) * W + x) * C + (long long)c4 * 4;
      if (dst) *reinterpret_cast<float4*>(dst + o) = a;
      if (dst_pair) {
        __half h0, l0, h1, l1, h2, l2, h3, l3;
        split_f16x2(a.x, h0, l0);
        split_f16x2(a.y, h1, l1);
        split_f16x2(a.z, h2, l2);
        split_f16x2(a.w, h3, l3);
        *reinterpret_cast<uint2*>(dst_pair + o) = make_uint2(pack_h2(h0, h1), pack_h2(h2, h3));
        *reinterpret_cast<uint2*>(dst_pair + plane + o) = make_uint2(pack_h2(l0, l1), pack_h2(l2, l3));
      }
    }
  }
}

// The same depthwise convolution with the input staged by TMA (what the Xception paths run): a persistent CTA walks
// tiles of TH x TW pixels x 32 channels; one 4-D box {32 ch, TW+2D, TH+2D, 1} per tile lands in shared memory with the
// halo and TMA's zero fill outside the image (= 'SAME' padding), kDwStages tiles ahead of the arithmetic.  The register
// kernel above keeps its in-flight loads in registers -- 3 x 16 B per thread at 25 % occupancy, ~24 KB per SM, which by
// Little's law is ~2.6 TB/s of requests against a 3x re-read of every input --; here the in-flight bytes live in shared
// memory (3 boxes, ~70 KB per SM) and the halo costs 1.4x.  A thread owns one float4 of channels for 4 pixels of the
// tile; a quarter-warp reads 128 contiguous bytes of one pixel (no bank conflicts).  Same taps, same fmaf order.
constexpr int kDwTH = 8, kDwTW = 16, kDwCB = 32, kDwStages = 4, kDwThreads = 256;

struct DwTile {
  int n, y0, x0, c0;
};

template <int D>
__global__ void __launch_bounds__(kDwThreads, 2) depthwise3x3_f32_tma_kernel(
    const __grid_constant__ CUtensorMap map_src, const float* __restrict__ w9c, float* __restrict__ dst,
    __half* __restrict__ dst_pair, long long plane, int N, int H, int W, int C, int relu_in, int tiles_x, int tiles_y,
    int tiles_c, int total_tiles) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a PDL-launched convolution may come up behind us
  constexpr int HW_ = kDwTW + 2 * D, HH_ = kDwTH + 2 * D;
  constexpr uint32_t kBox = (uint32_t)HW_ * HH_ * kDwCB * 4;
  extern __shared__ unsigned char dw_smem_raw[];
  unsigned char* smem = dw_smem_raw + ((128u - (ptx::smem_u32(dw_smem_raw) & 127u)) & 127u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)kDwStages * kBox);
  const int tid = threadIdx.x;
  if (tid == 0) {
    ptx::prefetch_tmap(&map_src);
    for (int s_ = 0; s_ < kDwStages; ++s_) ptx::mbar_init(&full_bar[s_], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  auto decode = [&](int t) {
    DwTile d;
    d.c0 = (t % tiles_c) * kDwCB;
    t /= tiles_c;
    d.x0 = (t % tiles_x) * kDwTW;
    t /= tiles_x;
    d.y0 = (t % tiles_y) * kDwTH;
    d.n = t / tiles_y;
    return d;
  };
  auto issue = [&](int t, int stage) {
    const DwTile d = decode(t);
    ptx::mbar_arrive_expect_tx(&full_bar[stage], kBox);
    ptx::tma_load_4d(smem + (size_t)stage * kBox, &map_src, &full_bar[stage], d.c0, d.x0 - D, d.y0 - D, d.n);
  };
  const int first = blockIdx.x, stride = gridDim.x;
  if (tid == 0)
    for (int s_ = 0; s_ < kDwStages - 1; ++s_)
      if (first + s_ * stride < total_tiles) issue(first + s_ * stride, s_);

  const int cv = tid & 7, p0 = tid >> 3;  // float4 of channels inside the 32-channel box; first of this thread's 4 pixels
  int it = 0;
  for (int t = first; t < total_tiles; t += stride, ++it) {
    const int stage = it % kDwStages;
    // the stage that tile it + kDwStages - 1 will use was read in iteration it - 1: every thread is past it (barrier below)
    if (tid == 0) {
      const int tn = t + (kDwStages - 1) * stride;
      if (tn < total_tiles) issue(tn, (it + kDwStages - 1) % kDwStages);
    }
    const DwTile d = decode(t);
    const int c = d.c0 + cv * 4;
    float4 wt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k)
      wt[k] = c < C ? __ldg(reinterpret_cast<const float4*>(w9c + (long long)k * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    ptx::mbar_wait(&full_bar[stage], (uint32_t)(it / kDwStages) & 1u);
    const float4* tile = reinterpret_cast<const float4*>(smem + (size_t)stage * kBox);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int pix = p0 + 32 * q, py = pix / kDwTW, px = pix % kDwTW;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float4 v = tile[((py + kh * D) * HW_ + (px + kw * D)) * (kDwCB / 4) + cv];
          if (relu_in) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
          const float4 k = wt[kh * 3 + kw];
          a.x = __fmaf_rn(v.x, k.x, a.x);
          a.y = __fmaf_rn(v.y, k.y, a.y);
          a.z = __fmaf_rn(v.z, k.z, a.z);
          a.w = __fmaf_rn(v.w, k.w, a.w);
        }
      }
      const int y = d.y0 + py, x = d.x0 + px;
      if (y < H && x < W && c < C) {
        const long long o = (((long long)d.n * H + y) * W + x) * C + c;
        if (dst) *reinterpret_cast<float4*>(dst + o) = a;
        if (dst_pair) {
          __half h0, l0, h1, l1, h2, l2, h3, l3;
          split_f16x2(a.x, h0, l0);
          split_f16x2(a.y, h1, l1);
          split_f16x2(a.z, h2, l2);
          split_f16x2(a.w, h3, l3);
          *reinterpret_cast<uint2*>(dst_pair + o) = make_uint2(pack_h2(h0, h1), pack_h2(h2, h3));
          *reinterpret_cast<uint2*>(dst_pair + plane + o) = make_uint2(pack_h2(l0, l1), pack_h2(l2, l3));
        }
      }
    }
    __syncthreads();  // every thread has read this stage: it may be refilled (by the issue at the top of the next turn)
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
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

int encode_f16(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* estr = nullptr) {
  auto fn = encode_fn();
  if (!fn) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                  box, estr ? estr : ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return XDET_OK;
}

// N tile: 32 / 64 / 128.  Cost model as in conv_gemm.cu, with three products per k-block and twice the operand bytes.
int pick_bn(int cout, long long m_tiles, int num_k_blocks) {
  const int cands[3] = {128, 64, 32};
  int best = 0;
  double best_cost = 0;
  for (int bn : cands) {
    const long long tiles = m_tiles * ((cout + bn - 1) / bn);
    const long long waves = (tiles + kNumSMs - 1) / kNumSMs;
    const double per_kb = std::max(6.0 * bn, (32768.0 + bn * 256.0) / 56.0);
    const double cost = (double)waves * (num_k_blocks * per_kb + 8.0 * bn + 800.0);
    if (best == 0 || cost < best_cost * 0.98) {
      best = bn;
      best_cost = cost;
    }
  }
  return best;
}

template <int CG, bool CL>
cudaError_t launch(const cudaLaunchConfig_t& cfg, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                   const CUtensorMap& b_hi, const CUtensorMap& b_lo, const PairArgs& a) {
  cudaError_t e = cudaFuncSetAttribute(conv_gemm_f16x2_kernel<CG, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaLaunchKernelEx(&cfg, conv_gemm_f16x2_kernel<CG, CL>, a_hi, a_lo, b_hi, b_lo, a);
}

}  // namespace
}  // namespace xdet

using namespace xdet;

extern "C" int xdet_conv2d_f16x2(const void* d_in_pair, const xdet_conv_f16x2_desc* d, void* stream) {
  if (!d || !d_in_pair) return fail(XDET_EINVAL, "null argument");
  if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->KH <= 0 || d->KW <= 0)
    return fail(XDET_EINVAL, "conv2d_f16x2: non-positive dimension");
  if (d->Hout <= 0 || d->Wout <= 0) return fail(XDET_EINVAL, "conv2d_f16x2: non-positive output size");
  if (d->in_cs < d->Cin || (d->in_cs % 8) != 0)
    return fail(XDET_EINVAL, "conv2d_f16x2: input channel stride (%d) must be >= Cin and a multiple of 8", d->in_cs);
  if ((reinterpret_cast<uintptr_t>(d_in_pair) & 15) || (reinterpret_cast<uintptr_t>(d->weights) & 15) ||
      (d->in_plane % 8) != 0 || (d->w_plane % 8) != 0)
    return fail(XDET_EINVAL, "conv2d_f16x2: operand planes must be 16-byte aligned");
  const int sh = d->stride_h <= 0 ? 1 : d->stride_h, sw = d->stride_w <= 0 ? 1 : d->stride_w;
  if (sh > 2 || sw > 2) return fail(XDET_EINVAL, "conv2d_f16x2: strides 1 and 2 are supported");
  const int dil_h = d->dil_h <= 0 ? 1 : d->dil_h, dil_w = d->dil_w <= 0 ? 1 : d->dil_w;
  const bool fold = d->fold_w != 0;
  if (fold && (d->KW * d->in_cs > kBK || dil_w != 1 || d->in_wp < (d->Wout - 1) * sw + kBK / d->in_cs))
    return fail(XDET_EINVAL, "conv2d_f16x2: fold_w needs KW*in_cs <= 64, dil_w == 1 and in_wp >= (Wout-1)*stride_w + 64/in_cs");
  if (!d->out && !d->out_pair && !d->out2 && !d->out2_pair) return fail(XDET_EINVAL, "conv2d_f16x2: no output");
  if ((d->out2 || d->out2_pair) && (!d->scale2 || !d->bias2))
    return fail(XDET_EINVAL, "conv2d_f16x2: second output needs scale2 and bias2");
  if ((d->out_pair || d->out2_pair) && (d->pair_cs < d->Cout || d->pair_cs % 8 != 0 || d->pair_plane % 8 != 0))
    return fail(XDET_EINVAL, "conv2d_f16x2: pair_cs must be >= Cout and a multiple of 8");

  int N = d->N, H = d->H, W = d->W, Hout = d->Hout, Wout = d->Wout;
  long long out_sn = d->out_sn, out_sy = d->out_sy, out_sx = d->out_sx;
  const bool pointwise = !fold && d->KH == 1 && d->KW == 1 && sh == 1 && sw == 1 && d->pad_top == 0 && d->pad_left == 0 &&
                         Hout == H && Wout == W;
  if (pointwise && out_sy == out_sx * W && out_sn == out_sy * H && (long long)N * H * W < (1ll << 31)) {
    W = Wout = N * H * W;
    H = Hout = 1;
    N = 1;
    out_sy = out_sn = out_sx * W;
  }
  const int kcpt = fold ? 1 : (d->Cin + kBK - 1) / kBK;
  const int taps = fold ? d->KH : d->KH * d->KW;
  const int ktot = taps * kcpt * kBK;

  PairArgs a{};
  int BW = 8;
  while (BW < Wout && BW < kBM) BW <<= 1;
  a.BW = BW;
  a.BH = kBM / BW;
  a.tiles_x = (Wout + a.BW - 1) / a.BW;
  a.tiles_y = (Hout + a.BH - 1) / a.BH;
  const long long m_tiles = (long long)a.tiles_x * a.tiles_y * N;
  a.Hout = Hout;
  a.Wout = Wout;
  a.Cout = d->Cout;
  a.taps_w = fold ? 1 : d->KW;
  a.dil_h = dil_h;
  a.dil_w = fold ? 0 : dil_w;
  a.pad_top = d->pad_top;
  a.pad_left = fold ? 0 : d->pad_left;
  a.mul_x = fold ? 1 : sw;
  a.mul_y = sh;
  a.k_chunks_per_tap = kcpt;
  a.num_k_blocks = taps * kcpt;
  // equal chunks of at most chunk_kb k-blocks (default 2 = 128 reduction elements = 8 accumulation steps: measured,
  // tests/manual/fp64_arbiter.py -- the truncation bias is then below the fp32 CPU oracle's own rounding noise)
  const int max_chunk = d->chunk_kb > 0 ? d->chunk_kb : 2;
  const int nchunks = (a.num_k_blocks + max_chunk - 1) / max_chunk;
  a.chunk_kb = (a.num_k_blocks + nchunks - 1) / nchunks;

  int BN = d->block_n > 0 ? d->block_n : pick_bn(d->Cout, m_tiles, a.num_k_blocks);
  if (BN != 32 && BN != 64 && BN != 128) return fail(XDET_EINVAL, "conv2d_f16x2: block_n must be 32, 64 or 128");
  const size_t tail = 4 * kMaxBN * sizeof(float) + (2 * kMaxStages + 4) * sizeof(uint64_t) + 64;
  const size_t cap = d->max_ctas > 0 ? kSmemCapShared : kSmemCapAlone;
  const size_t stage_bytes = 2 * (size_t)kBM * kBK * 2 + 2 * ((((size_t)BN * kBK * 2) + 1023) & ~(size_t)1023);
  int stages = (int)((cap - 1024 - tail) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail(XDET_EINVAL, "conv2d_f16x2: tile does not fit shared memory");
  // clusters of two CTAs sharing the B tile: on request (2), never (1); auto (0) = off unless the tuner picks it
  const bool cluster = d->cluster == 2 && m_tiles >= 2;
  a.BN = BN;
  a.stages = stages;
  a.tmem_cols = 4 * BN;  // two buffers x (acc0 + acc1): 128 / 256 / 512 columns
  a.n_tiles_n = (d->Cout + BN - 1) / BN;
  const long long total = m_tiles * a.n_tiles_n;
  if (total >= (1ll << 31)) return fail(XDET_EINVAL, "conv2d_f16x2: too many tiles");
  a.total_tiles = (int)total;
  a.scale = d->scale;
  a.bias = d->bias;
  a.relu = d->relu;
  a.residual = d->residual;
  a.out = d->out;
  a.out_sn = out_sn;
  a.out_sy = out_sy;
  a.out_sx = out_sx;
  a.out_sc = d->out_sc;
  a.out_pair = reinterpret_cast<__half*>(d->out_pair);
  a.pair_plane = d->pair_plane;
  a.pair_cs = d->pair_cs;
  a.out2 = d->out2;
  a.scale2 = d->scale2;
  a.bias2 = d->bias2;
  a.out2_pair = reinterpret_cast<__half*>(d->out2_pair);
  auto aligned = [&](int bytes) {
    const uintptr_t m = (uintptr_t)bytes - 1;
    const long long e = bytes / 4;
    return d->out_sc == 1 && out_sx % e == 0 && out_sy % e == 0 && out_sn % e == 0 &&
           (reinterpret_cast<uintptr_t>(d->out) & m) == 0 && (reinterpret_cast<uintptr_t>(d->out2) & m) == 0 &&
           (reinterpret_cast<uintptr_t>(d->residual) & m) == 0;
  };
  a.vec_ok = aligned(32) ? 2 : (aligned(16) ? 1 : 0);
  a.pair_vec32 = (d->pair_cs % 16 == 0 && d->pair_plane % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out_pair) & 31) == 0 &&
                  (reinterpret_cast<uintptr_t>(d->out2_pair) & 31) == 0) ? 1 : 0;

  CUtensorMap map_a[2], map_b[2];
  for (int pl = 0; pl < 2; ++pl) {
    const __half* base = reinterpret_cast<const __half*>(d_in_pair) + (size_t)pl * d->in_plane;
    if (fold) {
      const cuuint64_t dims[4] = {(cuuint64_t)kBK, (cuuint64_t)Wout, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)sw * d->in_cs * 2, (cuuint64_t)d->in_wp * d->in_cs * 2,
                                     (cuuint64_t)d->in_wp * d->in_cs * 2 * H};
      const cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)a.BW, (cuuint32_t)(a.BH * sh), 1};
      const cuuint32_t estr[4] = {1, 1, (cuuint32_t)sh, 1};
      XDET_TRY(encode_f16(&map_a[pl], base, 4, dims, strides, box, estr));
    } else {
      if (a.BW * sw > 256 || a.BH * sh > 256) return fail(XDET_EINVAL, "conv2d_f16x2: strided tile exceeds the TMA box limit");
      const cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)d->in_cs * 2, (cuuint64_t)d->in_cs * 2 * W,
                                     (cuuint64_t)d->in_cs * 2 * W * H};
      const cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)(a.BW * sw), (cuuint32_t)(a.BH * sh), 1};
      const cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
      XDET_TRY(encode_f16(&map_a[pl], base, 4, dims, strides, box, estr));
    }
    const __half* wbase = reinterpret_cast<const __half*>(d->weights) + (size_t)pl * d->w_plane;
    const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)d->Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)(cluster ? BN / 2 : BN)};  // cluster: each CTA loads half
    XDET_TRY(encode_f16(&map_b[pl], wbase, 2, dims, strides, box));
  }
  const size_t smem = (size_t)stages * stage_bytes + tail + 1024;
  int sms = kNumSMs;
  if (d->max_ctas > 0 && d->max_ctas < sms) sms = d->max_ctas;
  sms &= ~1;
  a.m_tiles = (int)m_tiles;
  a.total_pairs = (int)((m_tiles + 1) / 2) * a.n_tiles_n;
  const int grid = cluster ? 2 * (a.total_pairs < sms / 2 ? a.total_pairs : sms / 2)
                           : (a.total_tiles < sms ? a.total_tiles : sms);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (d->max_ctas <= 0) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le;
  if (cluster) {
    if (BN == 128) le = launch<64, true>(cfg, map_a[0], map_a[1], map_b[0], map_b[1], a);
    else if (BN == 64) le = launch<32, true>(cfg, map_a[0], map_a[1], map_b[0], map_b[1], a);
    else le = launch<16, true>(cfg, map_a[0], map_a[1], map_b[0], map_b[1], a);
  } else {
    if (BN == 128) le = launch<64, false>(cfg, map_a[0], map_a[1], map_b[0], map_b[1], a);
    else if (BN == 64) le = launch<32, false>(cfg, map_a[0], map_a[1], map_b[0], map_b[1], a);
    else le = launch<16, false>(cfg, map_a[0], map_a[1], map_b[0], map_b[1], a);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(le != cudaSuccess ? le : cudaGetLastError(), "conv_gemm_f16x2_kernel");
}

extern "C" int xdet_split2_f16(const float* d_src, long long sn, long long sy, long long sx, long long sc, int N, int H,
                               int W, int C, void* d_dst, int cs, int Wp, int x_off, long long plane, int relu,
                               void* stream) {
  if (N < 0 || H < 0 || W < 0 || C <= 0) return fail(XDET_EINVAL, "split2: bad shape");
  if (cs < C || cs % 8 != 0 || plane % 8 != 0 || Wp < W + x_off || x_off < 0)
    return fail(XDET_EINVAL, "split2: cs (%d) must be >= C and %% 8 == 0, Wp >= W + x_off", cs);
  if (reinterpret_cast<uintptr_t>(d_dst) & 15) return fail(XDET_EINVAL, "split2: destination must be 16-byte aligned");
  const long long total = (long long)N * H * W * (cs / 8);
  if (total == 0) return XDET_OK;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  split2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_src, sn, sy, sx, sc, H, W, C,
                                                                   reinterpret_cast<__half*>(d_dst), cs, Wp, x_off, plane,
                                                                   relu, total);
  return after_launch("split2_kernel");
}

extern "C" int xdet_maxpool3x3s2_f32x(const float* d_src, float* d_dst, void* d_dst_pair, float* d_dst2, void* d_dst2_pair,
                                      const float* d_scale2, const float* d_bias2, const float* d_residual,
                                      long long pair_plane, int N, int H, int W, int C, int Ho, int Wo, int pad_top,
                                      int pad_left, void* stream) {
  if (N < 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return fail(XDET_EINVAL, "maxpool_f32x: C must be a positive multiple of 8");
  if ((d_dst2 || d_dst2_pair) && (!d_scale2 || !d_bias2)) return fail(XDET_EINVAL, "maxpool_f32x: second output needs scale2 and bias2");
  if (!d_dst && !d_dst_pair && !d_dst2 && !d_dst2_pair) return fail(XDET_EINVAL, "maxpool_f32x: no output");
  if ((d_dst_pair || d_dst2_pair) && pair_plane % 8 != 0) return fail(XDET_EINVAL, "maxpool_f32x: pair planes must be 16-byte aligned");
  const long long total = (long long)N * Ho * Wo * (C / 8);
  if (total == 0) return XDET_OK;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  maxpool3x3s2_f32x_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      d_src, d_dst, reinterpret_cast<__half*>(d_dst_pair), d_dst2, reinterpret_cast<__half*>(d_dst2_pair), d_scale2,
      d_bias2, d_residual, pair_plane, H, W, C, Ho, Wo, pad_top, pad_left, total);
  return after_launch("maxpool3x3s2_f32x_kernel");
}

static int g_dw_tma = 1;
// 0: the register-window kernel (kept as the second implementation the tests compare against), 1: the TMA-staged one
extern "C" void xdet_set_depthwise_f32_tma(int enabled) { g_dw_tma = enabled ? 1 : 0; }

extern "C" int xdet_depthwise3x3_f32x(const float* d_src, const float* d_weights, float* d_dst, void* d_dst_pair,
                                      long long pair_plane, int N, int H, int W, int C, int dilation, int relu_in,
                                      void* stream) {
  if (N < 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return fail(XDET_EINVAL, "depthwise_f32x: C must be a positive multiple of 8");
  if (dilation != 1 && dilation != 2) return fail(XDET_EINVAL, "depthwise_f32x: dilation must be 1 or 2");
  if (!d_dst && !d_dst_pair) return fail(XDET_EINVAL, "depthwise_f32x: no output");
  if (d_dst_pair && pair_plane % 8 != 0) return fail(XDET_EINVAL, "depthwise_f32x: pair planes must be 16-byte aligned");
  if (N == 0) return XDET_OK;
  cudaStream_t st = (cudaStream_t)stream;
  __half* pp = reinterpret_cast<__half*>(d_dst_pair);
  if ((reinterpret_cast<uintptr_t>(d_src) & 15) == 0 && C % 4 == 0 && W >= 4 && H >= 4 && g_dw_tma) {
    // TMA-staged kernel: fp32 tensor map {C, W, H, N}, box {32, TW+2D, TH+2D, 1}, no swizzle, zero fill outside
    auto fn = encode_fn();
    if (!fn) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)C * 4 * W, (cuuint64_t)C * 4 * W * H};
    const cuuint32_t box[4] = {(cuuint32_t)kDwCB, (cuuint32_t)(kDwTW + 2 * dilation), (cuuint32_t)(kDwTH + 2 * dilation), 1};
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d_src), dims, strides, box, ones,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled(depthwise) failed with CUresult %d", (int)r);
    const int tiles_x = (W + kDwTW - 1) / kDwTW, tiles_y = (H + kDwTH - 1) / kDwTH, tiles_c = (C + kDwCB - 1) / kDwCB;
    const long long total_t = (long long)N * tiles_y * tiles_x * tiles_c;
    if (total_t < (1ll << 31)) {
      const size_t box_bytes = (size_t)(kDwTW + 2 * dilation) * (kDwTH + 2 * dilation) * kDwCB * 4;
      const size_t smem = kDwStages * box_bytes + kDwStages * sizeof(uint64_t) + 128;
      const int per_sm = smem <= 110 * 1024 ? 2 : 1;
      const int grid = (int)std::min<long long>(total_t, (long long)kNumSMs * per_sm);
      if (dilation == 1) {
        XDET_TRY(check_cuda(cudaFuncSetAttribute(depthwise3x3_f32_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem), "cudaFuncSetAttribute(depthwise tma)"));
        depthwise3x3_f32_tma_kernel<1><<<grid, kDwThreads, smem, st>>>(map, d_weights, d_dst, pp, pair_plane, N, H, W, C,
                                                                      relu_in, tiles_x, tiles_y, tiles_c, (int)total_t);
      } else {
        XDET_TRY(check_cuda(cudaFuncSetAttribute(depthwise3x3_f32_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem), "cudaFuncSetAttribute(depthwise tma)"));
        depthwise3x3_f32_tma_kernel<2><<<grid, kDwThreads, smem, st>>>(map, d_weights, d_dst, pp, pair_plane, N, H, W, C,
                                                                      relu_in, tiles_x, tiles_y, tiles_c, (int)total_t);
      }
      return after_launch("depthwise3x3_f32_tma_kernel");
    }
  }
  // strip height: tall strips amortise the window fill, but keep >= ~8 CTAs per SM in flight
  int YS = 16;
  while (YS > 2 && (long long)N * ((H + YS - 1) / YS) * W * (C / 4) < 8ll * kNumSMs * 256) YS /= 2;
  const long long total = (long long)N * ((H + YS - 1) / YS) * W * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > (1ll << 30)) blocks = 1ll << 30;
  if (dilation == 1)
    depthwise3x3_f32x_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(d_src, d_weights, d_dst, pp, pair_plane, N, H, W, C, relu_in, YS, total);
  else
    depthwise3x3_f32x_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(d_src, d_weights, d_dst, pp, pair_plane, N, H, W, C, relu_in, YS, total);
  return after_launch("depthwise3x3_f32x_kernel");
}
