// Implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (sm_100a).
//
//   D[m, n] = sum_{tap, c} A[m shifted by tap, c] * Wt[n, tap, c]      (bf16 x bf16 -> fp32 in TMEM)
//
// Replaces the cuDNN / cuBLAS calls TensorFlow makes for every tf.layers.conv2d / dense on the
// Light-Head R-CNN path (net/resnet_v2.py:89-100,320-325, net/xception_body.py:243-376,381-400,450-475,
// 540-558).  One PERSISTENT kernel covers 1x1 / 3x3 / 7x7 / 15x1 / 1x15 convolutions (dilated, stride 1
// or 2) and plain GEMMs (dense layers; 1x1 convs are flattened to [N*H*W, C]):
//
//   * activations are NHWC bf16; a tile of 128 output pixels is a BH x BW patch of one image.  For
//     each filter tap the A operand is ONE TMA 4-D box load {64 ch, BW, BH, 1} at the tap's shifted
//     coordinates -- out-of-bounds rows/columns/channels are zero-filled by TMA, which implements
//     SAME padding and channel tails without any im2col buffer; stride-2 convolutions use the tensor
//     map's element strides (every 2nd pixel of a 2*BW x 2*BH box);
//   * "fold_w" mode (the 3-channel 7x7/s2 stem): the KW taps of one filter row and the <= 8 padded input
//     channels are one contiguous run of KW*in_cs <= 64 elements, so a K chunk is a whole filter row and
//     the tensor map's pixel stride is stride_w*in_cs elements (overlapping windows);
//   * weights are [Cout][tap][Cin padded to 64] bf16 (K-major), loaded as 2-D boxes {64, BN};
//   * both operands land in shared memory in the 128-byte-swizzled K-major layout tcgen05 consumes
//     directly through shared-memory matrix descriptors;
//   * each CTA (one per SM) walks tiles blockIdx.x, +gridDim.x, ...: warp 0 = TMA producer, warp 1 = MMA
//     issuer (one elected thread, tcgen05.mma cta_group::1, M=128, N=BN<=256, K=16) and TMEM owner,
//     warps 2-5 = epilogue (tcgen05.ld 32x32b, each warp its own 32-lane TMEM quadrant).  Three
//     pipelines: a full/empty mbarrier ring of shared-memory stages (TMA <-> MMA), TWO accumulator
//     buffers in TMEM with full/empty barriers (MMA <-> epilogue: the epilogue of tile i overlaps the
//     main loop of tile i+1), and the epilogue's own double-buffered staging tiles;
//   * epilogue: y = acc*scale[c] + bias[c] (+ residual) (ReLU), and optionally a second output
//     relu(y*scale2[c]+bias2[c]) (the next pre-activation BN+ReLU of a ResNet-v2 block,
//     net/resnet_v2.py:163-164) so that no separate normalisation pass exists.  bf16 NHWC outputs go
//     registers -> swizzled shared memory -> TMA tile store (coalesced, edges clipped by TMA) and the
//     residual tile arrives by TMA load, prefetched two 64-channel chunks ahead; fp32 / strided outputs
//     (NCHW fp32 for PsRoIAlign, fp32 NHWC for the RPN decode) are written directly from registers.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace xdet {
namespace {

constexpr int kBM = 128;          // rows (output pixels) per tile
constexpr int kBK = 64;           // K per stage: 64 bf16 = one 128-byte swizzle atom
constexpr int kUmmaK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int kGemmThreads = 320; // 10 warps: TMA, MMA, 2 epilogue groups x 4
constexpr int kEpiThreads = 128;
constexpr int kMaxStages = 8;
constexpr int kMaxBN = 256;
constexpr int kMaxSlots = 3;
constexpr uint32_t kChunkBytes = kBM * 64 * 2;  // one 128 x 64-channel bf16 staging tile
// dynamic shared memory per CTA: leaves ~16 KB of the SM for small CTAs of concurrent kernels (the NMS
// bit-matrix kernel of the proposal stream runs beside the backbone convolutions)
constexpr size_t kSmemCapShared = 211 * 1024;  // while another stream runs beside the convolutions (max_ctas set)
constexpr size_t kSmemCapAlone = 227 * 1024;

struct ConvGemmArgs {
  int tiles_x, tiles_y, n_tiles_n, total_tiles;
  int BW, BH;
  int Hout, Wout, Cout;
  int taps_w, dil_h, dil_w, pad_top, pad_left, mul_x, mul_y;
  int k_chunks_per_tap, num_k_blocks;
  int BN, bn_pad, stages, tmem_cols;
  const float* scale;
  const float* bias;
  int relu;
  int tma_epilogue, has_res, has_out2, skip_out;
  int epi_groups, n_slots, n_out2;  // epilogue warpgroups; staging slots / second-output buffers per group
  void* out;
  int out_fp32;
  long long out_sn, out_sy, out_sx, out_sc;
  const float* scale2;
  const float* bias2;
  float* stats;    // [2][Cout] sums / sums of squares of the stored bf16 outputs (NULL: off)
  int stats_cols;  // Cout rounded up to 64: length of one of the CTA's two shared-memory accumulators
};

struct TileCoord {
  int x0, y0, img, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvGemmArgs& p, int t) {
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

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                 const __grid_constant__ CUtensorMap map_out2, const ConvGemmArgs p) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte-swizzled tiles need 1024-byte aligned bases: align manually (the launch adds slack)
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  // carve: [stages][A 16 KB][B BN*128 B] | epilogue staging (out x2, out2 x2, residual x2) | scale/bias | barriers
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = (uint32_t)p.BN * kBK * 2;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  unsigned char* tiles = smem;
  unsigned char* ebuf = smem + (size_t)p.stages * stage_bytes;  // [groups][n_slots + n_out2][16 KB]
  float* sbuf = reinterpret_cast<float*>(ebuf + (size_t)p.epi_groups * (p.n_slots + p.n_out2) * kChunkBytes);  // [2][4][kMaxBN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbuf + 2 * 4 * kMaxBN);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* res_full_bar = tmem_empty_bar + 2;       // [2 groups][kMaxSlots]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full_bar + 2 * kMaxSlots);
  // statistics (p.stats): [2 groups][4 warps][8 units][16] partials of the chunk in flight, then the groups' accumulators
  // [2 groups][2][stats_cols] (16-byte aligned: see `tail`)
  float* s_part = reinterpret_cast<float*>(tmem_ptr + 16);
  float* s_stats = s_part + 2 * 4 * 8 * 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.stats)
    for (int i = threadIdx.x; i < 4 * p.stats_cols; i += kGemmThreads) s_stats[i] = 0.f;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_a);
    ptx::prefetch_tmap(&map_b);
    if (p.tma_epilogue) ptx::prefetch_tmap(&map_out);
    if (p.has_res) ptx::prefetch_tmap(&map_res);
    if (p.has_out2) ptx::prefetch_tmap(&map_out2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&tmem_full_bar[s], 1);
        ptx::mbar_init(&tmem_empty_bar[s], 4 * p.epi_groups);  // one arrival per epilogue warp
      }
      for (int s = 0; s < 2 * kMaxSlots; ++s) ptx::mbar_init(&res_full_bar[s], 1);
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
  // Programmatic dependent launch: this kernel may have started while its predecessor was still draining (its
  // prologue above touches no tensor data).  Let OUR successor start its prologue as SMs free up, and make every
  // thread that reads upstream tensors (the TMA producer, the residual loaders, the epilogue's scale/bias reads of
  // batch-norm statistics) wait for the predecessor's memory first.
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const int ax = tc.x0 * p.mul_x - p.pad_left, ay = tc.y0 * p.mul_y - p.pad_top;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          const int tap = kb / p.k_chunks_per_tap, cc = kb - tap * p.k_chunks_per_tap;
          const int kh = tap / p.taps_w, kw = tap - kh * p.taps_w;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = tiles + (size_t)stage * stage_bytes;
          unsigned char* sb = sa + a_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
          ptx::tma_load_4d(sa, &map_a, &full_bar[stage], cc * kBK, ax + kw * p.dil_w, ay + kh * p.dil_h, tc.img);
          ptx::tma_load_2d(sb, &map_b, &full_bar[stage], kb * kBK, tc.n0);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = ptx::make_idesc_bf16(kBM, p.BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator buffer
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.bn_pad);
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(tiles + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(sa + k * kUmmaK * 2);
            const uint64_t db = ptx::make_smem_desc_sw128(sb + k * kUmmaK * 2);
            ptx::mma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&empty_bar[stage]);                                 // slot reusable once these MMAs retire
          if (kb == p.num_k_blocks - 1) ptx::mma_commit(&tmem_full_bar[as]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else if (((warp - 2) >> 2) < p.epi_groups) {
    // ===== epilogue: warps 2..5 (group 0) and 6..9 (group 1), TMEM quadrant = warp % 4 =====
    // The 64-channel chunks of the tile sequence are dealt round-robin to the groups; each group owns its
    // staging slots, scale/bias vectors, named barrier and TMA store/load stream.
    const int grp = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int m = quad * 32 + lane;                        // row of the tile == TMEM lane
    const int e = (warp - 2 - grp * 4) * 32 + lane;        // 0..127 inside the group
    const bool leader = (e == 0);                          // issues the group's TMA loads / stores
    const uint32_t bar_id = 1u + (uint32_t)grp;
    const int nchunks = (p.BN + 63) / 64;
    const int G = p.epi_groups;
    const uint32_t swz = (uint32_t)(m & 7);
    unsigned char* gbase = ebuf + (size_t)grp * (size_t)(p.n_slots + p.n_out2) * kChunkBytes;
    float* sb = sbuf + grp * 4 * kMaxBN;
    uint64_t* res_bar = res_full_bar + grp * kMaxSlots;
    int as = 0;
    uint32_t aphase = 0;
    int q = 0;               // chunks this group has processed (slot = q % n_slots)
    uint32_t res_phase = 0;  // bit s = parity of the next fill of residual slot s

    // residual tile of this group's qq-th chunk (running chunk index j = qq*G + grp over the CTA's tile sequence)
    auto issue_residual = [&](int qq) {
      const int j = qq * G + grp;
      const int ti = j / nchunks, c = j - ti * nchunks;
      const long long t = (long long)blockIdx.x + (long long)ti * gridDim.x;
      if (t >= p.total_tiles) return;
      const TileCoord tc = decode_tile(p, (int)t);
      if (tc.n0 + c * 64 >= p.Cout) return;  // chunk fully outside: the consumer skips it as well
      const int sl = qq % p.n_slots;
      ptx::mbar_arrive_expect_tx(&res_bar[sl], kChunkBytes);
      ptx::tma_load_4d(gbase + (size_t)sl * kChunkBytes, &map_res, &res_bar[sl], tc.n0 + c * 64, tc.x0, tc.y0, tc.img);
    };
    if (p.has_res && leader) {
      issue_residual(0);
      issue_residual(1);
    }

    int ti = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
      const TileCoord tc = decode_tile(p, t);
      const int ncols = min(p.BN, p.Cout - tc.n0);
      // this group's chunks of the tile: c = c_first, c_first + G, ...
      int c_first = (grp - (ti * nchunks) % G + G) % G;
      const bool mine = c_first < nchunks;
      if (mine) {
        // per-tile scale / bias vectors -> shared memory (this group's previous readers passed their last barrier)
        for (int i = e; i < p.BN; i += kEpiThreads) {
          const int c = tc.n0 + i;
          const bool ok = c < p.Cout;
          sb[i] = (ok && p.scale) ? __ldg(p.scale + c) : 1.f;
          sb[kMaxBN + i] = (ok && p.bias) ? __ldg(p.bias + c) : 0.f;
          if (p.has_out2) {
            sb[2 * kMaxBN + i] = ok ? __ldg(p.scale2 + c) : 1.f;
            sb[3 * kMaxBN + i] = ok ? __ldg(p.bias2 + c) : 0.f;
          }
        }
        ptx::named_bar_sync(bar_id, kEpiThreads);
      }
      ptx::mbar_wait(&tmem_full_bar[as], aphase);
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * p.bn_pad);
      bool released = false;
      auto release_tmem = [&]() {  // accumulator fully read by this warp: hand the buffer back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
        released = true;
      };

      if (p.tma_epilogue) {
        for (int c = c_first; c < nchunks; c += G, ++q) {
          const int col0 = c * 64;
          const bool last = c + G >= nchunks;
          if (col0 >= ncols) {  // whole chunk beyond Cout (only when BN > remaining channels)
            if (p.has_res && leader) {
              ptx::bulk_wait_read<0>();  // no store of its own follows: every earlier one must be done
              issue_residual(q + 2);     // keep the prefetch chain going
            }
            continue;
          }
          const int sl = q % p.n_slots;
          unsigned char* obuf = gbase + (size_t)sl * kChunkBytes;
          unsigned char* o2buf = gbase + (size_t)(p.n_slots + (q % max(p.n_out2, 1))) * kChunkBytes;
          if (leader) {  // the stores that last read the buffers written below have finished reading them
            if (p.n_out2 == 1) ptx::bulk_wait_read<0>(); else ptx::bulk_wait_read<1>();
          }
          ptx::named_bar_sync(bar_id, kEpiThreads);
          uint32_t r0[32], r1[32];
          ptx::tmem_ld_32x32(t_acc + (uint32_t)col0, r0);
          ptx::tmem_ld_32x32(t_acc + (uint32_t)col0 + 32u, r1);
          ptx::tmem_ld_wait();
          if (last) release_tmem();
          if (p.has_res) {
            ptx::mbar_wait(&res_bar[sl], (res_phase >> sl) & 1u);
            res_phase ^= (1u << sl);
          }
          unsigned char* crow = obuf + (size_t)m * 128;   // residual in, result out: same 16-byte units, same thread
          unsigned char* c2row = o2buf + (size_t)m * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u) {  // 8 channels = one 16-byte unit of the swizzled row
            float v[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int cj = u * 8 + jj;
              const float acc = __uint_as_float(cj < 32 ? r0[cj & 31] : r1[cj & 31]);
              v[jj] = fmaf(acc, sb[col0 + cj], sb[kMaxBN + col0 + cj]);
            }
            const uint32_t off = ((uint32_t)u ^ swz) << 4;
            if (p.has_res) {
              const uint4 rr = *reinterpret_cast<const uint4*>(crow + off);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const float2 f = __bfloat1622float2(h[jj]);
                v[2 * jj] += f.x;
                v[2 * jj + 1] += f.y;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) v[jj] = fmaxf(v[jj], 0.f);
            }
            if (!p.skip_out)
              *reinterpret_cast<uint4*>(crow + off) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            if (p.has_out2) {
              float w[8];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const int cj = u * 8 + jj;
                w[jj] = fmaxf(fmaf(v[jj], sb[2 * kMaxBN + col0 + cj], sb[3 * kMaxBN + col0 + cj]), 0.f);
              }
              *reinterpret_cast<uint4*>(c2row + off) = make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]),
                                                                  pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
            }
          }
          ptx::fence_proxy_async();  // generic-proxy writes -> visible to the TMA (async proxy)
          ptx::named_bar_sync(bar_id, kEpiThreads);
          if (leader) {
            if (!p.skip_out) ptx::tma_store_4d(&map_out, obuf, tc.n0 + col0, tc.x0, tc.y0, tc.img);
            if (p.has_out2) ptx::tma_store_4d(&map_out2, o2buf, tc.n0 + col0, tc.x0, tc.y0, tc.img);
            ptx::bulk_commit();
            if (p.has_res) {
              // slot (q+2) % 3 was last read by the store of chunk q-1: all but the newest store must be done
              ptx::bulk_wait_read<1>();
              issue_residual(q + 2);
            }
          }
          if (p.stats) {
            // Batch-norm statistics of what was just stored (the bf16 values, as the next layer reads them): the staged
            // 128 x 64 tile is re-read column-wise -- thread = (16-byte unit u of 8 channels, row group rg), rows
            // rg, rg+16, ... -- reduced over the warp's four row groups by shuffles and added to the CTA's accumulators.
            // (The slot is not rewritten before every thread of the group has passed the next chunk's barrier.)
            const int u = e & 7, rg = e >> 3;
            const int bw_shift = __ffs(p.BW) - 1;
            float sa[8], sq[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) sa[jj] = sq[jj] = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int r = rg + 16 * k;
              const bool ok = (tc.y0 + (r >> bw_shift) < p.Hout) && (tc.x0 + (r & (p.BW - 1)) < p.Wout);
              if (ok) {
                const uint4 w4 = *reinterpret_cast<const uint4*>(obuf + (size_t)r * 128 + ((uint32_t)(u ^ (r & 7)) << 4));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w4);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  const float2 f = __bfloat1622float2(h[jj]);
                  sa[2 * jj] += f.x;
                  sa[2 * jj + 1] += f.y;
                  sq[2 * jj] = fmaf(f.x, f.x, sq[2 * jj]);
                  sq[2 * jj + 1] = fmaf(f.y, f.y, sq[2 * jj + 1]);
                }
              }
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              sa[jj] += __shfl_xor_sync(0xffffffffu, sa[jj], 8);
              sq[jj] += __shfl_xor_sync(0xffffffffu, sq[jj], 8);
              sa[jj] += __shfl_xor_sync(0xffffffffu, sa[jj], 16);
              sq[jj] += __shfl_xor_sync(0xffffffffu, sq[jj], 16);
            }
            // the four warps' partials meet in shared memory (plain stores: shared-memory float atomics are CAS loops);
            // after the barrier thread e owns value e & 15 (8 sums, 8 squares) of unit e >> 4 and adds it to the GROUP's
            // accumulator -- the only writer of that element, since a group works through its chunks one at a time
            float* part = s_part + grp * (4 * 8 * 16);
            if (lane < 8) {
              float4* dst = reinterpret_cast<float4*>(part + ((warp & 3) * 8 + u) * 16);
              dst[0] = make_float4(sa[0], sa[1], sa[2], sa[3]);
              dst[1] = make_float4(sa[4], sa[5], sa[6], sa[7]);
              dst[2] = make_float4(sq[0], sq[1], sq[2], sq[3]);
              dst[3] = make_float4(sq[4], sq[5], sq[6], sq[7]);
            }
            ptx::named_bar_sync(bar_id, kEpiThreads);
            {
              const int uu = e >> 4, kk = e & 15;
              const int c = tc.n0 + col0 + uu * 8 + (kk & 7);
              if (c < p.Cout) {
                const float v = part[(0 * 8 + uu) * 16 + kk] + part[(1 * 8 + uu) * 16 + kk] +
                                part[(2 * 8 + uu) * 16 + kk] + part[(3 * 8 + uu) * 16 + kk];
                s_stats[(grp * 2 + (kk >> 3)) * p.stats_cols + c] += v;
              }
            }
          }
        }
        if (!released) release_tmem();  // no chunk of this tile (or only skipped ones) was this group's
      } else {
        // direct register -> global path (fp32 and/or strided outputs); single group
        const int py = tc.y0 + m / p.BW, px = tc.x0 + m % p.BW;
        const bool row_ok = (py < p.Hout) && (px < p.Wout);
        const long long pix_off =
            (long long)tc.img * p.out_sn + (long long)py * p.out_sy + (long long)px * p.out_sx;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_acc + (uint32_t)c0, r);
          ptx::tmem_ld_wait();
          if (!row_ok) continue;
          const int nv = min(32, ncols - c0);
          const long long cbase = tc.n0 + c0;
          if (p.out_fp32) {
            float* op = reinterpret_cast<float*>(p.out) + pix_off + cbase * p.out_sc;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nv) {
                float a = fmaf(__uint_as_float(r[j]), sb[c0 + j], sb[kMaxBN + c0 + j]);
                if (p.relu) a = fmaxf(a, 0.f);
                op[(long long)j * p.out_sc] = a;
              }
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix_off + cbase * p.out_sc;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nv) {
                float a = fmaf(__uint_as_float(r[j]), sb[c0 + j], sb[kMaxBN + c0 + j]);
                if (p.relu) a = fmaxf(a, 0.f);
                op[(long long)j * p.out_sc] = __float2bfloat16_rn(a);
              }
            }
          }
        }
        release_tmem();
        ptx::named_bar_sync(bar_id, kEpiThreads);  // sb is rewritten at the top of the next tile
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (p.tma_epilogue && leader) ptx::bulk_wait_all();  // staging tiles must outlive their stores
    if (p.stats) {
      // every epilogue warp of the CTA has added its last chunk: the groups' accumulators are summed and leave as one
      // reduction per element (a 148th of what per-tile atomics would send to the same few cache lines)
      ptx::named_bar_sync(5u, (uint32_t)(kEpiThreads * p.epi_groups));
      const int ei = grp * kEpiThreads + e, nthr = kEpiThreads * p.epi_groups;
      for (int i = ei * 4; i < 2 * p.Cout; i += nthr * 4) {  // (Cout % 8 == 0: a group of four never straddles the halves)
        const int h = i >= p.Cout ? 1 : 0, c = i - h * p.Cout;
        float4 v = *reinterpret_cast<const float4*>(s_stats + h * p.stats_cols + c);
        if (p.epi_groups == 2) {
          const float4 w = *reinterpret_cast<const float4*>(s_stats + (2 + h) * p.stats_cols + c);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
          asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.stats + i), "f"(v.x), "f"(v.y),
                       "f"(v.z), "f"(v.w)
                       : "memory");
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- host side -------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* estr = nullptr) {
  auto fn = get_encode_fn();
  if (!fn) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                  strides_bytes, box, estr ? estr : ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return XDET_OK;
}

int next_pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// N-tile heuristic: estimated clocks of the busiest SM = waves * (k-blocks * per-block time + epilogue).
// Per k-block the tile is bound by the MMA (2*BN clocks for 128 x BN x 64 at 8192 MAC-flops/clk/SM) or by
// the operand traffic from L2 ((16 KB + BN*128 B) at ~kL2BytesPerClk per SM with every SM pulling).
constexpr double kL2BytesPerClk = 56.0;
int pick_block_n(int cout, long long m_tiles, int num_k_blocks, bool multiples_of_64) {
  int cands[4] = {256, 128, 64, 0};
  if (!multiples_of_64 && cout < 256) cands[3] = ((cout + 15) / 16) * 16;  // exact fit
  int best = 0;
  double best_cost = 0;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (bn == 0) continue;
    const long long tiles = m_tiles * ((cout + bn - 1) / bn);
    const long long waves = (tiles + kNumSMs - 1) / kNumSMs;
    const double per_kb = std::max(2.0 * bn, (16384.0 + bn * 128.0) / kL2BytesPerClk);
    const double cost = (double)waves * (num_k_blocks * per_kb + 6.0 * bn + 600.0);
    if (best == 0 || cost < best_cost * 0.98) {
      best = bn;
      best_cost = cost;
    }
  }
  return best;
}

}  // namespace
}  // namespace xdet

using namespace xdet;

static int g_conv_pdl = 1;
extern "C" void xdet_set_conv_pdl(int enabled) { g_conv_pdl = enabled ? 1 : 0; }

extern "C" int xdet_conv2d_bf16(const void* d_in, const xdet_conv_desc* d, void* stream) {
  if (!d || !d_in) return fail(XDET_EINVAL, "null argument");
  if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->KH <= 0 || d->KW <= 0)
    return fail(XDET_EINVAL, "conv2d: non-positive dimension");
  if (d->Hout <= 0 || d->Wout <= 0) return fail(XDET_EINVAL, "conv2d: non-positive output size");
  if (d->in_cs < d->Cin || (d->in_cs % 8) != 0)
    return fail(XDET_EINVAL, "conv2d: input channel stride (%d) must be >= Cin and a multiple of 8 (TMA 16-byte strides)",
                d->in_cs);
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d->weights) & 15))
    return fail(XDET_EINVAL, "conv2d: input and weights must be 16-byte aligned");
  const int sh = d->stride_h <= 0 ? 1 : d->stride_h, sw = d->stride_w <= 0 ? 1 : d->stride_w;
  if (sh > 2 || sw > 2) return fail(XDET_EINVAL, "conv2d: strides 1 and 2 are supported");
  const int dil_h = d->dil_h <= 0 ? 1 : d->dil_h, dil_w = d->dil_w <= 0 ? 1 : d->dil_w;
  const bool fold = d->fold_w != 0;
  if (fold && (d->KW * d->in_cs > kBK || dil_w != 1 || d->in_wp < (d->Wout - 1) * sw + kBK / d->in_cs))
    return fail(XDET_EINVAL, "conv2d: fold_w needs KW*in_cs <= 64, dil_w == 1 and in_wp >= (Wout-1)*stride_w + 64/in_cs");

  // geometry seen by the kernel (1x1 stride-1 convolutions over dense tensors flatten to one long row)
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

  ConvGemmArgs a{};
  // tile shape: widest power-of-two row segment that covers the output row, up to 128 (<= 128/stride so that the
  // strided box stays within TMA's 256-element limit)
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

  // epilogue mode
  if (!d->out && !d->out2) return fail(XDET_EINVAL, "conv2d: no output");
  a.skip_out = d->out ? 0 : 1;  // only the second output is wanted (its producer's raw value has no other reader)
  const bool out_bf16_nhwc = !d->out_fp32 && d->out_sc == 1 && (out_sx % 8) == 0 && (out_sy % 8) == 0 &&
                             (out_sn % 8) == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 && d->Cout >= 8;
  a.tma_epilogue = out_bf16_nhwc ? 1 : 0;
  a.has_res = d->residual ? 1 : 0;
  a.has_out2 = d->out2 ? 1 : 0;
  if (a.has_out2 && (!d->scale2 || !d->bias2)) return fail(XDET_EINVAL, "conv2d: out2 needs scale2 and bias2");
  if ((a.has_res || a.has_out2) && !a.tma_epilogue)
    return fail(XDET_EINVAL, "conv2d: residual / second output need a bf16 NHWC `out` (unit channel stride, 16-byte aligned "
                             "pixel strides); they share its layout");
  if ((a.has_res && (reinterpret_cast<uintptr_t>(d->residual) & 15)) || (a.has_out2 && (reinterpret_cast<uintptr_t>(d->out2) & 15)))
    return fail(XDET_EINVAL, "conv2d: residual / out2 must be 16-byte aligned");

  // epilogue staging: memory-bound layers (few k-blocks per tile) get two epilogue warpgroups; a residual is
  // prefetched two chunks ahead into a third slot and overwritten in place by the result
  a.epi_groups = (a.tma_epilogue && a.num_k_blocks <= 40) ? 2 : 1;  // measured: tools/conv_floor.py
  if (d->epi_groups == 1 || d->epi_groups == 2) a.epi_groups = a.tma_epilogue ? d->epi_groups : 1;
  a.n_slots = a.tma_epilogue ? (a.has_res ? 3 : 2) : 0;
  a.n_out2 = a.has_out2 ? (a.has_res ? 1 : 2) : 0;
  const int stats_cols = d->stats ? (d->Cout + 63) / 64 * 64 : 0;
  if (d->stats) {
    if (!a.tma_epilogue || a.skip_out)
      return fail(XDET_EINVAL, "conv2d: statistics need a stored bf16 NHWC `out`");
    if (d->Cout % 8 || d->Cout > 2048 || (reinterpret_cast<uintptr_t>(d->stats) & 15))
      return fail(XDET_EINVAL, "conv2d: statistics need Cout %% 8 == 0, Cout <= 2048 and a 16-byte aligned buffer");
  }
  const size_t tail = 2 * 4 * kMaxBN * sizeof(float) + (2 * kMaxStages + 4 + 2 * kMaxSlots) * sizeof(uint64_t) + 64 +
                      (d->stats ? (2 * 4 * 8 * 16 + 4 * (size_t)stats_cols) * sizeof(float) : 0);
  int BN = d->block_n > 0 ? d->block_n : pick_block_n(d->Cout, m_tiles, a.num_k_blocks, a.tma_epilogue != 0);
  if (BN % 16 != 0 || BN < 16 || BN > kMaxBN) return fail(XDET_EINVAL, "conv2d: block_n must be a multiple of 16 in [16,256]");
  if (a.tma_epilogue && BN % 64 != 0) return fail(XDET_EINVAL, "conv2d: block_n must be a multiple of 64 for bf16 NHWC outputs");
  // shared memory: [stages][A + B] + epilogue staging must fit; want >= 3 stages (>= 2 at the very least).
  // Shrink in this order: N tile 256 -> 128 (auto only), then one epilogue group instead of two.
  const size_t kSmemCap = d->max_ctas > 0 ? kSmemCapShared : kSmemCapAlone;
  size_t epi_bytes = 0, stage_bytes = 0;
  int stages = 0;
  for (;;) {
    epi_bytes = (size_t)kChunkBytes * a.epi_groups * (a.n_slots + a.n_out2);
    stage_bytes = (size_t)kBM * kBK * 2 + (((size_t)BN * kBK * 2 + 1023) & ~(size_t)1023);
    const size_t fixed = 1024 + tail + epi_bytes;
    stages = fixed < kSmemCap ? (int)((kSmemCap - fixed) / stage_bytes) : 0;
    const int want = a.num_k_blocks >= 3 ? 3 : 2;
    if (stages >= want) break;
    if (d->block_n <= 0 && BN > 128) { BN = 128; continue; }
    if (a.epi_groups == 2) { a.epi_groups = 1; continue; }
    if (stages >= 2) break;
    if (d->block_n <= 0 && BN > 64) { BN = 64; continue; }
    return fail(XDET_EINVAL, "conv2d: tile does not fit shared memory");
  }
  if (stages > kMaxStages) stages = kMaxStages;
  a.BN = BN;
  a.bn_pad = ((BN + 31) / 32) * 32;
  a.tmem_cols = next_pow2_cols(2 * a.bn_pad);
  a.n_tiles_n = (d->Cout + BN - 1) / BN;
  const long long total = m_tiles * a.n_tiles_n;
  if (total >= (1ll << 31)) return fail(XDET_EINVAL, "conv2d: too many tiles");
  a.total_tiles = (int)total;
  a.stages = stages;
  a.scale = d->scale;
  a.bias = d->bias;
  a.relu = d->relu;
  a.out = d->out;
  a.out_fp32 = d->out_fp32;
  a.out_sn = out_sn;
  a.out_sy = out_sy;
  a.out_sx = out_sx;
  a.out_sc = d->out_sc;
  a.scale2 = d->scale2;
  a.bias2 = d->bias2;
  a.stats = d->stats;
  a.stats_cols = stats_cols;

  CUtensorMap map_a, map_b, map_out, map_res, map_out2;
  if (fold) {
    // dims {64-element window, Wout windows (stride_w pixels apart, overlapping), H rows, N images}
    const cuuint64_t dims[4] = {(cuuint64_t)kBK, (cuuint64_t)Wout, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)sw * d->in_cs * 2, (cuuint64_t)d->in_wp * d->in_cs * 2,
                                   (cuuint64_t)d->in_wp * d->in_cs * 2 * H};
    const cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)a.BW, (cuuint32_t)(a.BH * sh), 1};
    const cuuint32_t estr[4] = {1, 1, (cuuint32_t)sh, 1};
    XDET_TRY(encode_map(&map_a, d_in, 4, dims, strides, box, estr));
  } else {
    if (a.BW * sw > 256 || a.BH * sh > 256) return fail(XDET_EINVAL, "conv2d: strided tile exceeds the TMA box limit");
    const cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)d->in_cs * 2, (cuuint64_t)d->in_cs * 2 * W,
                                   (cuuint64_t)d->in_cs * 2 * W * H};
    const cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)(a.BW * sw), (cuuint32_t)(a.BH * sh), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
    XDET_TRY(encode_map(&map_a, d_in, 4, dims, strides, box, estr));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)d->Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)BN};
    XDET_TRY(encode_map(&map_b, d->weights, 2, dims, strides, box));
  }
  map_out = map_a;  // placeholders when unused (never dereferenced by the kernel)
  map_res = map_a;
  map_out2 = map_a;
  if (a.tma_epilogue) {
    const cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)Wout, (cuuint64_t)Hout, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)out_sx * 2, (cuuint64_t)out_sy * 2, (cuuint64_t)out_sn * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)a.BW, (cuuint32_t)a.BH, 1};
    XDET_TRY(encode_map(&map_out, d->out ? d->out : d->out2, 4, dims, strides, box));
    if (a.has_res) XDET_TRY(encode_map(&map_res, d->residual, 4, dims, strides, box));
    if (a.has_out2) XDET_TRY(encode_map(&map_out2, d->out2, 4, dims, strides, box));
  }
  const size_t smem = (size_t)stages * stage_bytes + epi_bytes + tail + 1024;  // +1024: manual alignment slack
  XDET_TRY(check_cuda(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                      "cudaFuncSetAttribute(conv_gemm)"));
  int sms = kNumSMs;
  if (d->max_ctas > 0 && d->max_ctas < sms) sms = d->max_ctas;  // leave SMs to a concurrent stream
  const int grid = a.total_tiles < sms ? a.total_tiles : sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // not while another stream shares the GPU (max_ctas set): early-launched CTAs would sit on the SMs left free for it
  cfg.numAttrs = (g_conv_pdl && d->max_ctas <= 0) ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, conv_gemm_kernel, map_a, map_b, map_out, map_res, map_out2, a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(le != cudaSuccess ? le : cudaGetLastError(), "conv_gemm_kernel");
}
