// Implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (sm_100a).
//
//   D[m, n] = sum_{tap, c} A[m shifted by tap, c] * Wt[n, tap, c]      (bf16 x bf16 -> fp32 in TMEM)
//
// Replaces the cuDNN / cuBLAS calls TensorFlow makes for every tf.layers.conv2d / dense on the
// Light-Head R-CNN path (net/resnet_v2.py:89-100, net/xception_body.py:243-376,381-400,450-475,
// 540-558).  One kernel covers 1x1 / 3x3 / 15x1 / 1x15 (dilated) stride-1 convolutions and plain
// GEMMs (dense layers, 1x1 convs flattened to [N*H*W, C]):
//
//   * activations are NHWC bf16; a tile of 128 output pixels is a BH x BW patch of one image.  For
//     each filter tap the A operand is ONE TMA 4-D box load {64 ch, BW, BH, 1} at the tap's shifted
//     coordinates -- out-of-bounds rows/columns/channels are zero-filled by TMA, which implements
//     SAME padding and channel tails without any im2col buffer;
//   * weights are [Cout][tap][Cin padded to 64] bf16 (K-major), loaded as 2-D boxes {64, BN};
//   * both operands land in shared memory in the 128-byte-swizzled K-major layout tcgen05 consumes
//     directly through shared-memory matrix descriptors;
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread, tcgen05.mma cta_group::1,
//     M=128, N=BN, K=16) and TMEM owner, warps 2-5 = epilogue (tcgen05.ld 32x32b, each warp its own
//     32-lane TMEM quadrant); full/empty mbarrier ring of `stages` smem slots, tcgen05.commit frees
//     a slot / publishes the accumulator;
//   * epilogue: y = acc*scale[c] + bias[c] (+ residual) (ReLU), written as bf16 or fp32 with arbitrary
//     element strides (NHWC bf16 for the next layer, NCHW fp32 for PsRoIAlign / the RPN decode), and
//     optionally a second output relu(y*scale2[c]+bias2[c]) (the next pre-activation BN+ReLU of a
//     ResNet-v2 block, net/resnet_v2.py:163-164) so that no separate normalisation pass exists.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace xdet {
namespace {

constexpr int kBM = 128;          // rows (output pixels) per tile
constexpr int kBK = 64;           // K per stage: 64 bf16 = one 128-byte swizzle atom
constexpr int kUmmaK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int kGemmThreads = 192; // 6 warps: TMA, MMA, 4x epilogue
constexpr int kMaxStages = 8;

struct ConvGemmArgs {
  int tiles_x, tiles_y, n_img;  // spatial tiling of the output (tiles_x * tiles_y * n_img M-tiles)
  int BW, BH;                   // tile shape, BW*BH == 128
  int Hout, Wout, Cout;
  int taps_w, dil_h, dil_w, pad_top, pad_left;
  int k_chunks_per_tap, num_k_blocks;
  int BN, stages, tmem_cols;
  const float* scale;
  const float* bias;
  int relu;
  const __nv_bfloat16* residual;
  void* out;
  int out_fp32;
  long long out_sn, out_sy, out_sx, out_sc;
  __nv_bfloat16* out2;
  const float* scale2;
  const float* bias2;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const ConvGemmArgs p) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte-swizzled operand tiles need 1024-byte aligned bases: align manually (the launch adds slack)
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  // carve: [stages][A 16 KB][B BN*128 B] | barriers | tmem ptr
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = (uint32_t)p.BN * kBK * 2;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  unsigned char* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int t = blockIdx.x;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int img = t / p.tiles_y;
  const int x0 = tx * p.BW, y0 = ty * p.BH;
  const int n0 = blockIdx.y * p.BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_a);
    ptx::prefetch_tmap(&map_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
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

  if (warp == 0) {
    // ===== TMA producer =====
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        const int tap = kb / p.k_chunks_per_tap, cc = kb - tap * p.k_chunks_per_tap;
        const int kh = tap / p.taps_w, kw = tap - kh * p.taps_w;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        unsigned char* sa = tiles + (size_t)stage * stage_bytes;
        unsigned char* sb = sa + a_bytes;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
        ptx::tma_load_4d(sa, &map_a, &full_bar[stage], cc * kBK, x0 + kw * p.dil_w - p.pad_left,
                         y0 + kh * p.dil_h - p.pad_top, img);
        ptx::tma_load_2d(sb, &map_b, &full_bar[stage], kb * kBK, n0);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = ptx::make_idesc_bf16(kBM, p.BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      ptx::mbar_wait(&full_bar[stage], phase);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t sa = ptx::smem_u32(tiles + (size_t)stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          const uint64_t da = ptx::make_smem_desc_sw128(sa + k * kUmmaK * 2);
          const uint64_t db = ptx::make_smem_desc_sw128(sb + k * kUmmaK * 2);
          ptx::mma_bf16_ss(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        ptx::mma_commit(&empty_bar[stage]);                           // slot reusable once these MMAs retire
        if (kb == p.num_k_blocks - 1) ptx::mma_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int m = quad * 32 + lane;  // row of the tile == TMEM lane
    const int py = y0 + m / p.BW, px = x0 + m % p.BW;
    const bool row_ok = (py < p.Hout) && (px < p.Wout);
    const long long pix_off = (long long)img * p.out_sn + (long long)py * p.out_sy + (long long)px * p.out_sx;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const int ncols = min(p.BN, p.Cout - n0);
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (!row_ok) continue;
      const int cbase = n0 + c0;
      const int nv = min(32, ncols - c0);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = cbase + j;
        float a = __uint_as_float(r[j]);
        if (j < nv) {
          const float sc = p.scale ? __ldg(p.scale + c) : 1.f;
          const float bi = p.bias ? __ldg(p.bias + c) : 0.f;
          a = fmaf(a, sc, bi);
        }
        v[j] = a;
      }
      const bool vec_ok = (p.out_sc == 1) && !p.out_fp32 && (nv == 32) && ((p.out_sx & 7) == 0) &&
                          ((p.out_sy & 7) == 0) && ((p.out_sn & 7) == 0) && ((cbase & 7) == 0);
      if (p.residual) {
        const __nv_bfloat16* rp = p.residual + pix_off + cbase;
        if (vec_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + q);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h[e]);
              v[q * 8 + e * 2] += f.x;
              v[q * 8 + e * 2 + 1] += f.y;
            }
          }
        } else {
          for (int j = 0; j < nv; ++j) v[j] += __bfloat162float(rp[(long long)j * p.out_sc]);
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.out_fp32) {
        float* op = reinterpret_cast<float*>(p.out) + pix_off + (long long)cbase * p.out_sc;
        for (int j = 0; j < nv; ++j) op[(long long)j * p.out_sc] = v[j];
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix_off + (long long)cbase * p.out_sc;
        if (vec_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[q * 8 + e * 2], v[q * 8 + e * 2 + 1]);
            reinterpret_cast<uint4*>(op)[q] = u;
          }
        } else {
          for (int j = 0; j < nv; ++j) op[(long long)j * p.out_sc] = __float2bfloat16_rn(v[j]);
        }
      }
      if (p.out2) {
        __nv_bfloat16* op2 = p.out2 + pix_off + (long long)cbase * p.out_sc;
        float w[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = cbase + j;
          float a = v[j];
          if (j < nv) a = fmaxf(fmaf(a, __ldg(p.scale2 + c), __ldg(p.bias2 + c)), 0.f);
          w[j] = a;
        }
        if (vec_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(w[q * 8 + e * 2], w[q * 8 + e * 2 + 1]);
            reinterpret_cast<uint4*>(op2)[q] = u;
          }
        } else {
          for (int j = 0; j < nv; ++j) op2[(long long)j * p.out_sc] = __float2bfloat16_rn(w[j]);
        }
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
               const cuuint32_t* box) {
  auto fn = get_encode_fn();
  if (!fn) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(XDET_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return XDET_OK;
}

int next_pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace
}  // namespace xdet

using namespace xdet;

extern "C" int xdet_conv2d_bf16(const void* d_in, const xdet_conv_desc* d, void* stream) {
  if (!d || !d_in) return fail(XDET_EINVAL, "null argument");
  if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->KH <= 0 || d->KW <= 0)
    return fail(XDET_EINVAL, "conv2d: non-positive dimension");
  if (d->in_cs < d->Cin || (d->in_cs % 8) != 0)
    return fail(XDET_EINVAL, "conv2d: input channel stride (%d) must be >= Cin and a multiple of 8 (TMA 16-byte strides)",
                d->in_cs);
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d->weights) & 15))
    return fail(XDET_EINVAL, "conv2d: input and weights must be 16-byte aligned");
  const int kcpt = (d->Cin + kBK - 1) / kBK;
  const int ktot = d->KH * d->KW * kcpt * kBK;

  ConvGemmArgs a{};
  // tile shape: widest power-of-two row segment that covers the output row, up to 128
  int BW = 8;
  while (BW < d->Wout && BW < kBM) BW <<= 1;
  a.BW = BW;
  a.BH = kBM / BW;
  a.tiles_x = (d->Wout + a.BW - 1) / a.BW;
  a.tiles_y = (d->Hout + a.BH - 1) / a.BH;
  a.n_img = d->N;
  a.Hout = d->Hout;
  a.Wout = d->Wout;
  a.Cout = d->Cout;
  a.taps_w = d->KW;
  a.dil_h = d->dil_h;
  a.dil_w = d->dil_w;
  a.pad_top = d->pad_top;
  a.pad_left = d->pad_left;
  a.k_chunks_per_tap = kcpt;
  a.num_k_blocks = d->KH * d->KW * kcpt;
  // N tile: 128 by default, the whole (16-aligned) Cout when it is smaller, 256 never (TMEM/epilogue balance)
  int BN = d->Cout >= 128 ? 128 : ((d->Cout + 15) / 16) * 16;
  if (d->block_n > 0) BN = d->block_n;
  if (BN % 16 != 0 || BN < 16 || BN > 256) return fail(XDET_EINVAL, "conv2d: block_n must be a multiple of 16 in [16,256]");
  a.BN = BN;
  a.tmem_cols = next_pow2_cols(((BN + 31) / 32) * 32);
  const size_t stage_bytes = (size_t)kBM * kBK * 2 + (((size_t)BN * kBK * 2 + 1023) & ~(size_t)1023);
  const size_t tail = 3 * kMaxStages * sizeof(uint64_t) + 64;
  int stages = (int)((227 * 1024 - tail - 1024) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > a.num_k_blocks) stages = a.num_k_blocks < 2 ? 2 : a.num_k_blocks;
  if (stages < 2) return fail(XDET_EINVAL, "conv2d: tile does not fit shared memory");
  a.stages = stages;
  a.scale = d->scale;
  a.bias = d->bias;
  a.relu = d->relu;
  a.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  a.out = d->out;
  a.out_fp32 = d->out_fp32;
  a.out_sn = d->out_sn;
  a.out_sy = d->out_sy;
  a.out_sx = d->out_sx;
  a.out_sc = d->out_sc;
  a.out2 = reinterpret_cast<__nv_bfloat16*>(d->out2);
  a.scale2 = d->scale2;
  a.bias2 = d->bias2;
  if (a.out2 && (!a.scale2 || !a.bias2)) return fail(XDET_EINVAL, "conv2d: out2 needs scale2 and bias2");
  if ((a.residual || a.out2) && d->out_fp32)
    return fail(XDET_EINVAL, "conv2d: residual / second output share the (bf16) layout of `out`");

  CUtensorMap map_a, map_b;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    const cuuint64_t strides[3] = {(cuuint64_t)d->in_cs * 2, (cuuint64_t)d->in_cs * 2 * d->W,
                                   (cuuint64_t)d->in_cs * 2 * d->W * d->H};
    const cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)a.BW, (cuuint32_t)a.BH, 1};
    XDET_TRY(encode_map(&map_a, d_in, 4, dims, strides, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)d->Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)BN};
    XDET_TRY(encode_map(&map_b, d->weights, 2, dims, strides, box));
  }
  const size_t smem = (size_t)stages * stage_bytes + tail + 1024;  // +1024: manual alignment slack
  XDET_TRY(check_cuda(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "cudaFuncSetAttribute(conv_gemm)"));
  dim3 grid((unsigned)(a.tiles_x * a.tiles_y * a.n_img), (unsigned)((d->Cout + BN - 1) / BN));
  conv_gemm_kernel<<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(map_a, map_b, a);
  return after_launch("conv_gemm_kernel");
}
