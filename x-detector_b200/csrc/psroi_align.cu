// PsRoiAlign forward / backward for sm_100a.
//
// Replaces the reference's TF custom op (cpp/PSROIPooling/ps_roi_align_op.{h,cc,cu},
// ps_roi_align_grad_op.{cc,cu}).  The arithmetic contract is the reference's CPU functor
// (ps_roi_align_op.cc:94-193, ps_roi_align_grad_op.cc:200-315), reproduced operation for
// operation so that results are bit-identical:
//   * RoI / bin geometry in fp32, one rounding per operation -> every mul/add/div below is an
//     explicit round-to-nearest intrinsic (no FMA contraction; the file is also built with
//     -fmad=false);
//   * 4-tap blend: the first three products carry a `1.` double literal in the reference and are
//     formed in fp64 as ((wa*wb)*P); the fourth, `fx*fy*P` (ps_roi_align_op.cc:176), is all-float
//     and formed in fp32; the terms are summed left-to-right in fp64 and rounded once to fp32;
//   * max: strict '<' from -FLT_MAX (first maximum wins); mean: fp32 running sum in sample order
//     and one fp32 divide.
// No tensor cores: this is irregular, HBM/L1-bound gather work.
#include <cfloat>
#include <cstdint>

#include "common.cuh"

namespace xdet {
namespace {

struct RoiGeom {
  float ymin, xmin;      // clipped RoI origin (feature-map pixels)
  float bin_h, bin_w;    // bin extent
  float step_h, step_w;  // sample pitch inside a bin
  int nh, nw;            // samples per bin; nh == 0 marks a degenerate RoI
};

// ps_roi_align_op.cc:123-158.  std::max(a,b) == (a<b)?b:a and std::min(a,b) == (b<a)?b:a are
// spelled out so that NaN handling is the reference's, not fmaxf's.
__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi, int H, int W, int gw, int gh) {
  RoiGeom g;
  const float r0 = roi[0], r1 = roi[1], r2 = roi[2], r3 = roi[3];
  if (r2 < FLT_MIN || r3 < FLT_MIN) {
    g.nh = 0;
    g.nw = 0;
    g.ymin = g.xmin = g.bin_h = g.bin_w = g.step_h = g.step_w = 0.f;
    return g;
  }
  const float fH = (float)H, fW = (float)W;
  const float cy = __fmul_rn(r0, fH);
  const float cx = __fmul_rn(r1, fW);
  float rh = __fmul_rn(r2, fH);
  rh = (rh < 1.0f) ? 1.0f : rh;
  float rw = __fmul_rn(r3, fW);
  rw = (rw < 1.0f) ? 1.0f : rw;
  const float hh = __fmul_rn(rh, 0.5f);  // (float)(rh / 2.) : exact for rh >= 1
  const float hw = __fmul_rn(rw, 0.5f);
  float ymin = __fsub_rn(cy, hh);
  ymin = (ymin < 0.f) ? 0.f : ymin;
  float xmin = __fsub_rn(cx, hw);
  xmin = (xmin < 0.f) ? 0.f : xmin;
  float ymax = __fadd_rn(cy, hh);
  ymax = (fH < ymax) ? fH : ymax;  // (float)H - FLT_MIN == (float)H in fp32
  float xmax = __fadd_rn(cx, hw);
  xmax = (fW < xmax) ? fW : xmax;
  g.ymin = ymin;
  g.xmin = xmin;
  g.bin_w = __fdiv_rn(__fsub_rn(xmax, xmin), (float)gw);
  g.bin_h = __fdiv_rn(__fsub_rn(ymax, ymin), (float)gh);
  g.nw = __float2int_rz(g.bin_w) + 1;
  g.nh = __float2int_rz(g.bin_h) + 1;
  g.step_w = __fdiv_rn(g.bin_w, (float)g.nw);
  g.step_h = __fdiv_rn(g.bin_h, (float)g.nh);
  return g;
}

// float(start + step*i + step/2.)  (ps_roi_align_op.cc:163-164): fp32 mul, fp32 add, then a
// double add rounded once to float.  When step/2 is a normal float the double sum of the two
// floats is either exact (exponent gap <= 29) or dominated by one addend, so the result equals a
// single fp32 add; the literal fp64 form is kept for (sub)denormal steps.
__device__ __forceinline__ float sample_coord(float start, float step, int i) {
  const float b = __fadd_rn(start, __fmul_rn(step, (float)i));
  if (fabsf(step) >= 4.0f * FLT_MIN) return __fadd_rn(b, __fmul_rn(step, 0.5f));
  return __double2float_rn(__dadd_rn((double)b, __dmul_rn((double)step, 0.5)));
}

struct Tap {
  int o00, o01, o10, o11;  // plane offsets: (iy,ix) (iy+1,ix) (iy,ix+1) (iy+1,ix+1), +1 clamped
  double w00, w01, w10;    // fp64 weights of the first three taps
  float w11;               // fp32 weight fx*fy of the fourth
};

// Per-sample weights, shared by every channel of a bin (ps_roi_align_op.cc:166-176).
__device__ __forceinline__ void sample_weights(float x, float y, int H, int W, int stride_row, Tap& t) {
  int ix = __float2int_rz(x), iy = __float2int_rz(y);
  const float fx = __fsub_rn(x, (float)ix), fy = __fsub_rn(y, (float)iy);
  const int ix1 = min(ix + 1, W - 1), iy1 = min(iy + 1, H - 1);
  ix = min(ix, W - 1);  // deviation (out-of-contract input only): clamp instead of reading past the plane
  iy = min(iy, H - 1);
  ix = max(ix, 0);
  iy = max(iy, 0);
  t.o00 = iy * stride_row + ix;
  t.o01 = iy1 * stride_row + ix;
  t.o10 = iy * stride_row + ix1;
  t.o11 = iy1 * stride_row + ix1;
  const double dfx = (double)fx, dfy = (double)fy;
  const double ax = __dsub_rn(1.0, dfx), ay = __dsub_rn(1.0, dfy);
  t.w00 = __dmul_rn(ax, ay);
  t.w01 = __dmul_rn(ax, dfy);
  t.w10 = __dmul_rn(dfx, ay);
  t.w11 = __fmul_rn(fx, fy);
}

__device__ __forceinline__ float blend(const Tap& t, float p00, float p01, float p10, float p11) {
  double s = __dmul_rn(t.w00, (double)p00);
  s = __dadd_rn(s, __dmul_rn(t.w01, (double)p01));
  s = __dadd_rn(s, __dmul_rn(t.w10, (double)p10));
  s = __dadd_rn(s, (double)__fmul_rn(t.w11, p11));
  return __double2float_rn(s);
}

// Pool one output element: `plane` points at the (image, channel) plane (global or shared),
// with `stride_row` elements between rows.
template <bool kMax>
__device__ __forceinline__ void pool_one(const float* __restrict__ plane, int stride_row, const RoiGeom& g, float x0,
                                         float y0, int H, int W, float& out_v, int& out_i) {
  float acc = kMax ? -FLT_MAX : 0.f;
  int arg = 0;
  for (int hi = 0; hi < g.nh; ++hi) {
    const float y = sample_coord(y0, g.step_h, hi);
    for (int wi = 0; wi < g.nw; ++wi) {
      const float x = sample_coord(x0, g.step_w, wi);
      Tap t;
      sample_weights(x, y, H, W, stride_row, t);
      const float v = blend(t, plane[t.o00], plane[t.o01], plane[t.o10], plane[t.o11]);
      if (kMax) {
        if (acc < v) {
          acc = v;
          arg = g.nw * hi + wi;
        }
      } else {
        acc = __fadd_rn(acc, v);
      }
    }
  }
  if (!kMax) acc = __fdiv_rn(acc, (float)(g.nh * g.nw));
  out_v = acc;
  out_i = arg;
}

// ------------------------------------------------------------------------------------------
// Variant GATHER: one thread per output element, taps gathered straight from global memory
// (read-only path, L1/L2 cached).  Works for any shape; the only path for maps whose planes do
// not fit shared memory.  Adjacent threads = adjacent output channels -> coalesced stores.
// ------------------------------------------------------------------------------------------
template <bool kMax>
__global__ void __launch_bounds__(256) psroi_fwd_gather_kernel(const float* __restrict__ inputs,
                                                               const float* __restrict__ rois,
                                                               float* __restrict__ pooled, int32_t* __restrict__ index,
                                                               int C, int H, int W, int R, int gw, int gh,
                                                               long long total) {
  const int G = gw * gh, bank = C / G;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % C);
    const long long pr = e / C;  // image * R + roi
    const int img = (int)(pr / R);
    const int bin = c / bank;
    const int row = bin / gw, col = bin - row * gw;
    const RoiGeom g = roi_geometry(rois + pr * 4, H, W, gw, gh);
    float v = 0.f;
    int a = 0;
    if (g.nh > 0) {
      const float x0 = __fadd_rn(g.xmin, __fmul_rn(g.bin_w, (float)col));
      const float y0 = __fadd_rn(g.ymin, __fmul_rn(g.bin_h, (float)row));
      const float* plane = inputs + ((long long)img * C + c) * H * W;
      pool_one<kMax>(plane, W, g, x0, y0, H, W, v, a);
    }
    pooled[e] = v;
    index[e] = a;
  }
}

// ------------------------------------------------------------------------------------------
// Variant PLANES: the feature map of a detector head is tiny (490x30x30 fp32 = 1.76 MB) and
// lives in L2; what the op really moves is its OUTPUT (8 B per element).  Each CTA therefore
// pins a channel slice (whole bins: `bins_per_cta * bank` planes) in shared memory once and
// streams a range of RoIs past it, so every tap is a shared-memory read and HBM sees only the
// map once plus the output stream.
//   grid = (slices, roi_splits, N); block = 256 threads
//   smem = slice planes [cs][H*W + pad] fp32 (odd plane pitch -> conflict-free across channels)
//          + double-buffered per-RoI geometry for a chunk of kChunk RoIs
// Inside a chunk the work items (roi, channel) are laid out channel-fastest so that a warp's
// stores are contiguous runs of `cs` floats per RoI.
// ------------------------------------------------------------------------------------------
constexpr int kPlanesThreads = 256;
constexpr int kChunk = 32;

template <bool kMax>
__global__ void __launch_bounds__(kPlanesThreads) psroi_fwd_planes_kernel(
    const float* __restrict__ inputs, const float* __restrict__ rois, float* __restrict__ pooled,
    int32_t* __restrict__ index, int C, int H, int W, int R, int gw, int gh, int bins_per_cta, int plane_pitch,
    int rois_per_split) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = gw * gh, bank = C / G, HW = H * W;
  const int bin0 = blockIdx.x * bins_per_cta;
  const int nbins = min(bins_per_cta, G - bin0);
  const int cs = nbins * bank;  // channels in this CTA's slice
  const int c0 = bin0 * bank;
  const int img = blockIdx.z;
  const int r_begin = blockIdx.y * rois_per_split;
  const int r_end = min(R, r_begin + rois_per_split);

  float* planes = reinterpret_cast<float*>(smem_raw);
  RoiGeom* geom = reinterpret_cast<RoiGeom*>(smem_raw + (size_t)bins_per_cta * bank * plane_pitch * sizeof(float));

  // Stage the slice: global [cs][HW] contiguous -> shared [cs][plane_pitch].
  {
    const float* src = inputs + ((long long)img * C + c0) * HW;
    const int n = cs * HW;
    for (int i = threadIdx.x; i < n; i += kPlanesThreads) {
      const int ch = i / HW, p = i - ch * HW;
      planes[ch * plane_pitch + p] = __ldg(src + i);
    }
  }

  const int items = kChunk * cs;
  int buf = 0;
  for (int rc = r_begin; rc < r_end; rc += kChunk, buf ^= 1) {
    RoiGeom* gbuf = geom + buf * kChunk;
    if (threadIdx.x < kChunk) {
      const int r = rc + threadIdx.x;
      if (r < r_end) gbuf[threadIdx.x] = roi_geometry(rois + ((long long)img * R + r) * 4, H, W, gw, gh);
    }
    __syncthreads();  // geometry visible (and, first time round, the staged planes)
    for (int it = threadIdx.x; it < items; it += kPlanesThreads) {
      const int rl = it / cs, ch = it - rl * cs;
      const int r = rc + rl;
      if (r >= r_end) break;
      const RoiGeom g = gbuf[rl];
      const int bin = bin0 + ch / bank;
      const int row = bin / gw, col = bin - row * gw;
      float v = 0.f;
      int a = 0;
      if (g.nh > 0) {
        const float x0 = __fadd_rn(g.xmin, __fmul_rn(g.bin_w, (float)col));
        const float y0 = __fadd_rn(g.ymin, __fmul_rn(g.bin_h, (float)row));
        pool_one<kMax>(planes + ch * plane_pitch, W, g, x0, y0, H, W, v, a);
      }
      const long long o = ((long long)img * R + r) * C + c0 + ch;
      pooled[o] = v;
      index[o] = a;
    }
    // no second barrier: the next chunk writes the other geometry buffer, and a thread can only
    // reach the chunk after that (overwriting this buffer) by passing the next __syncthreads().
  }
}

// ------------------------------------------------------------------------------------------
// Backward.  Bit-exactness fixes the order in which a map cell receives its additions: the
// reference CPU functor (ps_roi_align_grad_op.cc:212-311) walks, for each cell, the RoIs in index
// order and inside a RoI the samples (h-major) and taps (00, +row, +col, +row+col).  Different
// planes are independent, so: one THREAD owns one (image, channel) plane and walks the RoIs
// serially; a CTA (one warp) owns up to 32 consecutive planes whose accumulators live in shared
// memory [planes][H*W+pad], are zero-filled and finally written out cooperatively (the planes of
// a CTA are contiguous in global memory, so those stores are fully coalesced).  This is also
// run-to-run deterministic, unlike the reference's atomicAdd kernel (ps_roi_align_grad_op.cu:100-135).
//   max : only the arg-max sample of each (roi, channel) scatters (:259-286)
//   mean: an fp32 partial per cell over the RoI's samples, divided by nh*nw, then added (:287-309);
//         the partials use a second shared plane per thread, cleared over the touched rectangle.
// ------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 32;

template <bool kMax>
__global__ void __launch_bounds__(kBwdThreads) psroi_bwd_kernel(const float* __restrict__ rois,
                                                                const float* __restrict__ gout,
                                                                const int32_t* __restrict__ index,
                                                                float* __restrict__ gin, int C, int H, int W, int R,
                                                                int gw, int gh, int planes_per_cta, int pitch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* acc_all = reinterpret_cast<float*>(smem_raw);
  const int G = gw * gh, bank = C / G, HW = H * W;
  const int img = blockIdx.y;
  const int c_begin = blockIdx.x * planes_per_cta;
  const int np = min(planes_per_cta, C - c_begin);
  const int nbuf = kMax ? 1 : 2;
  for (int i = threadIdx.x; i < planes_per_cta * nbuf * pitch; i += kBwdThreads) acc_all[i] = 0.f;
  __syncwarp();

  if ((int)threadIdx.x < np) {
    const int c = c_begin + threadIdx.x;
    const int bin = c / bank, row = bin / gw, col = bin - row * gw;
    float* plane = acc_all + threadIdx.x * pitch;
    float* part = acc_all + (planes_per_cta + threadIdx.x) * pitch;  // mean only
    for (int r = 0; r < R; ++r) {
      const RoiGeom g = roi_geometry(rois + ((long long)img * R + r) * 4, H, W, gw, gh);
      if (g.nh == 0) continue;
      const long long o = ((long long)img * R + r) * C + c;
      const float gv = gout[o];
      const double dg = (double)gv;
      const float x0 = __fadd_rn(g.xmin, __fmul_rn(g.bin_w, (float)col));
      const float y0 = __fadd_rn(g.ymin, __fmul_rn(g.bin_h, (float)row));
      const int s_lo = kMax ? index[o] : 0;
      const int s_hi = kMax ? s_lo + 1 : g.nh * g.nw;
      int y_lo = H, y_hi = -1, x_lo = W, x_hi = -1;
      float* dst = kMax ? plane : part;
      for (int s = s_lo; s < s_hi; ++s) {
        const int hi = s / g.nw, wi = s - hi * g.nw;
        const float x = sample_coord(x0, g.step_w, wi), y = sample_coord(y0, g.step_h, hi);
        const int ix = __float2int_rz(x), iy = __float2int_rz(y);
        const float fx = __fsub_rn(x, (float)ix), fy = __fsub_rn(y, (float)iy);
        const int ix1 = min(ix + 1, W - 1), iy1 = min(iy + 1, H - 1);
        const double dfx = (double)fx, dfy = (double)fy;
        const double ax = __dsub_rn(1.0, dfx), ay = __dsub_rn(1.0, dfy);
        const float t00 = __double2float_rn(__dmul_rn(__dmul_rn(ax, ay), dg));
        const float t01 = __double2float_rn(__dmul_rn(__dmul_rn(ax, dfy), dg));
        const float t10 = __double2float_rn(__dmul_rn(__dmul_rn(dfx, ay), dg));
        const float t11 = __fmul_rn(__fmul_rn(fx, fy), gv);  // all-float in the reference (:280-282)
        // a tap lands only where a map cell equals its (row, col) (:271-282); ix1/iy1 are clamped
        const bool xin = ix >= 0 && ix < W, yin = iy >= 0 && iy < H;
        if (xin && yin) dst[iy * W + ix] = __fadd_rn(dst[iy * W + ix], t00);
        if (xin) dst[iy1 * W + ix] = __fadd_rn(dst[iy1 * W + ix], t01);
        if (yin) dst[iy * W + ix1] = __fadd_rn(dst[iy * W + ix1], t10);
        dst[iy1 * W + ix1] = __fadd_rn(dst[iy1 * W + ix1], t11);
        if (!kMax) {
          y_lo = min(y_lo, yin ? iy : iy1);
          y_hi = max(y_hi, iy1);
          x_lo = min(x_lo, xin ? ix : ix1);
          x_hi = max(x_hi, ix1);
        }
      }
      if (!kMax) {
        // Cells of the rectangle no tap touched hold +0 partials: adding (+0/cnt) leaves the
        // plane value unchanged bit-for-bit (x + 0 == x; the plane never holds -0 because it
        // starts at +0 and (+0) + (-0) == +0), so sweeping the whole rectangle is exact.
        const float cnt = (float)(g.nh * g.nw);
        for (int yy = y_lo; yy <= y_hi; ++yy)
          for (int xx = x_lo; xx <= x_hi; ++xx) {
            const float a = part[yy * W + xx];
            if (a != 0.f) plane[yy * W + xx] = __fadd_rn(plane[yy * W + xx], __fdiv_rn(a, cnt));
            part[yy * W + xx] = 0.f;
          }
      }
    }
  }
  __syncwarp();
  float* out = gin + ((long long)img * C + c_begin) * HW;
  for (int i = threadIdx.x; i < np * HW; i += kBwdThreads) {
    const int p = i / HW, q = i - p * HW;
    out[i] = acc_all[p * pitch + q];
  }
}

int validate(int N, int C, int H, int W, int R, int gw, int gh) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || R < 0) return fail(XDET_EINVAL, "negative dimension");
  if (gw <= 0 || gh <= 0) return fail(XDET_EINVAL, "Need Attr grid_dim_width/height > 0, got %d x %d", gw, gh);
  if (C % (gw * gh) != 0)
    return fail(XDET_EINVAL, "channels (%d) must be divisible by grid_dim_width*grid_dim_height (%d)", C, gw * gh);
  return XDET_OK;
}

constexpr size_t kMaxSmem = 227 * 1024;

// Largest number of whole bins whose planes (+geometry buffers) fit in shared memory; 0 if none.
int planes_bins_per_cta(int bank, int H, int W, int G, int* pitch_out, size_t* smem_out) {
  const int HW = H * W;
  const int pitch = (HW % 2 == 0) ? HW + 1 : HW;  // odd pitch: channels land in distinct banks
  const size_t geom_bytes = 2 * kChunk * sizeof(RoiGeom);
  const size_t per_bin = (size_t)bank * pitch * sizeof(float);
  if (per_bin + geom_bytes > kMaxSmem) return 0;
  // One bin per CTA keeps the slice small (more CTAs per SM, more slices to spread over the 148
  // SMs); tiny banks (15x15 grids: bank 4) are grouped until a slice is at least ~32 channels so
  // that a RoI's output run stays a reasonable store width.
  int bins = 1;
  while (bins < G && (bins * bank) < 32 && (size_t)(bins + 1) * per_bin + geom_bytes <= kMaxSmem) ++bins;
  *pitch_out = pitch;
  *smem_out = (size_t)bins * per_bin + geom_bytes;
  return bins;
}

template <bool kMax>
int launch_fwd(const float* in, const float* rois, float* pooled, int32_t* index, int N, int C, int H, int W, int R,
               int gw, int gh, int variant, cudaStream_t st) {
  const long long total = (long long)N * R * C;
  if (total == 0) return XDET_OK;
  const int G = gw * gh, bank = C / G;
  int pitch = 0;
  size_t smem = 0;
  const int bins = (bank > 0 && H > 0 && W > 0) ? planes_bins_per_cta(bank, H, W, G, &pitch, &smem) : 0;
  if (variant == XDET_PSROI_PLANES && bins == 0)
    return fail(XDET_EINVAL, "PLANES variant needs bank*H*W*4 B (=%zu) to fit shared memory",
                (size_t)bank * H * W * 4);
  if (variant == XDET_PSROI_AUTO) variant = (bins > 0 && (long long)R * N >= 64) ? XDET_PSROI_PLANES : XDET_PSROI_GATHER;

  if (variant == XDET_PSROI_PLANES) {
    const int slices = (G + bins - 1) / bins;
    // Enough RoI splits to give every SM ~2 CTAs, but never fewer than kChunk RoIs per split.
    const int ctas_per_sm = (int)(kMaxSmem / smem) >= 2 ? 2 : 1;
    int splits = (kNumSMs * ctas_per_sm + slices * N - 1) / (slices * N);
    const int max_splits = (R + kChunk - 1) / kChunk;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int per = (R + splits - 1) / splits;
    per = (per + kChunk - 1) / kChunk * kChunk;
    splits = (R + per - 1) / per;
    auto kern = psroi_fwd_planes_kernel<kMax>;
    XDET_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(psroi_fwd_planes)"));
    dim3 grid(slices, splits, N);
    kern<<<grid, kPlanesThreads, smem, st>>>(in, rois, pooled, index, C, H, W, R, gw, gh, bins, pitch, per);
    return after_launch("psroi_fwd_planes_kernel");
  }
  if (variant != XDET_PSROI_GATHER) return fail(XDET_EINVAL, "unknown PsRoiAlign variant %d", variant);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
  psroi_fwd_gather_kernel<kMax><<<(unsigned)blocks, 256, 0, st>>>(in, rois, pooled, index, C, H, W, R, gw, gh, total);
  return after_launch("psroi_fwd_gather_kernel");
}

}  // namespace
}  // namespace xdet

using namespace xdet;

extern "C" int xdet_psroi_align_fwd_ex(const float* d_inputs, const float* d_rois, float* d_pooled, int32_t* d_index,
                                       int N, int C, int H, int W, int R, int gw, int gh, int use_max, int variant,
                                       void* stream) {
  XDET_TRY(validate(N, C, H, W, R, gw, gh));
  cudaStream_t st = (cudaStream_t)stream;
  return use_max ? launch_fwd<true>(d_inputs, d_rois, d_pooled, d_index, N, C, H, W, R, gw, gh, variant, st)
                 : launch_fwd<false>(d_inputs, d_rois, d_pooled, d_index, N, C, H, W, R, gw, gh, variant, st);
}

extern "C" int xdet_psroi_align_fwd(const float* d_inputs, const float* d_rois, float* d_pooled, int32_t* d_index,
                                    int N, int C, int H, int W, int R, int gw, int gh, int use_max, void* stream) {
  return xdet_psroi_align_fwd_ex(d_inputs, d_rois, d_pooled, d_index, N, C, H, W, R, gw, gh, use_max,
                                 XDET_PSROI_AUTO, stream);
}

template <bool kMax>
static int launch_bwd(const float* rois, const float* gout, const int32_t* idx, float* gin, int N, int C, int H, int W,
                      int R, int gw, int gh, cudaStream_t st) {
  const int HW = H * W;
  const int pitch = (HW % 2 == 0) ? HW + 1 : HW;
  const size_t per_plane = (size_t)(kMax ? 1 : 2) * pitch * sizeof(float);
  int ppc = (int)(kMaxSmem / per_plane);
  if (ppc < 1) return fail(XDET_EINVAL, "PsRoiAlignGrad: a %dx%d plane does not fit shared memory", H, W);
  if (ppc > kBwdThreads) ppc = kBwdThreads;
  // spread the planes over at least ~2 waves of CTAs when there are enough of them
  while (ppc > 4 && (long long)((C + ppc - 1) / ppc) * N < 2 * kNumSMs) ppc /= 2;
  const size_t smem = (size_t)ppc * per_plane;
  auto kern = psroi_bwd_kernel<kMax>;
  XDET_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "cudaFuncSetAttribute(psroi_bwd)"));
  dim3 grid((C + ppc - 1) / ppc, N);
  kern<<<grid, kBwdThreads, smem, st>>>(rois, gout, idx, gin, C, H, W, R, gw, gh, ppc, pitch);
  return after_launch("psroi_bwd_kernel");
}

extern "C" int xdet_psroi_align_bwd(const float* d_rois, const float* d_pooled_grad, const int32_t* d_index,
                                    float* d_grad, int N, int C, int H, int W, int R, int gw, int gh, int use_max,
                                    void* stream) {
  XDET_TRY(validate(N, C, H, W, R, gw, gh));
  if ((long long)N * C == 0 || H == 0 || W == 0) return XDET_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return use_max ? launch_bwd<true>(d_rois, d_pooled_grad, d_index, d_grad, N, C, H, W, R, gw, gh, st)
                 : launch_bwd<false>(d_rois, d_pooled_grad, d_index, d_grad, N, C, H, W, R, gw, gh, st);
}

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t n) { return check_cuda(cudaMalloc(&p, n ? n : 1), "cudaMalloc"); }
};
}  // namespace

extern "C" int xdet_psroi_align_fwd_host(const float* h_inputs, const float* h_rois, float* h_pooled, int32_t* h_index,
                                         int N, int C, int H, int W, int R, int gw, int gh, int use_max) {
  XDET_TRY(validate(N, C, H, W, R, gw, gh));
  const size_t n_in = (size_t)N * C * H * W, n_roi = (size_t)N * R * 4, n_out = (size_t)N * R * C;
  DevBuf in, roi, out, idx;
  XDET_TRY(in.alloc(n_in * 4));
  XDET_TRY(roi.alloc(n_roi * 4));
  XDET_TRY(out.alloc(n_out * 4));
  XDET_TRY(idx.alloc(n_out * 4));
  XDET_TRY(check_cuda(cudaMemcpy(in.p, h_inputs, n_in * 4, cudaMemcpyHostToDevice), "H2D inputs"));
  XDET_TRY(check_cuda(cudaMemcpy(roi.p, h_rois, n_roi * 4, cudaMemcpyHostToDevice), "H2D rois"));
  XDET_TRY(xdet_psroi_align_fwd((const float*)in.p, (const float*)roi.p, (float*)out.p, (int32_t*)idx.p, N, C, H, W, R,
                                gw, gh, use_max, nullptr));
  XDET_TRY(check_cuda(cudaMemcpy(h_pooled, out.p, n_out * 4, cudaMemcpyDeviceToHost), "D2H pooled"));
  XDET_TRY(check_cuda(cudaMemcpy(h_index, idx.p, n_out * 4, cudaMemcpyDeviceToHost), "D2H index"));
  return XDET_OK;
}

extern "C" int xdet_psroi_align_bwd_host(const float* h_rois, const float* h_pooled_grad, const int32_t* h_index,
                                         float* h_grad, int N, int C, int H, int W, int R, int gw, int gh,
                                         int use_max) {
  XDET_TRY(validate(N, C, H, W, R, gw, gh));
  const size_t n_in = (size_t)N * C * H * W, n_roi = (size_t)N * R * 4, n_out = (size_t)N * R * C;
  DevBuf roi, g, idx, gin;
  XDET_TRY(roi.alloc(n_roi * 4));
  XDET_TRY(g.alloc(n_out * 4));
  XDET_TRY(idx.alloc(n_out * 4));
  XDET_TRY(gin.alloc(n_in * 4));
  XDET_TRY(check_cuda(cudaMemcpy(roi.p, h_rois, n_roi * 4, cudaMemcpyHostToDevice), "H2D rois"));
  XDET_TRY(check_cuda(cudaMemcpy(g.p, h_pooled_grad, n_out * 4, cudaMemcpyHostToDevice), "H2D grad"));
  XDET_TRY(check_cuda(cudaMemcpy(idx.p, h_index, n_out * 4, cudaMemcpyHostToDevice), "H2D index"));
  XDET_TRY(xdet_psroi_align_bwd((const float*)roi.p, (const float*)g.p, (const int32_t*)idx.p, (float*)gin.p, N, C, H,
                                W, R, gw, gh, use_max, nullptr));
  XDET_TRY(check_cuda(cudaMemcpy(h_grad, gin.p, n_in * 4, cudaMemcpyDeviceToHost), "D2H grad"));
  return XDET_OK;
}
