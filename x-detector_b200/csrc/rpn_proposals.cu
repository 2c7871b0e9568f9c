// RPN proposal generation on the GPU (the reference pins this whole stage to /cpu:0,
// net/xception_body.py:425, and it is the largest stall of its step: ~100 ms/img).
//
//   xdet_rpn_decode   rpn head outputs -> objectness + decoded boxes
//                       (light_head_rfcn_eval.py:389-399, preprocessing/anchor_manipulator.py:641-669)
//   xdet_rpn_select   clip -> filter -> top-k -> NMS -> upsample -> RoIs (+ their (cy,cx,h,w) form)
//                       (net/xception_body.py:402-444 with _bboxes_clip :173-194, _filter_and_sort_boxes
//                        :133-158, _bboxes_nms :57-67, _upsample_rois :196-213, _point2center :215-218)
//
// The selection stage is integer/compare work on fp32 values with one rounding per operation (this
// file is built without FMA contraction and uses explicit _rn intrinsics), so that, given the same
// scores and boxes, the selected set is bit-identical to the CPU restatement (oracle/proposals.py):
//   * top-k: 64-bit keys (score bits << 32 | ~index) -> descending order, ties to the lower index
//     (tf.nn.top_k); radix-select of the k-th key, compaction into shared memory, bitonic sort;
//   * NMS (TF r1.6 NonMaxSuppressionV2): IoU on min/max-normalised corners, 0 if an area <= 0,
//     suppress iff IoU > threshold; all-pairs suppression bit-matrix spread over the whole GPU, then
//     one CTA per image resolves 64 candidates at a time (sequential only inside the 64x64 diagonal
//     block, survivors' rows OR-reduced in parallel);
//   * upsample: tile + "random" remainder, with tf.random_shuffle replaced by an injected key array
//     (stable argsort of the first n keys) so that both sides can be driven identically.
// Neither HBM- nor tensor-bound: latency / dependency bound, reported as us per image.
#include <cfloat>
#include <cstdint>

#include "common.cuh"

namespace xdet {
namespace {

constexpr int kSelThreads = 1024;
constexpr int kDetThreads = 256;  // CTA size of the per-(image, class) selections of xdet_det_postprocess

// ---- decode ------------------------------------------------------------------------------------
// rpn_out: [N, h, w, ch_stride] fp32 with the 2A class logits at channel cls_off + a*2 + {0,1} and the
// 4A box deltas at channel box_off + a*4 + {cy,cx,h,w}.  Anchor order is (y, x, a), a fastest.
__global__ void __launch_bounds__(256) rpn_decode_kernel(const float* __restrict__ rpn_out, int ch_stride, int cls_off,
                                                         int box_off, const float* __restrict__ yref,
                                                         const float* __restrict__ xref,
                                                         const float* __restrict__ href,
                                                         const float* __restrict__ wref, int hw, int A,
                                                         float* __restrict__ scores, float* __restrict__ boxes,
                                                         long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int a = (int)(e % A);
  const long long cell = e / A;  // n*hw + (y*w + x)
  const int pos = (int)(cell % hw);
  const float* px = rpn_out + cell * ch_stride;
  const float z0 = __ldg(px + cls_off + a * 2), z1 = __ldg(px + cls_off + a * 2 + 1);
  // tf.nn.softmax(...)[:, -1]: exp(z - max) / sum
  const float m = fmaxf(z0, z1);
  const float e0 = expf(__fsub_rn(z0, m)), e1 = expf(__fsub_rn(z1, m));
  scores[e] = __fdiv_rn(e1, __fadd_rn(e0, e1));
  const float t0 = __ldg(px + box_off + a * 4), t1 = __ldg(px + box_off + a * 4 + 1);
  const float t2 = __ldg(px + box_off + a * 4 + 2), t3 = __ldg(px + box_off + a * 4 + 3);
  const float ha = __ldg(href + a), wa = __ldg(wref + a);
  const float ph = __fmul_rn(expf(t2), ha), pw = __fmul_rn(expf(t3), wa);
  const float cy = __fadd_rn(__fmul_rn(t0, ha), __ldg(yref + pos));
  const float cx = __fadd_rn(__fmul_rn(t1, wa), __ldg(xref + pos));
  const float hh = __fmul_rn(ph, 0.5f), hw2 = __fmul_rn(pw, 0.5f);  // x / 2. is exact halving
  float4 b;
  b.x = __fsub_rn(cy, hh);
  b.y = __fsub_rn(cx, hw2);
  b.z = __fadd_rn(cy, hh);
  b.w = __fadd_rn(cx, hw2);
  reinterpret_cast<float4*>(boxes)[e] = b;
}

// ---- clip + filter -------------------------------------------------------------------------------
__device__ __forceinline__ float4 clip_box(float4 b) {  // _bboxes_clip against [0,0,1,1]
  float ymin = fmaxf(b.x, 0.f), xmin = fmaxf(b.y, 0.f);
  const float ymax = fminf(b.z, 1.f), xmax = fminf(b.w, 1.f);
  ymin = fminf(ymin, ymax);
  xmin = fminf(xmin, xmax);
  return make_float4(ymin, xmin, ymax, xmax);
}

__device__ __forceinline__ bool keep_box(float4 c, float min_size) {  // _filter_and_sort_boxes mask
  const float ws = __fsub_rn(c.w, c.y), hs = __fsub_rn(c.z, c.x);
  const float xc = __fadd_rn(c.y, __fmul_rn(ws, 0.5f)), yc = __fadd_rn(c.x, __fmul_rn(hs, 0.5f));
  return ws > min_size && hs > min_size && xc > 0.f && yc > 0.f && xc < 1.f && yc < 1.f;
}

// One CTA per image: keys -> radix-select the K-th largest -> compact -> bitonic sort -> sorted outputs.
__global__ void __launch_bounds__(kSelThreads) rpn_topk_kernel(const float* __restrict__ scores,
                                                               const float* __restrict__ boxes, int A_tot, int K,
                                                               int P /* pow2 >= K */, float min_size,
                                                               int prefiltered /* 1: candidates are valid iff score > 0,
                                                                                  boxes are used as they are */,
                                                               unsigned long long* __restrict__ keys_ws,
                                                               float* __restrict__ top_scores,
                                                               float* __restrict__ top_boxes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* sk = reinterpret_cast<unsigned long long*>(smem_raw);  // [P]
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_count, s_nvalid;
  const int img = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const float* sc = scores + (long long)img * A_tot;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (long long)img * A_tot;
  unsigned long long* keys = keys_ws + (long long)img * A_tot;

  const bool direct = A_tot <= P;  // few candidates: sort them all in shared memory, no selection pass
  if (tid == 0) s_nvalid = 0;
  if (direct)
    for (int i = tid; i < P; i += nthr) sk[i] = 0ull;
  __syncthreads();
  int local_valid = 0;
  for (int i = tid; i < A_tot; i += nthr) {
    const float s = sc[i];
    unsigned long long key = 0ull;
    // a non-positive score can never outlive _upsample_rois
    if (s > 0.f && (prefiltered || keep_box(clip_box(bx[i]), min_size))) {
      key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
      ++local_valid;
    }
    if (direct) sk[i] = key; else keys[i] = key;
  }
  atomicAdd(&s_nvalid, local_valid);
  __syncthreads();
  const int keff = min(K, s_nvalid);
  if (keff == 0) {  // nothing to select (e.g. a class without detections): the zero padding is the whole answer
    for (int j = tid; j < K; j += nthr) {
      top_scores[(long long)img * K + j] = 0.f;
      reinterpret_cast<float4*>(top_boxes)[(long long)img * K + j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }

  if (!direct) {
    // radix select: find the keff-th largest key (keys are unique, 0 = filtered out)
    unsigned long long thr = 1ull;  // s_nvalid <= K: every valid key is selected, no selection pass needed
    if (s_nvalid > K) {
      if (tid == 0) {
        s_prefix = 0ull;
        s_remaining = keff;
        s_count = 0;  // (doubles as the "resolved far enough" flag of the pass loop)
      }
      __syncthreads();
      // from the top byte down, but only until the candidates fit the sort buffer: after a pass G keys lie above the
      // threshold bucket (all selected) and E inside it; once G + E <= P the bucket is simply sorted along (its surplus
      // lands behind position keff).  Typically one or two passes instead of eight; four loads in flight per thread.
      constexpr int U = 4;
      for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += nthr) hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int i0 = tid; i0 < A_tot; i0 += U * nthr) {
          unsigned long long kk[U];
#pragma unroll
          for (int u = 0; u < U; ++u) kk[u] = keys[min(i0 + u * nthr, A_tot - 1)];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const unsigned long long k = kk[u];
            const bool match = (pass == 0) || ((k >> (shift + 8)) == (prefix >> (shift + 8)));
            if (i0 + u * nthr < A_tot && match && k != 0ull) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
          }
        }
        __syncthreads();
        if (tid == 0) {
          int rem = s_remaining;
          int b = 255;
          for (; b > 0; --b) {
            if ((int)hist[b] >= rem) break;
            rem -= (int)hist[b];
          }
          s_prefix = prefix | ((unsigned long long)b << shift);
          s_remaining = rem;
          if ((keff - rem) + (int)hist[b] <= P) s_count = 1;
        }
        __syncthreads();
        const int done = s_count;
        __syncthreads();
        if (done) break;
      }
      thr = s_prefix;  // (low bits zero when the loop left early: the whole bucket qualifies)
      if (thr == 0ull) thr = 1ull;  // never the filtered-out keys
    }

    // compact the selected keys into shared memory, pad with zeros
    if (tid == 0) s_count = 0;
    for (int i = tid; i < P; i += nthr) sk[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < A_tot; i += nthr) {
      const unsigned long long k = keys[i];
      if (k != 0ull && k >= thr) sk[atomicAdd(&s_count, 1)] = k;
    }
    __syncthreads();
  }
  // sort descending
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (P >> 1); t += nthr) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);  // first half of each `size` block descending
        const unsigned long long a = sk[lo], b = sk[hi];
        if ((a < b) == desc) {
          sk[lo] = b;
          sk[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  // sorted outputs (zero padded to K, as _pad_axis does)
  for (int j = tid; j < K; j += nthr) {
    const unsigned long long k = sk[j];
    float s = 0.f;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k != 0ull) {
      const unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
      s = __uint_as_float((unsigned)(k >> 32));
      c = prefiltered ? bx[idx] : clip_box(bx[idx]);
    }
    top_scores[(long long)img * K + j] = s;
    reinterpret_cast<float4*>(top_boxes)[(long long)img * K + j] = c;
  }
}

// ---- NMS -------------------------------------------------------------------------------------------
struct NBox {
  float ymin, xmin, ymax, xmax, area;
};
__device__ __forceinline__ NBox norm_box(float4 b) {
  NBox n;
  n.ymin = fminf(b.x, b.z);
  n.xmin = fminf(b.y, b.w);
  n.ymax = fmaxf(b.x, b.z);
  n.xmax = fmaxf(b.y, b.w);
  n.area = __fmul_rn(__fsub_rn(n.ymax, n.ymin), __fsub_rn(n.xmax, n.xmin));
  return n;
}
// iou > thr with the reference's roundings (TF r1.6 NonMaxSuppressionV2: one fp32 division, then the compare).
// The division is only executed when the cheap bracket inter vs thr*union*(1 +- 1e-6) cannot decide: fl(a/b)
// differs from a/b by <= 2^-24 relative and fl(thr*union) likewise, so outside the bracket the two tests agree.
__device__ __forceinline__ bool iou_greater(const NBox& a, const NBox& b, float thr) {
  if (a.area <= 0.f || b.area <= 0.f) return false;
  const float iy0 = fmaxf(a.ymin, b.ymin), ix0 = fmaxf(a.xmin, b.xmin);
  const float iy1 = fminf(a.ymax, b.ymax), ix1 = fminf(a.xmax, b.xmax);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(iy1, iy0), 0.f), fmaxf(__fsub_rn(ix1, ix0), 0.f));
  const float uni = __fsub_rn(__fadd_rn(a.area, b.area), inter);
  if (thr >= 0.f && uni > 0.f) {
    if (inter <= 0.f) return false;  // iou == 0
    const float t = __fmul_rn(thr, uni);
    if (inter > __fmul_rn(t, 1.000001f)) return true;
    if (inter < __fmul_rn(t, 0.999999f)) return false;
  }
  return __fdiv_rn(inter, uni) > thr;
}

// grid (col blocks, row blocks, N), 64 threads: word (row i, col block) of the suppression matrix,
// bit j set iff box (cb*64+j) with j-index > i overlaps box i above the threshold.
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ top_boxes, int K, int words,
                                                      float thr, unsigned long long* __restrict__ mask) {
  const int cb = blockIdx.x, rb = blockIdx.y, img = blockIdx.z;
  if (cb < rb) return;  // only j > i matters
  __shared__ NBox cols[64];
  const float4* bx = reinterpret_cast<const float4*>(top_boxes) + (long long)img * K;
  const int cj = cb * 64 + threadIdx.x;
  if (cj < K) cols[threadIdx.x] = norm_box(bx[cj]);
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= K) return;
  const NBox me = norm_box(bx[i]);
  unsigned long long w = 0ull;
  const int ncol = min(64, K - cb * 64);
  for (int j = 0; j < ncol; ++j) {
    if (cb * 64 + j > i && iou_greater(cols[j], me, thr)) w |= (1ull << j);
  }
  mask[((long long)img * K + i) * words + cb] = w;
}

// One CTA per image: greedy scan in score order, then _upsample_rois and _point2center.
__global__ void __launch_bounds__(kSelThreads) nms_scan_kernel(
    const float* __restrict__ top_scores, const float* __restrict__ top_boxes,
    const unsigned long long* __restrict__ mask, int K, int words, int keep_n,
    const float* __restrict__ shuffle_keys /* [N, keep_n] or NULL */, int zero_pad /* 1: no _upsample_rois */,
    float* __restrict__ rois /* [N,keep_n,4] */,
    float* __restrict__ rois_yxhw /* [N,keep_n,4] or NULL */, float* __restrict__ roi_scores /* [N,keep_n] or NULL */,
    int* __restrict__ nms_keep_idx /* [N,keep_n] positions in the sorted list, -1 padded, or NULL */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* removed = reinterpret_cast<unsigned long long*>(smem_raw);  // [words]
  int* kept = reinterpret_cast<int*>(removed + words);                            // [keep_n]
  int* shuf = kept + keep_n;                                                      // [keep_n]
  const int img = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const unsigned long long* m = mask + (long long)img * K * words;
  const float* sc = top_scores + (long long)img * K;
  const float4* bx = reinterpret_cast<const float4*>(top_boxes) + (long long)img * K;

  for (int i = tid; i < words; i += nthr) removed[i] = 0ull;
  __syncthreads();
  const int out_size = min(keep_n, K);
  // Greedy scan, 64 candidates (one word column "cb") per iteration, software-pipelined over the CTA's warps:
  //   warp 0 (resolver) settles block cb -- the 64 diagonal words live in its registers (prefetched one block ahead),
  //     the sequential resolution runs on shuffles -- and at once ORs the survivors' words of column cb+1 into
  //     removed[cb+1], so that the next block can start;
  //   the other warps (workers) meanwhile OR the rows of block cb-1's survivors into the columns >= cb+1.
  // One CTA barrier per block.  removed[cb] is complete when block cb is resolved: block cb-1 contributed through the
  // resolver in the previous iteration, blocks <= cb-2 through the workers of iterations <= cb-1.
  const int warp_id = tid >> 5, lane_id = tid & 31;
  const int nworkers = nthr - 32;
  __shared__ unsigned long long s_surv2[2];
  int nk = 0;  // every thread tracks the kept count in a register
  unsigned long long d0 = 0ull, d1 = 0ull, e0 = 0ull, e1 = 0ull;  // diagonal words / words of the NEXT column of block cb
  if (warp_id == 0) {
    d0 = lane_id < K ? m[(long long)lane_id * words] : 0ull;
    d1 = lane_id + 32 < K ? m[(long long)(lane_id + 32) * words] : 0ull;
    if (words > 1) {
      e0 = lane_id < K ? m[(long long)lane_id * words + 1] : 0ull;
      e1 = lane_id + 32 < K ? m[(long long)(lane_id + 32) * words + 1] : 0ull;
    }
  }
  for (int cb = 0; cb < words && nk < out_size; ++cb) {
    if (warp_id == 0) {
      // prefetch the next block's diagonal words AND its rows' words of the column after it (the survivors among them
      // are OR-ed into removed[cb + 2] in the next iteration: fetched for all 64 rows now, the load no longer waits for
      // the sequential resolution that decides which of them count)
      unsigned long long n0 = 0ull, n1 = 0ull, f0 = 0ull, f1 = 0ull;
      if (cb + 1 < words) {
        const int r0 = (cb + 1) * 64 + lane_id, r1 = r0 + 32;
        if (r0 < K) n0 = m[(long long)r0 * words + cb + 1];
        if (r1 < K) n1 = m[(long long)r1 * words + cb + 1];
        if (cb + 2 < words) {
          if (r0 < K) f0 = m[(long long)r0 * words + cb + 2];
          if (r1 < K) f1 = m[(long long)r1 * words + cb + 2];
        }
      }
      unsigned long long rem = removed[cb], surv = 0ull;
      int k = nk;
      const int n_in = min(64, K - cb * 64);
      if (n_in < 64) rem |= ~0ull << n_in;  // positions past the list can never survive
      // jump from survivor to survivor (the lowest clear bit of `rem`): a block with a dozen survivors costs a dozen
      // shuffle steps, not 64 -- this loop is the serial spine of the whole scan
      while (k < out_size && rem != ~0ull) {
        const int j = __ffsll((long long)~rem) - 1;
        const unsigned long long dj = __shfl_sync(0xffffffffu, j < 32 ? d0 : d1, j & 31);
        surv |= (1ull << j);
        ++k;
        rem |= dj | ((2ull << j) - 1ull);  // its victims (later positions) and everything up to itself
      }
      // survivors' positions in the kept list, and their words of column cb+1
      unsigned long long next_word = 0ull;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane_id + 32 * h;
        if ((surv >> j) & 1ull) {
          kept[nk + __popcll(surv & ((1ull << j) - 1ull))] = cb * 64 + j;
          next_word |= (h == 0 ? e0 : e1);  // (prefetched one iteration ago; 0 beyond the last column)
        }
      }
      const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)(next_word & 0xffffffffull));
      const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(next_word >> 32));
      if (lane_id == 0) {
        s_surv2[cb & 1] = surv;
        const unsigned long long v = ((unsigned long long)hi << 32) | lo;
        if (v) atomicOr(&removed[cb + 1], v);
      }
      d0 = n0;
      d1 = n1;
      e0 = f0;
      e1 = f1;
    } else if (cb >= 1) {
      const int pb = cb - 1;  // block whose survivors' rows are applied to the columns >= pb + 2
      const unsigned long long surv = s_surv2[pb & 1];
      const int nw = words - pb - 2;
      if (surv && nw > 0) {
        const int total = 64 * nw;
        const int wt = tid - 32;
        for (int base = wt; base < total; base += 4 * nworkers) {
          unsigned long long v[4];
          int ww[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx = base + u * nworkers;
            v[u] = 0ull;
            ww[u] = 0;
            if (idx < total) {
              const int j = idx / nw;
              ww[u] = pb + 2 + (idx - j * nw);
              if ((surv >> j) & 1ull) v[u] = m[((long long)(pb * 64 + j)) * words + ww[u]];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (v[u]) atomicOr(&removed[ww[u]], v[u]);
        }
      }
    }
    __syncthreads();
    nk += __popcll(s_surv2[cb & 1]);
  }
  if (nms_keep_idx)
    for (int j = tid; j < keep_n; j += nthr) nms_keep_idx[(long long)img * keep_n + j] = j < nk ? kept[j] : -1;

  // _upsample_rois: drop paddings (score <= 0).  The kept list is in descending score order, so the
  // positive-score entries are a prefix and their count is its length.
  __shared__ int s_n;
  if (tid == 0) s_n = 0;
  __syncthreads();
  {
    int c = 0;
    for (int j = tid; j < nk; j += nthr) c += (sc[kept[j]] > 0.f) ? 1 : 0;
    if (c) atomicAdd(&s_n, c);
  }
  __syncthreads();
  const int n = s_n;
  float* out = rois + (long long)img * keep_n * 4;
  if (zero_pad) {  // bboxes_nms (utility/eval_helper.py:449-472): the selected boxes, zero padded to keep_n
    for (int j = tid; j < keep_n; j += nthr) {
      const bool on = j < n;
      reinterpret_cast<float4*>(out)[j] = on ? bx[kept[j]] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (roi_scores) roi_scores[(long long)img * keep_n + j] = on ? sc[kept[j]] : 0.f;
    }
  } else if (n == 0) {
    for (int j = tid; j < keep_n; j += nthr) {
      reinterpret_cast<float4*>(out)[j] = make_float4(0.2f, 0.2f, 0.8f, 0.8f);
      if (roi_scores) roi_scores[(long long)img * keep_n + j] = 1.f;
    }
  } else {
    const int left = keep_n - n;  // n <= keep_n always
    const int rem_cnt = left > 0 ? left % n : 0;
    if (rem_cnt > 0) {
      // tf.random_shuffle(range(n))[:rem_cnt] := first rem_cnt of the stable argsort of keys[:n]
      const float* keys = shuffle_keys ? shuffle_keys + (long long)img * keep_n : nullptr;
      for (int i = tid; i < n; i += nthr) {
        int rank = i;
        if (keys) {
          const float ki = keys[i];
          rank = 0;
          for (int j = 0; j < n; ++j) {
            const float kj = keys[j];
            rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
          }
        }
        if (rank < keep_n) shuf[rank] = i;
      }
    }
    __syncthreads();
    const int tiled = left > 0 ? n * (left / n + 1) : keep_n;
    for (int j = tid; j < keep_n; j += nthr) {
      const int src = (j < tiled) ? (j % n) : shuf[j - tiled];
      const int pos = kept[src];
      reinterpret_cast<float4*>(out)[j] = bx[pos];
      if (roi_scores) roi_scores[(long long)img * keep_n + j] = sc[pos];
    }
  }
  __syncthreads();
  if (rois_yxhw) {  // _point2center
    for (int j = tid; j < keep_n; j += nthr) {
      const float4 b = reinterpret_cast<const float4*>(out)[j];
      const float h = __fsub_rn(b.z, b.x), w = __fsub_rn(b.w, b.y);
      reinterpret_cast<float4*>(rois_yxhw + (long long)img * keep_n * 4)[j] =
          make_float4(__fadd_rn(b.x, __fmul_rn(h, 0.5f)), __fadd_rn(b.y, __fmul_rn(w, 0.5f)), h, w);
    }
  }
}

// ---- head post-processing ------------------------------------------------------------------------
// softmax over the class scores and ext_decode_rois (preprocessing/anchor_manipulator.py:671-683,
// light_head_rfcn_eval.py:406-410): one thread per RoI.
__global__ void __launch_bounds__(256) head_decode_kernel(const float* __restrict__ rois,
                                                          const float* __restrict__ head_out, int ch_stride,
                                                          int cls_off, int num_classes, int loc_off,
                                                          float* __restrict__ probs, float* __restrict__ boxes,
                                                          long long* __restrict__ classes /* or NULL */,
                                                          float* __restrict__ best_prob /* or NULL */, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float* h = head_out + i * ch_stride;
  float mx = -FLT_MAX;
  for (int c = 0; c < num_classes; ++c) mx = fmaxf(mx, __ldg(h + cls_off + c));
  float sum = 0.f;
  for (int c = 0; c < num_classes; ++c) sum = __fadd_rn(sum, expf(__fsub_rn(__ldg(h + cls_off + c), mx)));
  float pbest = -1.f;
  int cbest = 0;
  for (int c = 0; c < num_classes; ++c) {
    const float pc = __fdiv_rn(expf(__fsub_rn(__ldg(h + cls_off + c), mx)), sum);
    probs[i * num_classes + c] = pc;
    if (pc > pbest) {  // tf.argmax / tf.reduce_max over the class axis (light_head_rfcn_eval.py:413-416): first maximum
      pbest = pc;
      cbest = c;
    }
  }
  if (classes) classes[i] = cbest;
  if (best_prob) best_prob[i] = pbest;
  const float4 r = reinterpret_cast<const float4*>(rois)[i];
  const float href = __fsub_rn(r.z, r.x), wref = __fsub_rn(r.w, r.y);
  const float yref = __fadd_rn(r.x, __fmul_rn(href, 0.5f)), xref = __fadd_rn(r.y, __fmul_rn(wref, 0.5f));
  const float t0 = __ldg(h + loc_off), t1 = __ldg(h + loc_off + 1), t2 = __ldg(h + loc_off + 2),
              t3 = __ldg(h + loc_off + 3);
  const float ph = __fmul_rn(expf(t2), href), pw = __fmul_rn(expf(t3), wref);
  const float cy = __fadd_rn(__fmul_rn(t0, href), yref), cx = __fadd_rn(__fmul_rn(t1, wref), xref);
  reinterpret_cast<float4*>(boxes)[i] = make_float4(__fsub_rn(cy, __fmul_rn(ph, 0.5f)), __fsub_rn(cx, __fmul_rn(pw, 0.5f)),
                                                    __fadd_rn(cy, __fmul_rn(ph, 0.5f)), __fadd_rn(cx, __fmul_rn(pw, 0.5f)));
}

int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace
}  // namespace xdet

using namespace xdet;

extern "C" int xdet_rpn_decode(const float* d_rpn_out, int ch_stride, int cls_off, int box_off, const float* d_yref,
                               const float* d_xref, const float* d_href, const float* d_wref, int N, int fh, int fw,
                               int A, float* d_scores, float* d_boxes, void* stream) {
  if (N <= 0 || fh <= 0 || fw <= 0 || A <= 0) return fail(XDET_EINVAL, "rpn_decode: non-positive dimension");
  const long long total = (long long)N * fh * fw * A;
  rpn_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_rpn_out, ch_stride, cls_off, box_off, d_yref, d_xref, d_href, d_wref, fh * fw, A, d_scores, d_boxes, total);
  return after_launch("rpn_decode_kernel");
}

extern "C" size_t xdet_rpn_select_workspace_bytes(int N, int A_tot, int pre_nms_top_n) {
  const size_t words = (size_t)(pre_nms_top_n + 63) / 64;
  return (size_t)N * A_tot * 8 + (size_t)N * pre_nms_top_n * (4 + 16) + (size_t)N * pre_nms_top_n * words * 8 + 256;
}

extern "C" int xdet_rpn_select(const float* d_scores, const float* d_boxes, int N, int A_tot, int pre_nms_top_n,
                               int post_nms_top_n, float nms_threshold, float min_size, const float* d_shuffle_keys,
                               float* d_rois, float* d_rois_yxhw, float* d_roi_scores, int* d_nms_keep_idx,
                               void* d_workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0 || A_tot <= 0 || pre_nms_top_n <= 0 || post_nms_top_n <= 0)
    return fail(XDET_EINVAL, "rpn_select: non-positive dimension");
  const int K = pre_nms_top_n, keep = post_nms_top_n;
  if (workspace_bytes < xdet_rpn_select_workspace_bytes(N, A_tot, K))
    return fail(XDET_EINVAL, "rpn_select: workspace too small (%zu < %zu)", workspace_bytes,
                xdet_rpn_select_workspace_bytes(N, A_tot, K));
  if ((reinterpret_cast<uintptr_t>(d_workspace) & 15) != 0) return fail(XDET_EINVAL, "rpn_select: workspace alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int words = (K + 63) / 64;
  unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);
  ws += (size_t)N * A_tot * 8;
  float* top_boxes = reinterpret_cast<float*>(ws);
  ws += (size_t)N * K * 16;
  float* top_scores = reinterpret_cast<float*>(ws);
  ws += (size_t)N * K * 4;
  ws += (16 - (reinterpret_cast<uintptr_t>(ws) & 15)) & 15;
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws);

  const int P = next_pow2(K);
  const size_t smem_topk = (size_t)P * 8;
  if (smem_topk > 200 * 1024) return fail(XDET_EINVAL, "rpn_select: pre_nms_top_n %d too large for the in-CTA sort", K);
  XDET_TRY(check_cuda(cudaFuncSetAttribute(rpn_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_topk),
                      "cudaFuncSetAttribute(rpn_topk)"));
  rpn_topk_kernel<<<N, kSelThreads, smem_topk, st>>>(d_scores, d_boxes, A_tot, K, P, min_size, 0, keys, top_scores,
                                                     top_boxes);
  XDET_TRY(after_launch("rpn_topk_kernel"));
  // (the lower triangle of the matrix is never written and never read)
  nms_mask_kernel<<<dim3(words, words, N), 64, 0, st>>>(top_boxes, K, words, nms_threshold, mask);
  XDET_TRY(after_launch("nms_mask_kernel"));
  const size_t smem_scan = (size_t)words * 8 + (size_t)keep * 8;
  XDET_TRY(check_cuda(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan),
                      "cudaFuncSetAttribute(nms_scan)"));
  nms_scan_kernel<<<N, kSelThreads, smem_scan, st>>>(top_scores, top_boxes, mask, K, words, keep, d_shuffle_keys, 0,
                                                     d_rois, d_rois_yxhw, d_roi_scores, d_nms_keep_idx);
  return after_launch("nms_scan_kernel");
}

// ---- detection post-processing (light_head_rfcn_eval.py:263-290, utility/eval_helper.py) ------------------------
namespace xdet {
namespace {
// One thread per (image, class >= 1, RoI): tf_bboxes_select_layer (eval_helper.py:556-588: score and box times the
// 0/1 mask of score > select_threshold), bboxes_clip against bbox_img (:365-404), filter_boxes (:278-317) and
// bboxes_resize (:423-447).  Candidates that are dropped become (score 0, box 0): exactly what the reference's
// zero padding looks like to bboxes_sort / bboxes_nms.
__global__ void __launch_bounds__(256) det_prepare_kernel(const float* __restrict__ probs,
                                                          const float* __restrict__ boxes,
                                                          const float* __restrict__ bbox_img,
                                                          const float* __restrict__ min_size, int R, int num_classes,
                                                          float select_threshold, float* __restrict__ cand_scores,
                                                          float* __restrict__ cand_boxes, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int r = (int)(e % R);
  const int c1 = (int)((e / R) % (num_classes - 1));
  const int n = (int)(e / ((long long)R * (num_classes - 1)));
  const float s0 = probs[((long long)n * R + r) * num_classes + c1 + 1];
  const float fmask = s0 > select_threshold ? 1.f : 0.f;
  const float s = __fmul_rn(s0, fmask);
  const float4 b = reinterpret_cast<const float4*>(boxes)[(long long)n * R + r];
  const float4 ref = reinterpret_cast<const float4*>(bbox_img)[n];
  float ymin = fmaxf(__fmul_rn(b.x, fmask), ref.x), xmin = fmaxf(__fmul_rn(b.y, fmask), ref.y);
  const float ymax = fminf(__fmul_rn(b.z, fmask), ref.z), xmax = fminf(__fmul_rn(b.w, fmask), ref.w);
  ymin = fminf(ymin, ymax);
  xmin = fminf(xmin, xmax);
  const float ms = min_size[n];
  const float ws = __fsub_rn(xmax, xmin), hs = __fsub_rn(ymax, ymin);
  const float xc = __fadd_rn(xmin, __fmul_rn(ws, 0.5f)), yc = __fadd_rn(ymin, __fmul_rn(hs, 0.5f));
  const bool keep = ws > ms && hs > ms && xc > 0.f && yc > 0.f && xc < 1.f && yc < 1.f;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  float so = 0.f;
  if (keep && s > 0.f) {
    const float sy = __fsub_rn(ref.z, ref.x), sx = __fsub_rn(ref.w, ref.y);
    o = make_float4(__fdiv_rn(__fsub_rn(ymin, ref.x), sy), __fdiv_rn(__fsub_rn(xmin, ref.y), sx),
                    __fdiv_rn(__fsub_rn(ymax, ref.x), sy), __fdiv_rn(__fsub_rn(xmax, ref.y), sx));
    so = s;
  }
  cand_scores[e] = so;
  reinterpret_cast<float4*>(cand_boxes)[e] = o;
}
}  // namespace
}  // namespace xdet

extern "C" size_t xdet_det_postprocess_workspace_bytes(int N, int R, int num_classes, int top_k) {
  const size_t M = (size_t)N * (num_classes - 1);
  return M * R * (4 + 16) + xdet_rpn_select_workspace_bytes((int)M, R, top_k) + 256;
}

extern "C" int xdet_det_postprocess(const float* d_probs, const float* d_boxes, const float* d_bbox_img,
                                    const float* d_min_size, int N, int R, int num_classes, float select_threshold,
                                    int top_k, int keep_top_k, float nms_threshold, float* d_out_scores,
                                    float* d_out_boxes, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0 || R <= 0 || num_classes < 2 || top_k <= 0 || keep_top_k <= 0)
    return fail(XDET_EINVAL, "det_postprocess: non-positive dimension");
  if (workspace_bytes < xdet_det_postprocess_workspace_bytes(N, R, num_classes, top_k))
    return fail(XDET_EINVAL, "det_postprocess: workspace too small");
  if ((reinterpret_cast<uintptr_t>(d_workspace) & 15) != 0) return fail(XDET_EINVAL, "det_postprocess: workspace alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int M = N * (num_classes - 1);
  // the sorted list is min(#candidates, top_k) long in the reference; entries past the candidates are zero padding
  const int K = top_k < R ? top_k : R;
  const int keep = keep_top_k;
  unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
  float* cand_boxes = reinterpret_cast<float*>(ws);
  ws += (size_t)M * R * 16;
  float* cand_scores = reinterpret_cast<float*>(ws);
  ws += (size_t)M * R * 4;
  ws += (16 - (reinterpret_cast<uintptr_t>(ws) & 15)) & 15;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);
  ws += (size_t)M * R * 8;
  float* top_boxes = reinterpret_cast<float*>(ws);
  ws += (size_t)M * K * 16;
  float* top_scores = reinterpret_cast<float*>(ws);
  ws += (size_t)M * K * 4;
  ws += (16 - (reinterpret_cast<uintptr_t>(ws) & 15)) & 15;
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws);

  const long long total = (long long)M * R;
  det_prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_probs, d_boxes, d_bbox_img, d_min_size, R,
                                                                      num_classes, select_threshold, cand_scores,
                                                                      cand_boxes, total);
  XDET_TRY(after_launch("det_prepare_kernel"));
  const int words = (K + 63) / 64;
  const int P = R <= 2048 ? next_pow2(R) : next_pow2(K);  // few candidates per class: sort them all (no radix select)
  const size_t smem_topk = (size_t)P * 8;
  if (smem_topk > 200 * 1024) return fail(XDET_EINVAL, "det_postprocess: top_k %d too large for the in-CTA sort", K);
  XDET_TRY(check_cuda(cudaFuncSetAttribute(rpn_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_topk),
                      "cudaFuncSetAttribute(rpn_topk)"));
  // many small selections (one per image and class): 256-thread CTAs, several per SM, one wave
  rpn_topk_kernel<<<M, kDetThreads, smem_topk, st>>>(cand_scores, cand_boxes, R, K, P, 0.f, 1, keys, top_scores,
                                                     top_boxes);
  XDET_TRY(after_launch("rpn_topk_kernel"));
  nms_mask_kernel<<<dim3(words, words, M), 64, 0, st>>>(top_boxes, K, words, nms_threshold, mask);
  XDET_TRY(after_launch("nms_mask_kernel"));
  const size_t smem_scan = (size_t)words * 8 + (size_t)keep * 8;
  XDET_TRY(check_cuda(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan),
                      "cudaFuncSetAttribute(nms_scan)"));
  nms_scan_kernel<<<M, kDetThreads, smem_scan, st>>>(top_scores, top_boxes, mask, K, words, keep, nullptr, 1,
                                                     d_out_boxes, nullptr, d_out_scores, nullptr);
  return after_launch("nms_scan_kernel");
}

extern "C" int xdet_head_decode_ex(const float* d_rois, const float* d_head_out, int ch_stride, int cls_off,
                                   int num_classes, int loc_off, long long M, float* d_probs, float* d_boxes,
                                   long long* d_classes, float* d_best_prob, void* stream) {
  if (M <= 0) return XDET_OK;
  if (num_classes <= 0) return fail(XDET_EINVAL, "head_decode: num_classes must be positive");
  head_decode_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_rois, d_head_out, ch_stride, cls_off, num_classes, loc_off, d_probs, d_boxes, d_classes, d_best_prob, M);
  return after_launch("head_decode_kernel");
}

extern "C" int xdet_head_decode(const float* d_rois, const float* d_head_out, int ch_stride, int cls_off, int num_classes,
                                int loc_off, long long M, float* d_probs, float* d_boxes, void* stream) {
  return xdet_head_decode_ex(d_rois, d_head_out, ch_stride, cls_off, num_classes, loc_off, M, d_probs, d_boxes, nullptr,
                             nullptr, stream);
}

// ---- TP / FP matching of detections against ground truth (utility/eval_helper.py:671-788) -------------------
namespace xdet {
namespace {
// One warp per (image, class): the detections are visited in order (they are sorted by score), each one is matched
// to the ground-truth box of its class with the largest Jaccard score (first maximum, as tf.argmax), exactly as
// bboxes_matching's tf.while_loop does; the lanes share the ground-truth boxes.
__global__ void __launch_bounds__(32) det_match_kernel(const float* __restrict__ det_boxes,
                                                       const int* __restrict__ glabels,
                                                       const float* __restrict__ gbboxes,
                                                       const int* __restrict__ gdifficult, int num_fg, int Kd, int G,
                                                       float thr, unsigned char* __restrict__ tp,
                                                       unsigned char* __restrict__ fp, int* __restrict__ n_gb) {
  extern __shared__ unsigned char gmatch[];  // [G]
  const int n = blockIdx.x / num_fg, c1 = blockIdx.x % num_fg, label = c1 + 1;
  const int lane = threadIdx.x;
  const int* gl = glabels + (long long)n * G;
  const int* gd = gdifficult + (long long)n * G;
  const float4* gb = reinterpret_cast<const float4*>(gbboxes) + (long long)n * G;
  int cnt = 0;
  for (int g = lane; g < G; g += 32) {
    gmatch[g] = 0;
    cnt += (gl[g] == label && gd[g] == 0) ? 1 : 0;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) n_gb[blockIdx.x] = cnt;
  __syncwarp();
  const float4* db = reinterpret_cast<const float4*>(det_boxes) + (long long)blockIdx.x * Kd;
  for (int i = 0; i < Kd; ++i) {
    const float4 r = db[i];
    const float rarea = __fmul_rn(__fsub_rn(r.z, r.x), __fsub_rn(r.w, r.y));
    float best = -1.f;  // Jaccard scores are >= 0, so the first candidate always replaces this
    int best_g = G;
    for (int g = lane; g < G; g += 32) {
      const float4 b = gb[g];
      const float h = fmaxf(__fsub_rn(fminf(b.z, r.z), fmaxf(b.x, r.x)), 0.f);
      const float w = fmaxf(__fsub_rn(fminf(b.w, r.w), fmaxf(b.y, r.y)), 0.f);
      const float inter = __fmul_rn(h, w);
      const float garea = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
      const float uni = __fadd_rn(__fadd_rn(-inter, garea), rarea);
      float j = uni > 0.f ? __fdiv_rn(inter, uni) : 0.f;
      j = __fmul_rn(j, gl[g] == label ? 1.f : 0.f);
      if (j > best) {  // strict: the first maximum wins inside a lane (g ascending)
        best = j;
        best_g = g;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {  // across lanes: larger score, ties -> smaller index
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int og = __shfl_xor_sync(0xffffffffu, best_g, o);
      if (ob > best || (ob == best && og < best_g)) {
        best = ob;
        best_g = og;
      }
    }
    bool t = false, f = true;  // no ground truth at all: a false positive
    if (best_g < G) {
      const bool match = best > thr;
      const bool existing = gmatch[best_g] != 0;
      const bool nd = gd[best_g] == 0;
      t = nd && match && !existing;
      f = nd && (existing || !match);
      __syncwarp();
      if (lane == 0 && nd && match) gmatch[best_g] = 1;
      __syncwarp();
    }
    if (lane == 0) {
      tp[(long long)blockIdx.x * Kd + i] = t ? 1 : 0;
      fp[(long long)blockIdx.x * Kd + i] = f ? 1 : 0;
    }
  }
}
}  // namespace
}  // namespace xdet

extern "C" int xdet_det_match(const float* d_det_boxes, const int* d_glabels, const float* d_gbboxes,
                              const int* d_gdifficult, int N, int num_classes, int Kd, int G,
                              float matching_threshold, unsigned char* d_tp, unsigned char* d_fp, int* d_n_gbboxes,
                              void* stream) {
  if (N <= 0 || num_classes < 2 || Kd <= 0 || G < 0) return fail(XDET_EINVAL, "det_match: bad dimension");
  if (G > 48 * 1024) return fail(XDET_EINVAL, "det_match: too many ground-truth boxes per image");
  const int M = N * (num_classes - 1);
  det_match_kernel<<<M, 32, (size_t)(G > 0 ? G : 1), (cudaStream_t)stream>>>(d_det_boxes, d_glabels, d_gbboxes,
                                                                           d_gdifficult, num_classes - 1, Kd, G,
                                                                           matching_threshold, d_tp, d_fp, d_n_gbboxes);
  return after_launch("det_match_kernel");
}
