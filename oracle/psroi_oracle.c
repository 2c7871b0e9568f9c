/* TEST INFRASTRUCTURE ONLY (oracle). Never imported, linked or executed by the
 * product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Plain-C restatement of the reference's PsRoiAlign CPU arithmetic:
 *   forward   /root/reference/cpp/PSROIPooling/ps_roi_align_op.cc:94-193
 *   backward  /root/reference/cpp/PSROIPooling/ps_roi_align_grad_op.cc:200-315
 * Parity pin: validated bit-for-bit against the reference's own unmodified C++
 * (oracle/_ref/libref_psroi.so, built by oracle/Makefile) in tests/test_oracle.py and
 * against the known-answer vectors of the reference fixture (cpp/PSROIPooling/test_op.py:52-81)
 * committed in tests/golden/.
 *
 * The loop nest is organised differently from the reference (geometry once per
 * (image, roi, bin) and shared by the bank channels; backward as an ordered
 * scatter per plane) but every floating-point operation that reaches an output
 * is the same operation, on the same operands, in the same order:
 *   - RoI/bin geometry in fp32, one rounding per operation (build with
 *     -ffp-contract=off so that no mul+add is fused);
 *   - the 4-tap blend: the first three products contain a `1.` double literal and are
 *     therefore formed in fp64 as ((wa*wb)*P); the FOURTH product, `fx * fy * P`
 *     (ps_roi_align_op.cc:176), has only float operands and is formed in fp32
 *     (two fp32 roundings) before being widened; the four terms are summed left to
 *     right in fp64 and rounded once to fp32 (:173-176).  The same holds for the four
 *     gradient taps (ps_roi_align_grad_op.cc:271-282).  Found by pinning against the
 *     compiled reference: an all-fp64 blend differs from it in ~25 % of outputs by 1 ulp.
 *   - max: strict '<' from -FLT_MAX, first maximum wins (:177-181);
 *     mean: fp32 running sum in sample order, one fp32 divide (:183,188).
 * Deliberate, documented deviations (both outside the reference's input contract,
 * ps_roi_align_op.cc:50-53):
 *   - degenerate RoIs (h or w < FLT_MIN): the reference leaves pooled_index
 *     unwritten (:123-126); this oracle writes 0.
 *   - a sample whose integer row/col reaches H/W through fp rounding is clamped to
 *     H-1/W-1 (the reference would read one plane row past the end).
 */
#include <float.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int degenerate;
  float ymin, xmin;       /* clipped RoI origin, feature-map pixels */
  float bin_h, bin_w;     /* bin extent */
  float step_h, step_w;   /* sample pitch inside a bin */
  int nh, nw;             /* samples per bin along h, w */
} roi_geom_t;

static inline float fmaxf_(float a, float b) { return a < b ? b : a; } /* std::max(a,b) */
static inline float fminf_(float a, float b) { return b < a ? b : a; } /* std::min(a,b) */
static inline int imin_(int a, int b) { return b < a ? b : a; }

/* ps_roi_align_op.cc:123-158 */
static void roi_geometry(const float* roi, int H, int W, int gw, int gh, roi_geom_t* g) {
  if (roi[2] < FLT_MIN || roi[3] < FLT_MIN) {
    g->degenerate = 1;
    return;
  }
  g->degenerate = 0;
  float cy = roi[0] * (float)H;
  float cx = roi[1] * (float)W;
  float rh = fmaxf_(roi[2] * (float)H, 1.0f);
  float rw = fmaxf_(roi[3] * (float)W, 1.0f);
  /* (float)(rh / 2.) : the double quotient of a float by 2 is exact */
  float half_h = (float)((double)rh / 2.);
  float half_w = (float)((double)rw / 2.);
  float ymin = fmaxf_(cy - half_h, 0.0f);
  float xmin = fmaxf_(cx - half_w, 0.0f);
  float ymax = fminf_(cy + half_h, (float)H - FLT_MIN);
  float xmax = fminf_(cx + half_w, (float)W - FLT_MIN);
  float ext_h = ymax - ymin;
  float ext_w = xmax - xmin;
  g->ymin = ymin;
  g->xmin = xmin;
  g->bin_w = ext_w / (float)gw;
  g->bin_h = ext_h / (float)gh;
  g->nw = (int)g->bin_w + 1;
  g->nh = (int)g->bin_h + 1;
  g->step_w = g->bin_w / (float)g->nw;
  g->step_h = g->bin_h / (float)g->nh;
}

/* sample coordinate: float(start + step*i + step/2.)  (ps_roi_align_op.cc:163-164).
 * `start + step*i` is an fp32 mul then an fp32 add; the last addend is a double. */
static inline float sample_coord(float start, float step, int i) {
  float a = step * (float)i;
  float b = start + a;
  return (float)((double)b + (double)step / 2.);
}

typedef struct {
  const float* in;
  const float* rois;
  float* out;
  int32_t* idx;
  int N, C, H, W, R, gw, gh, use_max;
  long job_lo, job_hi; /* range of (image*R + roi) */
} fwd_job_t;

static void* fwd_worker(void* p) {
  fwd_job_t* j = (fwd_job_t*)p;
  const int G = j->gw * j->gh, bank = j->C / G, H = j->H, W = j->W;
  for (long pr = j->job_lo; pr < j->job_hi; ++pr) {
    const int img = (int)(pr / j->R);
    roi_geom_t g;
    roi_geometry(j->rois + pr * 4, H, W, j->gw, j->gh, &g);
    float* o = j->out + pr * (long)j->C;
    int32_t* oi = j->idx + pr * (long)j->C;
    if (g.degenerate) {
      for (int c = 0; c < G * bank; ++c) { o[c] = 0.0f; oi[c] = 0; }
      continue;
    }
    for (int bin = 0; bin < G; ++bin) {
      const int row = bin / j->gw, col = bin % j->gw;
      const float x0 = g.xmin + g.bin_w * (float)col;
      const float y0 = g.ymin + g.bin_h * (float)row;
      for (int ch = 0; ch < bank; ++ch) {
        const float* plane = j->in + ((long)img * j->C + (long)bin * bank + ch) * H * W;
        float best = j->use_max ? -FLT_MAX : 0.0f;
        int best_i = 0;
        for (int hi = 0; hi < g.nh; ++hi) {
          for (int wi = 0; wi < g.nw; ++wi) {
            float x = sample_coord(x0, g.step_w, wi);
            float y = sample_coord(y0, g.step_h, hi);
            int ix = (int)x, iy = (int)y;
            float fx = x - (float)ix, fy = y - (float)iy;
            int ixc = imin_(ix, W - 1), iyc = imin_(iy, H - 1); /* deviation: clamp */
            int ix1 = imin_(ix + 1, W - 1), iy1 = imin_(iy + 1, H - 1);
            double ax = 1. - (double)fx, ay = 1. - (double)fy;
            double v = ax * ay * (double)plane[iyc * W + ixc] +
                       ax * (double)fy * (double)plane[iy1 * W + ixc] +
                       (double)fx * ay * (double)plane[iyc * W + ix1] +
                       (double)(fx * fy * plane[iy1 * W + ix1]); /* all-float term, see header */
            float vf = (float)v;
            if (j->use_max) {
              if (best < vf) { best = vf; best_i = g.nw * hi + wi; }
            } else {
              best += vf;
            }
          }
        }
        if (!j->use_max) best /= (float)(g.nh * g.nw);
        o[bin * bank + ch] = best;
        oi[bin * bank + ch] = j->use_max ? best_i : 0;
      }
    }
  }
  return NULL;
}

static int run_jobs(void* (*fn)(void*), void* jobs, size_t job_size, int n) {
  if (n == 1) { fn(jobs); return 0; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n);
  if (!th) return -1;
  for (int t = 0; t < n; ++t) pthread_create(&th[t], NULL, fn, (char*)jobs + job_size * (size_t)t);
  for (int t = 0; t < n; ++t) pthread_join(th[t], NULL);
  free(th);
  return 0;
}

/* inputs[N,C,H,W], rois[N,R,4] (cy,cx,h,w in [0,1]) -> pooled[N,R,G,C/G], index[N,R,G,C/G]. */
int oracle_psroi_align_fwd(const float* inputs, const float* rois, float* pooled, int32_t* index,
                           int N, int C, int H, int W, int R, int gw, int gh, int use_max, int threads) {
  if (gw <= 0 || gh <= 0 || C % (gw * gh) != 0) return -1;
  long total = (long)N * R;
  if (threads < 1) threads = 1;
  if (threads > total) threads = total > 0 ? (int)total : 1;
  fwd_job_t* jobs = (fwd_job_t*)malloc(sizeof(fwd_job_t) * (size_t)threads);
  if (!jobs) return -2;
  long per = (total + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    fwd_job_t jb = {inputs, rois, pooled, index, N, C, H, W, R, gw, gh, use_max, per * t,
                    per * (t + 1) > total ? total : per * (t + 1)};
    if (jb.job_lo > total) jb.job_lo = total;
    jobs[t] = jb;
  }
  int rc = run_jobs(fwd_worker, jobs, sizeof(fwd_job_t), threads);
  free(jobs);
  return rc;
}

typedef struct {
  const float* rois;
  const float* gout;
  const int32_t* idx;
  float* gin;
  int N, C, H, W, R, gw, gh, use_max;
  long job_lo, job_hi; /* range of planes (image*C + c) */
} bwd_job_t;

/* One input plane (image, c) at a time.  A cell of that plane receives, in the
 * reference (ps_roi_align_grad_op.cc:212-311), for roi = 0..R-1 in order:
 *   max : up to four `+= (float)(w_tap * g)` in the tap order 00,(+1 row),(+1 col),(+1,+1),
 *         for the arg-max sample only (:259-286);
 *   mean: an fp32 partial `acc` over samples (h-major) and taps in that order, then
 *         `+= acc / (float)(nh*nw)` (:287-309).
 * Visiting rois, samples and taps in that same order and adding into the plane
 * reproduces each cell's addition sequence exactly. */
static void* bwd_worker(void* p) {
  bwd_job_t* j = (bwd_job_t*)p;
  const int G = j->gw * j->gh, bank = j->C / G, H = j->H, W = j->W;
  float* acc = (float*)calloc((size_t)H * W, sizeof(float));
  unsigned char* touched = (unsigned char*)calloc((size_t)H * W, 1);
  for (long pl = j->job_lo; pl < j->job_hi; ++pl) {
    const int img = (int)(pl / j->C), c = (int)(pl % j->C);
    const int bin = c / bank, row = bin / j->gw, col = bin % j->gw;
    float* plane = j->gin + pl * (long)H * W;
    memset(plane, 0, sizeof(float) * (size_t)H * W);
    for (int r = 0; r < j->R; ++r) {
      roi_geom_t g;
      roi_geometry(j->rois + ((long)img * j->R + r) * 4, H, W, j->gw, j->gh, &g);
      if (g.degenerate) continue;
      const long o = ((long)img * j->R + r) * j->C + c;
      const float gv = j->gout[o];
      const float x0 = g.xmin + g.bin_w * (float)col;
      const float y0 = g.ymin + g.bin_h * (float)row;
      const int s_lo = j->use_max ? j->idx[o] : 0;
      const int s_hi = j->use_max ? s_lo + 1 : g.nh * g.nw;
      int y_lo = H, y_hi = -1, x_lo = W, x_hi = -1;
      for (int s = s_lo; s < s_hi; ++s) {
        const int hi = s / g.nw, wi = s % g.nw;
        float x = sample_coord(x0, g.step_w, wi);
        float y = sample_coord(y0, g.step_h, hi);
        int ix = (int)x, iy = (int)y;
        float fx = x - (float)ix, fy = y - (float)iy;
        int ix1 = imin_(ix + 1, W - 1), iy1 = imin_(iy + 1, H - 1);
        double ax = 1. - (double)fx, ay = 1. - (double)fy;
        const float t[4] = {(float)(ax * ay * (double)gv), (float)(ax * (double)fy * (double)gv),
                            (float)((double)fx * ay * (double)gv), fx * fy * gv /* all-float */};
        const int ty[4] = {iy, iy1, iy, iy1}, tx[4] = {ix, ix, ix1, ix1};
        float* dst = j->use_max ? plane : acc;
        for (int k = 0; k < 4; ++k) {
          /* a tap lands only if a map cell equals its (row, col) (:271-282) */
          if (ty[k] < 0 || ty[k] >= H || tx[k] < 0 || tx[k] >= W) continue;
          dst[ty[k] * W + tx[k]] += t[k];
          if (!j->use_max) {
            touched[ty[k] * W + tx[k]] = 1;
            if (ty[k] < y_lo) y_lo = ty[k];
            if (ty[k] > y_hi) y_hi = ty[k];
            if (tx[k] < x_lo) x_lo = tx[k];
            if (tx[k] > x_hi) x_hi = tx[k];
          }
        }
      }
      if (!j->use_max) {
        const float cnt = (float)(g.nw * g.nh);
        for (int yy = y_lo; yy <= y_hi; ++yy)
          for (int xx = x_lo; xx <= x_hi; ++xx)
            if (touched[yy * W + xx]) {
              plane[yy * W + xx] += acc[yy * W + xx] / cnt;
              acc[yy * W + xx] = 0.0f;
              touched[yy * W + xx] = 0;
            }
      }
    }
  }
  free(acc);
  free(touched);
  return NULL;
}

/* rois[N,R,4], pooled_grad[N,R,G,C/G], index -> grad[N,C,H,W] (fully overwritten). */
int oracle_psroi_align_bwd(const float* rois, const float* pooled_grad, const int32_t* index, float* grad,
                           int N, int C, int H, int W, int R, int gw, int gh, int use_max, int threads) {
  if (gw <= 0 || gh <= 0 || C % (gw * gh) != 0) return -1;
  long total = (long)N * C;
  if (threads < 1) threads = 1;
  if (threads > total) threads = total > 0 ? (int)total : 1;
  bwd_job_t* jobs = (bwd_job_t*)malloc(sizeof(bwd_job_t) * (size_t)threads);
  if (!jobs) return -2;
  long per = (total + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    bwd_job_t jb = {rois, pooled_grad, index, grad, N, C, H, W, R, gw, gh, use_max, per * t,
                    per * (t + 1) > total ? total : per * (t + 1)};
    if (jb.job_lo > total) jb.job_lo = total;
    jobs[t] = jb;
  }
  int rc = run_jobs(bwd_worker, jobs, sizeof(bwd_job_t), threads);
  free(jobs);
  return rc;
}
