// CPU EMULATION of the staged CUDA kernel x-detector_b200/csrc/depthwise_wgrad.cu (test infrastructure):
// the kernel's own source is compiled as host C++ against the few stand-ins below -- thread/block indices as
// thread-locals, __shared__ as block-wide static storage (one block runs at a time), __syncthreads as a pthread
// barrier over the block's 256 real threads, atomicAdd under a mutex, bf16 types with round-to-nearest-even
// conversion -- and driven over the launcher's own grid decomposition (depthwise_wgrad_grid).  It checks what can
// be wrong in such a kernel short of the hardware: the index arithmetic, the borders, the dilation, the channel
// tail, the slab split, the shared-memory fold and the accumulation into dW.  Built and run by
// tests/test_staged_emulation.py with g++.
#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define XDET_EMULATE_ON_CPU 1
struct dim3 {
  unsigned x = 1, y = 1, z = 1;
};
static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static pthread_barrier_t g_block_barrier;
static std::mutex g_atomic_mutex;
#define __global__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
static inline void __syncthreads() { pthread_barrier_wait(&g_block_barrier); }
static inline float atomicAdd(float* p, float v) {
  std::lock_guard<std::mutex> lock(g_atomic_mutex);
  const float old = *p;
  *p = old + v;
  return old;
}
struct __nv_bfloat16 {
  uint16_t bits;
};
struct __nv_bfloat162 {
  __nv_bfloat16 x, y;
};
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float bf16_to_float(__nv_bfloat16 h) {
  uint32_t u = (uint32_t)h.bits << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline __nv_bfloat16 float_to_bf16(float f) {  // round to nearest even (finite inputs)
  uint32_t u;
  std::memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return __nv_bfloat16{(uint16_t)(u >> 16)};
}
static inline float2 __bfloat1622float2(__nv_bfloat162 v) { return float2{bf16_to_float(v.x), bf16_to_float(v.y)}; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }

#include "../../x-detector_b200/csrc/depthwise_wgrad.cu"

using xdet::depthwise3x3_wgrad_kernel;

template <int DIL>
static void launch(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int N, int H, int W, int C, int relu_in,
                   int num_sms) {
  int chunks, slabs, per;
  depthwise_wgrad_grid((long long)N * H * W, C, num_sms, &chunks, &slabs, &per);
  gridDim.x = chunks;
  gridDim.y = slabs;
  blockDim.x = xdet::kWgThreads;
  pthread_barrier_init(&g_block_barrier, nullptr, blockDim.x);
  for (int by = 0; by < slabs; ++by)
    for (int bx = 0; bx < chunks; ++bx) {
      std::vector<std::thread> threads;
      for (unsigned t = 0; t < blockDim.x; ++t)
        threads.emplace_back([=] {
          threadIdx.x = t;
          blockIdx.x = bx;
          blockIdx.y = by;
          depthwise3x3_wgrad_kernel<DIL>(x, dy, dw, N, H, W, C, relu_in, per);
        });
      for (auto& th : threads) th.join();
    }
  pthread_barrier_destroy(&g_block_barrier);
  std::printf("    grid (%d chunks x %d slabs of %d pixels)\n", chunks, slabs, per);
}

static int run_case(int N, int H, int W, int C, int dil, int relu_in, int num_sms, unsigned seed) {
  const size_t n = (size_t)N * H * W * C;
  std::vector<__nv_bfloat16> x(n), dy(n);
  std::srand(seed);
  auto rnd = [] { return (float)std::rand() / RAND_MAX * 2.f - 1.f; };
  for (size_t i = 0; i < n; ++i) {
    x[i] = float_to_bf16(rnd());
    dy[i] = float_to_bf16(rnd());
  }
  std::vector<float> dw(9 * (size_t)C, 0.5f);  // the kernel ACCUMULATES: start from a non-zero buffer
  if (dil == 1)
    launch<1>(x.data(), dy.data(), dw.data(), N, H, W, C, relu_in, num_sms);
  else
    launch<2>(x.data(), dy.data(), dw.data(), N, H, W, C, relu_in, num_sms);
  double worst = 0.0, scale = 0.0;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw)
      for (int c = 0; c < C; ++c) {
        double s = 0.0;
        for (int nn = 0; nn < N; ++nn)
          for (int y = 0; y < H; ++y)
            for (int xx = 0; xx < W; ++xx) {
              const int yi = y + (kh - 1) * dil, xi = xx + (kw - 1) * dil;
              if (yi < 0 || yi >= H || xi < 0 || xi >= W) continue;
              double a = bf16_to_float(x[(((size_t)nn * H + yi) * W + xi) * C + c]);
              if (relu_in && a < 0) a = 0;
              s += a * bf16_to_float(dy[(((size_t)nn * H + y) * W + xx) * C + c]);
            }
        const double got = (double)dw[(size_t)(kh * 3 + kw) * C + c] - 0.5;
        worst = std::fmax(worst, std::fabs(got - s));
        scale = std::fmax(scale, std::fabs(s));
      }
  const bool ok = worst <= 2e-5 * std::fmax(1.0, scale);  // fp32 accumulation of exact bf16 products
  std::printf("%s  N=%d H=%d W=%d C=%d dil=%d relu_in=%d sms=%d  max|err|=%.3g (max|dW|=%.3g)\n", ok ? "ok  " : "FAIL", N,
              H, W, C, dil, relu_in, num_sms, worst, scale);
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run_case(2, 7, 9, 72, 1, 1, 148, 1);    // channel tail (72 = 64 + 8), one slab
  bad += run_case(1, 5, 4, 8, 2, 1, 148, 2);     // map smaller than the dilated window
  bad += run_case(3, 10, 12, 136, 2, 0, 148, 3);  // dilation 2 without the input ReLU, three chunks
  bad += run_case(1, 50, 50, 64, 1, 1, 2, 4);    // several slabs of uneven length (2 "SMs" -> 5 slabs of 500)
  bad += run_case(2, 19, 23, 128, 1, 0, 1, 5);
  bad += run_case(1, 1, 1, 8, 1, 1, 148, 6);     // a single pixel: only the centre tap
  std::printf(bad ? "EMULATION FAILED\n" : "emulation ok\n");
  return bad;
}
