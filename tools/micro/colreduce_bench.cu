// Micro-benchmark behind the design of the batch-norm statistics kernels: column sums (+ squares) of a [rows, C] bf16
// matrix with different ways of combining the blocks' partial sums.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   V0 no combine (partials to scratch)      V1 scalar atomics          V2 16-byte atomics
//   V3 16-byte atomics over K replicas       V4 = V3 + last-block collapse (ticket)   V5 = V2 + last block (ticket)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int V>
__global__ void __launch_bounds__(256) k(const __nv_bfloat16* __restrict__ x, long long rows, int C, int cgb, int K,
                                         float* __restrict__ sums, float* __restrict__ out, unsigned* ticket) {
  const int rl = 256 / cgb, cg = threadIdx.x % cgb, lane_row = threadIdx.x / cgb;
  const int c0 = (blockIdx.x * cgb + cg) * 8;
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = q[j] = 0.f;
  if (c0 < C) {
#pragma unroll 4
    for (long long r = (long long)blockIdx.y * rl + lane_row; r < rows; r += (long long)gridDim.y * rl) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + r * C + c0));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float2 f = __bfloat1622float2(h[kk]);
        a[2 * kk] += f.x; a[2 * kk + 1] += f.y;
        q[2 * kk] = fmaf(f.x, f.x, q[2 * kk]); q[2 * kk + 1] = fmaf(f.y, f.y, q[2 * kk + 1]);
      }
    }
  }
  __shared__ float s_a[2048], s_q[2048];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s_a[threadIdx.x * 8 + j] = a[j]; s_q[threadIdx.x * 8 + j] = q[j]; }
  __syncthreads();
  const int nch = cgb * 8;
  float t = 0.f, t2 = 0.f;
  int c = C;
  if ((int)threadIdx.x < nch) {
    c = blockIdx.x * nch + threadIdx.x;
    if (c < C) for (int r = 0; r < rl; ++r) { t += s_a[r * nch + threadIdx.x]; t2 += s_q[r * nch + threadIdx.x]; }
  }
  const float u1 = __shfl_down_sync(~0u, t, 1), u2 = __shfl_down_sync(~0u, t, 2), u3 = __shfl_down_sync(~0u, t, 3);
  const float v1 = __shfl_down_sync(~0u, t2, 1), v2 = __shfl_down_sync(~0u, t2, 2), v3 = __shfl_down_sync(~0u, t2, 3);
  float* dst = sums;
  if (V == 3 || V == 4) dst = sums + (size_t)(blockIdx.y % K) * 2 * C;
  if (c < C) {
    if (V == 0) { out[((size_t)blockIdx.y * 2) * C + c] = t; out[((size_t)blockIdx.y * 2 + 1) * C + c] = t2; }
    if (V == 1) { atomicAdd(dst + c, t); atomicAdd(dst + C + c, t2); }
    if (V >= 2 && (threadIdx.x & 3) == 0) {
      atomicAdd(reinterpret_cast<float4*>(dst + c), make_float4(t, u1, u2, u3));
      atomicAdd(reinterpret_cast<float4*>(dst + C + c), make_float4(t2, v1, v2, v3));
    }
  }
  if (V < 4) return;
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1u);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int reps = (V == 4) ? K : 1;
  for (int cc = threadIdx.x; cc < 2 * C; cc += 256) {
    float s = 0.f;
    for (int kk = 0; kk < reps; ++kk) { s += __ldcg(sums + (size_t)kk * 2 * C + cc); sums[(size_t)kk * 2 * C + cc] = 0.f; }
    out[cc] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

template <int V>
float run(const __nv_bfloat16* x, long long rows, int C, int per_sm, int min_rows, int K, float* sums, float* out,
          unsigned* ticket, int* blocks) {
  int cgb = 1;
  while (cgb * 2 <= 32 && cgb * 2 <= C / 8) cgb *= 2;
  const int gx = (C / 8 + cgb - 1) / cgb, rl = 256 / cgb;
  long long slabs = (rows + (long long)rl * min_rows - 1) / ((long long)rl * min_rows);
  long long cap = 148LL * per_sm / gx; if (cap < 1) cap = 1;
  if (slabs > cap) slabs = cap;
  dim3 grid(gx, (unsigned)slabs);
  *blocks = gx * (int)slabs;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) k<V><<<grid, 256>>>(x, rows, C, cgb, K, sums, out, ticket);
  CK(cudaDeviceSynchronize());
  const int reps = 50;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k<V><<<grid, 256>>>(x, rows, C, cgb, K, sums, out, ticket);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms * 1e3f / reps;
}

int main() {
  const long long shapes[][2] = {{115200, 64}, {115200, 256}, {28800, 128}, {28800, 512}, {7200, 256}, {7200, 512}, {7200, 1024}, {7200, 2048}};
  __nv_bfloat16* x; float *sums, *out; unsigned* ticket;
  CK(cudaMalloc(&x, 115200LL * 256 * 2 * 4));  // x4: rotate through copies? (kept simple: one copy, L2-warm for small shapes)
  CK(cudaMemset(x, 0, 115200LL * 256 * 2));
  CK(cudaMalloc(&sums, 64 * 2 * 2048 * 4)); CK(cudaMemset(sums, 0, 64 * 2 * 2048 * 4));
  CK(cudaMalloc(&out, 2048LL * 2 * 2048 * 4)); CK(cudaMalloc(&ticket, 4)); CK(cudaMemset(ticket, 0, 4));
  printf("us per launch (back to back, includes launch gap). columns: V0 none | V1 scalar | V2 v4 | V3 v4 xK | V4 v4 xK+last | V5 v4+last\n");
  for (auto& sh : shapes) {
    const long long rows = sh[0]; const int C = (int)sh[1];
    for (int per_sm = 4; per_sm <= 8; per_sm += 4)
      for (int min_rows = 2; min_rows <= 8; min_rows *= 2) {
        int K = 2048 / C; if (K > 16) K = 16; if (K < 1) K = 1;
        int nb;
        const float t0 = run<0>(x, rows, C, per_sm, min_rows, K, sums, out, ticket, &nb);
        const float t1 = run<1>(x, rows, C, per_sm, min_rows, K, sums, out, ticket, &nb);
        const float t2 = run<2>(x, rows, C, per_sm, min_rows, K, sums, out, ticket, &nb);
        const float t3 = run<3>(x, rows, C, per_sm, min_rows, K, sums, out, ticket, &nb);
        const float t4 = run<4>(x, rows, C, per_sm, min_rows, K, sums, out, ticket, &nb);
        const float t5 = run<5>(x, rows, C, per_sm, min_rows, K, sums, out, ticket, &nb);
        printf("%7lld x %4d per_sm %d min_rows %d blocks %4d K %2d : %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f\n", rows, C, per_sm,
               min_rows, nb, K, t0, t1, t2, t3, t4, t5);
      }
  }
  return 0;
}
