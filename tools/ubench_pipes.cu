// Micro-benchmark (development tool, not product): issue rate of the FP64 pipe and of the
// float<->double conversions on sm_100a, to budget PsRoiAlign's bit-exact fp64 blend.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, float seed, int iters) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float f0 = seed, f1 = seed + 1, f2 = seed + 2, f3 = seed + 3, f4 = seed + 4, f5 = seed + 5, f6 = seed + 6, f7 = seed + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c); a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c); }
    if (OP == 1) { a0 = __dmul_rn(a0, m); a1 = __dmul_rn(a1, m); a2 = __dmul_rn(a2, m); a3 = __dmul_rn(a3, m); a4 = __dmul_rn(a4, m); a5 = __dmul_rn(a5, m); a6 = __dmul_rn(a6, m); a7 = __dmul_rn(a7, m); }
    if (OP == 2) { a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c); a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c); }
    if (OP == 3) {  // float->double->float round trip: 1 F2F.F64.F32 + 1 F2F.F32.F64 per lane-op pair
      f0 = __double2float_rn((double)f0 ) + 1.f; f1 = __double2float_rn((double)f1) + 1.f; f2 = __double2float_rn((double)f2) + 1.f; f3 = __double2float_rn((double)f3) + 1.f;
      f4 = __double2float_rn((double)f4) + 1.f; f5 = __double2float_rn((double)f5) + 1.f; f6 = __double2float_rn((double)f6) + 1.f; f7 = __double2float_rn((double)f7) + 1.f; }
    if (OP == 4) {  // fp32 FMA reference
      f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f);
      f4 = fmaf(f4, 1.0000001f, 1e-9f); f5 = fmaf(f5, 1.0000001f, 1e-9f); f6 = fmaf(f6, 1.0000001f, 1e-9f); f7 = fmaf(f7, 1.0000001f, 1e-9f); }
    if (OP == 5) {  // float->double only, consumed by a DADD (so: 1 cvt + 1 dadd per op)
      a0 = __dadd_rn(a0, (double)f0); a1 = __dadd_rn(a1, (double)f1); a2 = __dadd_rn(a2, (double)f2); a3 = __dadd_rn(a3, (double)f3);
      a4 = __dadd_rn(a4, (double)f4); a5 = __dadd_rn(a5, (double)f5); a6 = __dadd_rn(a6, (double)f6); a7 = __dadd_rn(a7, (double)f7);
      f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; f4 += 1.f; f5 += 1.f; f6 += 1.f; f7 += 1.f; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}

template <int OP>
void run(const char* name, int ops_per_iter) {
  const int blocks = 148 * 4, threads = 256, iters = 4096;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<blocks, threads>>>(out, 1.0f, 64);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(out, 1.0f, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double lane_ops = (double)blocks * threads * iters * ops_per_iter;
  printf("%-28s %8.3f ms  %8.2f Tlane-op/s  (%.1f lane-ops/clk/SM @1.9GHz)\n", name, ms, lane_ops / ms * 1e-9,
         lane_ops / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
}

int main() {
  run<0>("DADD", 8); run<1>("DMUL", 8); run<2>("DFMA", 8); run<3>("F2F f32->f64->f32 (+FADD)", 8);
  run<4>("FFMA", 8); run<5>("F2F f32->f64 + DADD", 8);
  cudaError_t e = cudaDeviceSynchronize(); printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
