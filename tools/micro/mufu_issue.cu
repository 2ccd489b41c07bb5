// Microbenchmark: does a MUFU.EX2 (quarter-rate pipe) block the scheduler's dispatch for its 8 cycles,
// or can independent FFMA/FFMA2 from the same or another warp issue underneath it?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_issue mufu_issue.cu ; run on a B200
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool PACKED>
__global__ void kern(float* out, long long* cycles, int iters, int warps_mufu) {
  const int warp = threadIdx.x >> 5;
  float m[8];
  float2 f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -0.001f * (threadIdx.x + i); f[i] = make_float2(1.0f + i, 2.0f + i); }
  const float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(0.0001f, -0.0001f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (warp < warps_mufu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
    }
    if (warp >= warps_mufu || warps_mufu > 4) {   // same-warp mix when warps_mufu > 4 (all warps do both)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (PACKED) f[k & 7] = __ffma2_rn(f[k & 7], a, b);
        else { f[k & 7].x = fmaf(f[k & 7].x, a.x, b.x); }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += m[i] + f[i].x + f[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int K, bool PACKED>
void run(const char* label, int threads, int warps_mufu) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  kern<K, PACKED><<<148, threads>>>(out, cyc, iters, warps_mufu);
  kern<K, PACKED><<<148, threads>>>(out, cyc, iters, warps_mufu);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-46s K=%2d %s: %.2f cycles / iteration (8 MUFU%s)\n", label, K, PACKED ? "FFMA2" : "FFMA ",
         double(h) / iters, K ? " + K FMA" : "");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  // (a) one warp per scheduler doing both MUFU and FMA in one instruction stream
  run<0, false>("1 warp/SMSP, same warp", 128, 8);
  run<16, false>("1 warp/SMSP, same warp", 128, 8);
  run<32, false>("1 warp/SMSP, same warp", 128, 8);
  run<64, false>("1 warp/SMSP, same warp", 128, 8);
  run<32, true>("1 warp/SMSP, same warp", 128, 8);
  run<64, true>("1 warp/SMSP, same warp", 128, 8);
  // (b) two warps per scheduler: warps 0-3 MUFU only, warps 4-7 FMA only
  run<0, false>("2 warps/SMSP: MUFU warp + idle warp", 256, 4);
  run<32, false>("2 warps/SMSP: MUFU warp + FMA warp", 256, 4);
  run<64, false>("2 warps/SMSP: MUFU warp + FMA warp", 256, 4);
  run<64, true>("2 warps/SMSP: MUFU warp + FMA warp", 256, 4);
  // (c) two warps per scheduler both doing MUFU + FMA
  run<0, false>("2 warps/SMSP, both mixed", 256, 8);
  run<32, false>("2 warps/SMSP, both mixed", 256, 8);
  run<32, true>("2 warps/SMSP, both mixed", 256, 8);
  return 0;
}
