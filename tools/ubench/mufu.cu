// Micro-benchmark: MUFU throughput per SM sub-partition and the spline element code in isolation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I normalizingflow_b200/csrc -o gpurun_out/ubench_mufu tools/ubench/mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "rqs_math.cuh"
using namespace nfk;

template <int OP>
__device__ __forceinline__ float mu(float x) {
  float y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// NF FFMAs per MUFU, 8 independent chains per thread
template <int OP, int NF>
__global__ void k_mix(float* out, int iters, long long* clk) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.001f * (threadIdx.x + j);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = mu<OP>(v[j]);
#pragma unroll
      for (int f = 0; f < NF; ++f) v[j] = fmaf(v[j], 0.999f, 0.0001f * f);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

struct GenParams {
  float base;
  __device__ __forceinline__ float operator()(int i) const { return base * (float)((i * 7) % 11 - 5); }
  __device__ __forceinline__ float dyn(int b, int i) const { return base * (float)(i - 3); }
};

template <int MODE, bool INV, int JAM>
__global__ void k_elem(float* out, int iters, long long* clk, RqsConsts c) {
  float acc = 0.f;
  float x = -2.9f + 0.01f * threadIdx.x;
  float b = 0.1f + 0.001f * threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    RqsOut o[JAM];
#pragma unroll
    for (int j = 0; j < JAM; ++j) o[j] = rqs_element<MODE, 8, INV, true>(GenParams{b + 0.01f * j}, x + 0.1f * j, c);
#pragma unroll
    for (int j = 0; j < JAM; ++j) {
      acc += o[j].lad + o[j].y;
      b += 1e-4f * o[j].y;
    }
    x = x * 0.99f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// phase A only (knot chains) / phase B only (search + evaluation on fixed knots), JAM independent elements
template <int MODE, bool INV, int JAM>
__global__ void k_phaseA(float* out, int iters, long long* clk, RqsConsts c) {
  float acc = 0.f;
  float b = 0.1f + 0.001f * threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    float cw[JAM][9], ch[JAM][9];
#pragma unroll
    for (int j = 0; j < JAM; ++j) rqs_knots<MODE, 8, INV, true>(GenParams{b + 0.01f * j}, c, cw[j], ch[j]);
#pragma unroll
    for (int j = 0; j < JAM; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 1; k < 8; ++k) s += cw[j][k] * ch[j][k];
      acc += s;
    }
    b += 1e-6f * acc;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int MODE, bool INV, int JAM>
__global__ void k_phaseB(float* out, int iters, long long* clk, RqsConsts c) {
  float acc = 0.f;
  float x = -2.9f + 0.01f * threadIdx.x;
  float b = 0.1f + 0.001f * threadIdx.x;
  float cw[9], ch[9];
  rqs_knots<MODE, 8, INV, true>(GenParams{b}, c, cw, ch);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    RqsOut o[JAM];
#pragma unroll
    for (int j = 0; j < JAM; ++j) o[j] = rqs_eval<MODE, 8, INV, true>(GenParams{b + 0.01f * j}, x + 0.1f * j, c, cw, ch);
#pragma unroll
    for (int j = 0; j < JAM; ++j) {
      acc += o[j].lad + o[j].y;
      b += 1e-4f * o[j].y;
    }
    x = x * 0.99f;
    cw[3] += 1e-7f * acc;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// software-pipelined element (phase A of element i+1 next to phase B of element i); ALT: warps whose
// slice is odd run B then A, the others A then B, as separate basic blocks
template <int MODE, bool INV, bool ALT>
__global__ void k_alt(float* out, int iters, long long* clk, RqsConsts c) {
  float acc = 0.f;
  float x = -2.9f + 0.01f * threadIdx.x;
  float b = 0.1f + 0.001f * threadIdx.x;
  const bool swap = ALT && (((threadIdx.x >> 5) >> 2) & 1);
  float cw[9], ch[9];
  rqs_knots<MODE, 8, INV, true>(GenParams{b}, c, cw, ch);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    RqsOut o;
    if (swap) o = rqs_eval<MODE, 8, INV, true>(GenParams{b}, x, c, cw, ch);
    float cwn[9], chn[9];
    rqs_knots<MODE, 8, INV, true>(GenParams{b + 1e-3f * i}, c, cwn, chn);
    if (!swap) o = rqs_eval<MODE, 8, INV, true>(GenParams{b}, x, c, cw, ch);
#pragma unroll
    for (int k = 0; k < 9; ++k) { cw[k] = cwn[k]; ch[k] = chn[k]; }
    acc += o.lad + o.y;
    b += 1e-4f * o.y;
    x = x * 0.99f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + cw[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

static RqsConsts consts(int K, float B) {
  RqsConsts c;
  c.B = B; c.twoB = 2 * B; c.negB = -B; c.Bnudge = B + 1e-6f; c.min_bin = 1e-3f; c.one_m = 1.f - 1e-3f * K;
  c.min_d = 1e-3f; c.edge_c = 0.5397424f; c.edge_d = 1.f; c.g0 = c.twoB * LOG2E; c.q0 = c.twoB * c.one_m;
  c.kstep = c.twoB * 1e-3f; c.bin_eps = 1.2e-4f; c.K = K; c.scan_order = 1;
  return c;
}

template <class F>
static void run(const char* name, F launch, int threads, int iters, double per_iter_units) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  launch(out, iters, clk, threads); launch(out, iters, clk, threads);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  int wps = threads / 32 / 4;
  printf("%-34s warps/SMSP %2d : %8.1f clk per unit per warp, %8.1f clk per unit per SMSP  (%s)\n", name, wps,
         (double)h / (iters * per_iter_units), (double)h / (iters * per_iter_units * wps), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk);
}

int main() {
  const int iters = 2000;
  for (int threads : {512}) {
    run("EX2 only (unit = 1 MUFU)", [](float* o, int it, long long* c, int t) { k_mix<0, 0><<<148, t>>>(o, it, c); }, threads, iters, 8);
    run("RCP only", [](float* o, int it, long long* c, int t) { k_mix<1, 0><<<148, t>>>(o, it, c); }, threads, iters, 8);
    run("LG2 only", [](float* o, int it, long long* c, int t) { k_mix<2, 0><<<148, t>>>(o, it, c); }, threads, iters, 8);
    run("TANH only", [](float* o, int it, long long* c, int t) { k_mix<3, 0><<<148, t>>>(o, it, c); }, threads, iters, 8);
    run("EX2 + 3 FFMA (unit = group)", [](float* o, int it, long long* c, int t) { k_mix<0, 3><<<148, t>>>(o, it, c); }, threads, iters, 8);
    run("EX2 + 6 FFMA", [](float* o, int it, long long* c, int t) { k_mix<0, 6><<<148, t>>>(o, it, c); }, threads, iters, 8);
    run("EX2 + 8 FFMA", [](float* o, int it, long long* c, int t) { k_mix<0, 8><<<148, t>>>(o, it, c); }, threads, iters, 8);
    run("EX2 + 12 FFMA", [](float* o, int it, long long* c, int t) { k_mix<0, 12><<<148, t>>>(o, it, c); }, threads, iters, 8);
  }
  RqsConsts c = consts(8, 3.f);
  for (int threads : {512, 1024}) {
    run("rqs_element FAST fwd (unit=elem)", [c](float* o, int it, long long* k, int t) { k_elem<2, false, 1><<<148, t>>>(o, it, k, c); }, threads, 500, 1);
    run("rqs_element FAST inv", [c](float* o, int it, long long* k, int t) { k_elem<2, true, 1><<<148, t>>>(o, it, k, c); }, threads, 500, 1);
    run("rqs_element FAST fwd x2 jam", [c](float* o, int it, long long* k, int t) { k_elem<2, false, 2><<<148, t>>>(o, it, k, c); }, threads, 500, 2);
    run("sw-pipelined, all A;B", [c](float* o, int it, long long* k, int t) { k_alt<2, false, false><<<148, t>>>(o, it, k, c); }, threads, 500, 1);
    run("sw-pipelined, alternating", [c](float* o, int it, long long* k, int t) { k_alt<2, false, true><<<148, t>>>(o, it, k, c); }, threads, 500, 1);
    run("phase A fwd", [c](float* o, int it, long long* k, int t) { k_phaseA<2, false, 1><<<148, t>>>(o, it, k, c); }, threads, 500, 1);
    run("phase A fwd x2 jam", [c](float* o, int it, long long* k, int t) { k_phaseA<2, false, 2><<<148, t>>>(o, it, k, c); }, threads, 500, 2);
    run("phase B fwd", [c](float* o, int it, long long* k, int t) { k_phaseB<2, false, 1><<<148, t>>>(o, it, k, c); }, threads, 500, 1);
    run("phase B fwd x2 jam", [c](float* o, int it, long long* k, int t) { k_phaseB<2, false, 2><<<148, t>>>(o, it, k, c); }, threads, 500, 2);
    run("phase B fwd x4 jam", [c](float* o, int it, long long* k, int t) { k_phaseB<2, false, 4><<<148, t>>>(o, it, k, c); }, threads, 500, 4);
    run("phase B inv", [c](float* o, int it, long long* k, int t) { k_phaseB<2, true, 1><<<148, t>>>(o, it, k, c); }, threads, 500, 1);
    run("rqs_element HYBRID fwd", [c](float* o, int it, long long* k, int t) { k_elem<1, false, 1><<<148, t>>>(o, it, k, c); }, threads, 500, 1);
  }
  return 0;
}
