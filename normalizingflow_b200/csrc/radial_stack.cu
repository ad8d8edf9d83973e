// Radial layers (reference nf/flows_1.py:85-97, quirk Q9) at the HBM rate.
//
// (1) radial_stack_ps_kernel -- a run of L per-sample radial layers in ONE pass: 4 or 8 lanes share a row (16 B
//     accesses, 128 contiguous bytes per lane group), the row stays in registers across all layers, each layer
//     costs one 2- or 3-step shuffle reduction of ||x - x0||^2 per row; layer constants (x0, alpha, beta) sit in
//     shared memory.  Per-layer launches moved 2 * d * 4 B per row and layer; the stack moves them once.
// (2) the reference's batch-global mode (r = ONE Frobenius norm over the whole batch) needs the sum of squares
//     of a layer's input before it can transform a single row, so a layer is inherently two passes over the
//     batch.  radial_apply_sumsq_kernel makes the second pass of layer l ALSO the first pass of layer l+1:
//     it writes z = x + beta h (x - x0_l) and accumulates sum (z - x0_{l+1})^2 on the fly, so a stack costs one
//     read + one write per layer (plus one read for the first layer) instead of two reads + one write.
//     Across ranks the scalar is still all-reduced between launches by the host.
#include "nfk_common.cuh"

namespace nfk {

constexpr int RS_THREADS = 256;

template <int NV4, int G>      // G lanes share a row, NV4 float4 per lane: d = 4 * G * NV4
__global__ void __launch_bounds__(RS_THREADS)
radial_stack_ps_kernel(const float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ log_alpha,
                       const float* __restrict__ beta_raw, float* __restrict__ out, float* __restrict__ logdet,
                       long long N, int L, int accumulate) {
  constexpr int d = 4 * G * NV4;
  extern __shared__ float sm[];
  float* sx0 = sm;                    // [L][d]
  float* sal = sm + (size_t)L * d;    // [L] alpha
  float* sbe = sal + L;               // [L] beta
  for (int i = threadIdx.x; i < L * d; i += RS_THREADS) sx0[i] = x0[i];
  for (int i = threadIdx.x; i < L; i += RS_THREADS) {
    const float alpha = expf(log_alpha[i]);
    sal[i] = alpha;
    sbe[i] = -alpha + logf(1.f + expf(beta_raw[i]));                  // flows_1.py:92
  }
  __syncthreads();
  const int g = threadIdx.x & (G - 1);
  const long long rows_per_block = RS_THREADS / G;
  for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < N; row0 += (long long)gridDim.x * rows_per_block) {
    const long long row = row0 + threadIdx.x / G;
    const bool live = row < N;
    float4 v[NV4];
    const float4* xr = reinterpret_cast<const float4*>(x + (live ? row : 0) * d);
#pragma unroll
    for (int i = 0; i < NV4; ++i) v[i] = live ? ldg_stream4(xr + g + G * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float ld = 0.f;
    for (int l = 0; l < L; ++l) {
      const float4* c = reinterpret_cast<const float4*>(sx0 + (size_t)l * d);
      float4 df[NV4];
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const float4 cc = c[g + G * i];
        df[i] = make_float4(v[i].x - cc.x, v[i].y - cc.y, v[i].z - cc.z, v[i].w - cc.w);
        acc = fmaf(df[i].x, df[i].x, acc);
        acc = fmaf(df[i].y, df[i].y, acc);
        acc = fmaf(df[i].z, df[i].z, acc);
        acc = fmaf(df[i].w, df[i].w, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (G == 8) acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      const float alpha = sal[l], beta = sbe[l];
      const float r = sqrtf(acc);                                      // per-row norm (per_sample mode)
      const float ar = alpha + r;
      const float h = 1.f / ar;
      const float bh = beta * h;
#pragma unroll
      for (int i = 0; i < NV4; ++i) {                                  // flows_1.py:93
        v[i].x = fmaf(bh, df[i].x, v[i].x);
        v[i].y = fmaf(bh, df[i].y, v[i].y);
        v[i].z = fmaf(bh, df[i].z, v[i].z);
        v[i].w = fmaf(bh, df[i].w, v[i].w);
      }
      ld += (float)(d - 1) * logf(1.f + bh) + logf(1.f + bh - beta * r / (ar * ar));   // flows_1.py:94-95
    }
    if (live) {
      float4* orow = reinterpret_cast<float4*>(out + row * d);
#pragma unroll
      for (int i = 0; i < NV4; ++i) stg_stream4(orow + g + G * i, v[i]);
      if (g == 0) logdet[row] = accumulate ? logdet[row] + ld : ld;
    }
  }
}

// batch-global mode, vectorised: every thread keeps the same 4 columns across its grid-stride iterations
// (RS_THREADS * 4 is a multiple of d), so x0 / x0_next live in registers and there is no per-element modulo.
//   APPLY: out = x + beta h (x - x0), h = 1 / (alpha + sqrt(sumsq[0]));  SUMSQ: sumsq_next += sum (value - x0n)^2
template <bool APPLY, bool SUMSQ>
__global__ void __launch_bounds__(RS_THREADS)
radial_global_kernel(const float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ log_alpha,
                     const float* __restrict__ beta_raw, const float* __restrict__ sumsq, const float* __restrict__ x0n,
                     float* __restrict__ sumsq_next, float* __restrict__ out, float* __restrict__ logdet,
                     long long total4, int d, int accumulate) {
  const int col = (threadIdx.x * 4) % d;
  float4 c = make_float4(0.f, 0.f, 0.f, 0.f), cn = c;
  float bh = 0.f;
  if (APPLY) {
    c = *reinterpret_cast<const float4*>(x0 + col);
    const float alpha = expf(log_alpha[0]);
    const float beta = -alpha + logf(1.f + expf(beta_raw[0]));         // flows_1.py:92
    const float r = sqrtf(sumsq[0]);                                   // flows_1.py:90: ONE norm for the batch
    const float ar = alpha + r;
    bh = beta / ar;
    if (blockIdx.x == 0 && threadIdx.x == 0) {                         // flows_1.py:94-95: log_det has shape [1]
      const float ld = (float)(d - 1) * logf(1.f + bh) + logf(1.f + bh - beta * r / (ar * ar));
      logdet[0] = accumulate ? logdet[0] + ld : ld;
    }
  }
  if (SUMSQ) cn = *reinterpret_cast<const float4*>(x0n + col);
  float acc = 0.f;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* o4 = reinterpret_cast<float4*>(out);
  for (long long i = (long long)blockIdx.x * RS_THREADS + threadIdx.x; i < total4; i += (long long)gridDim.x * RS_THREADS) {
    float4 v = ldg_stream4(x4 + i);
    if (APPLY) {
      v.x = fmaf(bh, v.x - c.x, v.x);                                  // flows_1.py:93
      v.y = fmaf(bh, v.y - c.y, v.y);
      v.z = fmaf(bh, v.z - c.z, v.z);
      v.w = fmaf(bh, v.w - c.w, v.w);
      stg_stream4(o4 + i, v);
    }
    if (SUMSQ) {
      const float a0 = v.x - cn.x, a1 = v.y - cn.y, a2 = v.z - cn.z, a3 = v.w - cn.w;
      acc = fmaf(a0, a0, acc);
      acc = fmaf(a1, a1, acc);
      acc = fmaf(a2, a2, acc);
      acc = fmaf(a3, a3, acc);
    }
  }
  if (SUMSQ) {
    __shared__ float red[RS_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < RS_THREADS / 32 ? red[threadIdx.x] : 0.f;
      v = warp_sum(v);
      if (threadIdx.x == 0) atomicAdd(sumsq_next, v);
    }
  }
}

}  // namespace nfk

using namespace nfk;

extern "C" {

int nfk_radial_stack(const float* x, const float* x0, const float* log_alpha, const float* beta, float* out,
                     float* logdet, int64_t N, int d, int L, int accumulate, void* stream) {
  NFK_REQUIRE(N >= 0 && L >= 1, "radial_stack: bad shape");
  NFK_REQUIRE(d == 32 || d == 64 || d == 128 || d == 256, "radial_stack: d must be 32, 64, 128 or 256 (got %d)", d);
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && x0 && log_alpha && beta && out && logdet, "radial_stack: null device pointer");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x0)) & 15) == 0,
              "radial_stack: pointers must be 16-byte aligned");
  const size_t smem = ((size_t)L * d + 2 * (size_t)L) * sizeof(float);
  NFK_REQUIRE(smem <= 200 * 1024, "radial_stack: L*d = %d too large for shared memory; split the stack", L * d);
  cudaStream_t st = (cudaStream_t)stream;
  // lanes per row: 4 where a lane's share (8 float4) still fits the registers -- the per-layer scalar work
  // (sqrt, reciprocal, two logarithms) is replicated in every lane of a row, so fewer lanes per row is less work
#define NFK_RS(NV, GG)                                                                                      \
  do {                                                                                                      \
    long long grid = (N + RS_THREADS / GG - 1) / (RS_THREADS / GG);                                         \
    const long long cap = (long long)sm_count() * 8;                                                        \
    if (grid > cap) grid = cap;                                                                             \
    cudaFuncSetAttribute(radial_stack_ps_kernel<NV, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    radial_stack_ps_kernel<NV, GG><<<(unsigned)grid, RS_THREADS, smem, st>>>(x, x0, log_alpha, beta, out, logdet, N, L, accumulate); \
  } while (0)
  if (d == 32) NFK_RS(1, 8);
  else if (d == 64) NFK_RS(4, 4);
  else if (d == 128) NFK_RS(8, 4);
  else NFK_RS(8, 8);
#undef NFK_RS
  count_launch();
  return check_launch("radial_stack");
}

int nfk_radial_global(const float* x, const float* x0, const float* log_alpha, const float* beta, const float* sumsq,
                      const float* x0_next, float* sumsq_next, float* out, float* logdet, int64_t N, int d,
                      int accumulate, void* stream) {
  NFK_REQUIRE(N >= 0 && d > 0 && d % 4 == 0 && (RS_THREADS * 4) % d == 0,
              "radial_global: d must divide %d and be a multiple of 4 (got %d)", RS_THREADS * 4, d);
  const bool apply = x0 != nullptr, ssq = x0_next != nullptr;
  NFK_REQUIRE(apply || ssq, "radial_global: nothing to do");
  NFK_REQUIRE(!apply || (log_alpha && beta && sumsq && out && logdet), "radial_global: apply needs log_alpha, beta, sumsq, out, logdet");
  NFK_REQUIRE(!ssq || sumsq_next, "radial_global: x0_next needs sumsq_next");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x0) |
                     reinterpret_cast<uintptr_t>(x0_next)) & 15) == 0,
              "radial_global: pointers must be 16-byte aligned");
  const long long total4 = (long long)N * d / 4;
  long long grid = (total4 + RS_THREADS - 1) / RS_THREADS;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (apply && ssq)
    radial_global_kernel<true, true><<<(unsigned)grid, RS_THREADS, 0, st>>>(x, x0, log_alpha, beta, sumsq, x0_next, sumsq_next, out,
                                                                            logdet, total4, d, accumulate);
  else if (apply)
    radial_global_kernel<true, false><<<(unsigned)grid, RS_THREADS, 0, st>>>(x, x0, log_alpha, beta, sumsq, x0_next, sumsq_next,
                                                                             out, logdet, total4, d, accumulate);
  else
    radial_global_kernel<false, true><<<(unsigned)grid, RS_THREADS, 0, st>>>(x, x0, log_alpha, beta, sumsq, x0_next, sumsq_next,
                                                                             out, logdet, total4, d, accumulate);
  count_launch();
  return check_launch("radial_global");
}

}  // extern "C"
