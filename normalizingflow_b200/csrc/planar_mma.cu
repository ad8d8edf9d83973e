// Planar stack (up to 32 tanh planar layers, d = 128) in ONE pass at the HBM rate.
// Replaces L calls of Planar.forward (reference nf/flows_1.py:42-60, quirk Q8).
//
// The reference applies the layers one after the other,  z_{l+1} = z_l + uhat_l tanh(w_l . z_l + b_l);  written
// out, every pre-activation is a linear function of the INPUT row and of the earlier activations:
//     a_l = w_l . x + b_l + sum_{m<l} G[l][m] h_m,   h_l = tanh(a_l),   G[l][m] = w_l . uhat_m   (Gram matrix)
//     z   = x + sum_l h_l uhat_l,      log_det = sum_l log(|1 + (1 - h_l^2) G[l][l]| + 1e-4)
// so the stack is two small dense products (P = X W^T: [N,128] x [128,32];  Z = X + H Uhat: [N,32] x [32,128])
// around a 32-step scalar recurrence per row.  At 8,192 FMA per row the two products are FMA-bound on the CUDA
// cores (0.25 ms per 2^20 rows at 100 % of the fp32 pipe; the per-layer kernel they replace took 2.4 ms), so
// they run on the tensor cores with fp32-class operands: every fp32 value is an fp16 pair hi + lo
// (22 significant bits) and every product three MMAs hi*hi + lo*hi + hi*lo with fp32 accumulation.
// Everything of a 16-row block lives in ONE warp's registers -- the x block is loaded from HBM straight into
// the accumulator-fragment layout (8-byte accesses, full 32-byte sectors), converted to A fragments on the fly,
// and Z accumulates on top of the exact fp32 x -- so there is no shared-memory staging of activations at all
// (which is why these are register-operand mma.sync instructions: tcgen05 operands would have to be written to
// shared memory first, for a kernel whose only remaining cost is its one read and one write of HBM).
// The recurrence runs inside each lane quad (the four lanes that hold a row), activations broadcast by shuffles.
#include <cuda_fp16.h>

#include "nfk_common.cuh"

namespace nfk {

constexpr int PM_D = 128, PM_L = 32, PM_THREADS = 256;

struct PlanarMmaArgs {
  const float* x;
  const float* w;        // [L][128]
  const float* uhat;     // [L][128]
  const float* gram;     // [L][L]  G[l][m] = w_l . uhat_m
  const float* b;        // [L]
  float* out;
  float* logdet;
  long long N;
  int L;
  int accumulate;
};

__device__ __forceinline__ uint32_t pm_pack(float lo, float hi) {          // saturating, NaN stays NaN
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void pm_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pm_pack(a, b);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pm_pack(a - hf.x, b - hf.y);
}
__device__ __forceinline__ void pm_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float pm_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float pm_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float pm_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh to ~3e-7 absolute (MUFU.TANH carries 2^-11): 1 - 2 / (1 + e^(2x))
__device__ __forceinline__ float pm_tanh(float x) {
  const float t = pm_ex2(x * 2.885390081777927f);
  return fmaf(-2.f, pm_rcp(1.f + t), 1.f);
}

__global__ void __launch_bounds__(PM_THREADS, 2)
planar_stack_mma_kernel(const __grid_constant__ PlanarMmaArgs a) {
  // B-operand fragments of both products, hi and lo, in [fragment][lane] order (one 16-byte load per MMA triple)
  __shared__ uint4 s_wfrag[32 * 32];       // GEMM A: fragment kk*4 + j   (k-step kk of 8, layer tile j of 4)
  __shared__ uint4 s_ufrag[32 * 32];       // GEMM B: fragment kk*16 + m  (k-step kk of 2, column tile m of 16)
  __shared__ float4 s_g[32 * 4 * 2];       // [l][t][8]: G[l'][l] for the 8 layers l' = 8jj + 2t + e that lane t owns (0 if l' <= l)
  __shared__ float s_b[32], s_gd[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L;
  for (int i = tid; i < 32 * 32; i += PM_THREADS) {
    const int f = i >> 5, ln = i & 31, g = ln >> 2, t = ln & 3;
    {
      const int kk = f >> 2, j = f & 3;
      const int l = 8 * j + g, k0 = 16 * kk + 2 * t;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (l < L) {
        const float* wr = a.w + (size_t)l * PM_D;
        v[0] = wr[k0], v[1] = wr[k0 + 1], v[2] = wr[k0 + 8], v[3] = wr[k0 + 9];
      }
      uint4 q;
      pm_split(v[0], v[1], q.x, q.z);
      pm_split(v[2], v[3], q.y, q.w);
      s_wfrag[i] = q;                       // {hi R0, hi R1, lo R0, lo R1}
    }
    {
      const int kk = f >> 4, m = f & 15;
      const int col = 8 * m + g, l0 = 16 * kk + 2 * t;
      float v[4];
      const int ls[4] = {l0, l0 + 1, l0 + 8, l0 + 9};
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = ls[e] < L ? a.uhat[(size_t)ls[e] * PM_D + col] : 0.f;
      uint4 q;
      pm_split(v[0], v[1], q.x, q.z);
      pm_split(v[2], v[3], q.y, q.w);
      s_ufrag[i] = q;
    }
  }
  for (int i = tid; i < 32 * 4 * 8; i += PM_THREADS) {
    const int l = i >> 5, t = (i >> 3) & 3, e8 = i & 7;
    const int lp = 8 * (e8 >> 1) + 2 * t + (e8 & 1);                 // owned layer of lane t, slot e8
    reinterpret_cast<float*>(s_g)[i] = (lp > l && lp < L && l < L) ? a.gram[(size_t)lp * L + l] : 0.f;
  }
  if (tid < 32) {
    s_b[tid] = tid < L ? a.b[tid] : 0.f;
    s_gd[tid] = tid < L ? a.gram[(size_t)tid * L + tid] : 0.f;
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const long long n_blocks = (a.N + 15) / 16;
  const long long wstride = (long long)gridDim.x * (PM_THREADS / 32);
  for (long long blk = (long long)blockIdx.x * (PM_THREADS / 32) + warp; blk < n_blocks; blk += wstride) {
    const long long r0 = blk * 16 + g, r1 = r0 + 8;
    const bool live0 = r0 < a.N, live1 = r1 < a.N;
    const float* x0 = a.x + (live0 ? r0 : 0) * PM_D + 2 * t;
    const float* x1 = a.x + (live1 ? r1 : 0) * PM_D + 2 * t;
    // ---- the x block in accumulator-fragment layout: column tile m -> (row g: cols 8m+2t,+1), (row g+8: same)
    float c[16][4];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float2 p = live0 ? __ldg(reinterpret_cast<const float2*>(x0 + 8 * m)) : make_float2(0.f, 0.f);
      const float2 q = live1 ? __ldg(reinterpret_cast<const float2*>(x1 + 8 * m)) : make_float2(0.f, 0.f);
      c[m][0] = p.x, c[m][1] = p.y, c[m][2] = q.x, c[m][3] = q.y;
    }
    // ---- P = X W^T  (16 rows x 32 layers): layer tile j -> s[j][0..1] row g, s[j][2..3] row g+8
    float s[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      uint32_t ah[4], al[4];
      pm_split(c[2 * kk][0], c[2 * kk][1], ah[0], al[0]);
      pm_split(c[2 * kk][2], c[2 * kk][3], ah[1], al[1]);
      pm_split(c[2 * kk + 1][0], c[2 * kk + 1][1], ah[2], al[2]);
      pm_split(c[2 * kk + 1][2], c[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 bf = s_wfrag[(kk * 4 + j) * 32 + lane];
        pm_mma(s[j], al[0], al[1], al[2], al[3], bf.x, bf.y);      // lo * hi
        pm_mma(s[j], ah[0], ah[1], ah[2], ah[3], bf.z, bf.w);      // hi * lo
        pm_mma(s[j], ah[0], ah[1], ah[2], ah[3], bf.x, bf.y);      // hi * hi
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float b0 = s_b[8 * j + 2 * t], b1 = s_b[8 * j + 2 * t + 1];
      s[j][0] += b0, s[j][1] += b1, s[j][2] += b0, s[j][3] += b1;
    }
    // ---- the recurrence, inside the lane quad that holds a row pair: layer l lives in lane (l % 8) / 2
    float h[4][4];
    float ld0 = 0.f, ld1 = 0.f;
    const int qbase = lane & ~3;
#pragma unroll
    for (int l = 0; l < PM_L; ++l) {
      const int j = l >> 3, tl = (l & 7) >> 1, e = l & 1;
      const float a0 = __shfl_sync(0xffffffffu, s[j][e], qbase | tl);
      const float a1 = __shfl_sync(0xffffffffu, s[j][e + 2], qbase | tl);
      const float h0 = pm_tanh(a0), h1 = pm_tanh(a1);                  // flows_1.py:56-57
      if (t == tl) {
        h[j][e] = h0;
        h[j][e + 2] = h1;
      }
      if (l < L) {                                                      // flows_1.py:58-59
        const float gd = s_gd[l];
        ld0 += pm_lg2(fabsf(fmaf(fmaf(-h0, h0, 1.f), gd, 1.f)) + 1e-4f);
        ld1 += pm_lg2(fabsf(fmaf(fmaf(-h1, h1, 1.f), gd, 1.f)) + 1e-4f);
      }
      // later pre-activations owned by this lane pick up G[l'][l] h_l (earlier / own layers: zero entries; whole
      // layer tiles below j are skipped statically)
      const float4 g0 = s_g[(l * 4 + t) * 2], g1 = s_g[(l * 4 + t) * 2 + 1];
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        if (jj < j) continue;
        s[jj][0] = fmaf(gv[2 * jj], h0, s[jj][0]);
        s[jj][1] = fmaf(gv[2 * jj + 1], h0, s[jj][1]);
        s[jj][2] = fmaf(gv[2 * jj], h1, s[jj][2]);
        s[jj][3] = fmaf(gv[2 * jj + 1], h1, s[jj][3]);
      }
    }
    // ---- Z = X + H Uhat: the accumulators already hold the exact fp32 x
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t ah[4], al[4];
      pm_split(h[2 * kk][0], h[2 * kk][1], ah[0], al[0]);
      pm_split(h[2 * kk][2], h[2 * kk][3], ah[1], al[1]);
      pm_split(h[2 * kk + 1][0], h[2 * kk + 1][1], ah[2], al[2]);
      pm_split(h[2 * kk + 1][2], h[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const uint4 bf = s_ufrag[(kk * 16 + m) * 32 + lane];
        pm_mma(c[m], al[0], al[1], al[2], al[3], bf.x, bf.y);
        pm_mma(c[m], ah[0], ah[1], ah[2], ah[3], bf.z, bf.w);
        pm_mma(c[m], ah[0], ah[1], ah[2], ah[3], bf.x, bf.y);
      }
    }
    float* o0 = a.out + r0 * PM_D + 2 * t;
    float* o1 = a.out + r1 * PM_D + 2 * t;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      if (live0) *reinterpret_cast<float2*>(o0 + 8 * m) = make_float2(c[m][0], c[m][1]);
      if (live1) *reinterpret_cast<float2*>(o1 + 8 * m) = make_float2(c[m][2], c[m][3]);
    }
    if (t == 0) {
      constexpr float LN2F = 0.6931471805599453f;
      if (live0) a.logdet[r0] = a.accumulate ? a.logdet[r0] + LN2F * ld0 : LN2F * ld0;
      if (live1) a.logdet[r1] = a.accumulate ? a.logdet[r1] + LN2F * ld1 : LN2F * ld1;
    }
  }
}

// G[l][m] = w_l . uhat_m  (L x L, fp32; one warp per entry)
__global__ void planar_gram_kernel(const float* __restrict__ w, const float* __restrict__ uhat, float* __restrict__ gram,
                                   int d, int L) {
  const int idx = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (idx >= L * L) return;
  const int l = idx / L, m = idx % L;
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) acc = fmaf(w[(size_t)l * d + k], uhat[(size_t)m * d + k], acc);
  acc = warp_sum(acc);
  if (lane == 0) gram[idx] = acc;
}

}  // namespace nfk

using namespace nfk;

extern "C" {

int nfk_planar_gram(const float* w, const float* uhat, float* gram, int d, int L, void* stream) {
  NFK_REQUIRE(d > 0 && L > 0, "planar_gram: bad shape");
  NFK_REQUIRE(w && uhat && gram, "planar_gram: null device pointer");
  const int per_block = 8;
  planar_gram_kernel<<<(L * L + per_block - 1) / per_block, per_block * 32, 0, (cudaStream_t)stream>>>(w, uhat, gram, d, L);
  count_launch();
  return check_launch("planar_gram");
}

int nfk_planar_stack_mma(const float* x, const float* w, const float* uhat, const float* gram, const float* b, float* out,
                         float* logdet, int64_t N, int d, int L, int accumulate, void* stream) {
  NFK_REQUIRE(N >= 0 && d == PM_D && L >= 1 && L <= PM_L, "planar_stack_mma: d must be %d and 1 <= L <= %d (got d=%d L=%d)",
              PM_D, PM_L, d, L);
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && w && uhat && gram && b && out && logdet, "planar_stack_mma: null device pointer");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 7) == 0,
              "planar_stack_mma: x / out must be 8-byte aligned");
  PlanarMmaArgs a{};
  a.x = x;
  a.w = w;
  a.uhat = uhat;
  a.gram = gram;
  a.b = b;
  a.out = out;
  a.logdet = logdet;
  a.N = N;
  a.L = L;
  a.accumulate = accumulate;
  const long long blocks16 = (N + 15) / 16;
  long long grid = (blocks16 + PM_THREADS / 32 - 1) / (PM_THREADS / 32);
  const long long cap = (long long)sm_count() * 2;
  if (grid > cap) grid = cap;
  planar_stack_mma_kernel<<<(unsigned)grid, PM_THREADS, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  return check_launch("planar_stack_mma");
}

}  // extern "C"
