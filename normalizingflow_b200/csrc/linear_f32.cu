// fp32 CUDA-core conditioner layer (parity mode): Y = act(X W^T + b), one nn.Linear (+Tanh) of
// FCNN (reference nf/flows.py:26-35).  Plain FFMA accumulation in fp32, k ascending, so the
// result sits at fp32 round-off of the reference's addmm; the tensor-core path for throughput is
// linear_bf16.cu.  Also the generic fp32 GEMM used by the conditioner backward.
#include "nfk_common.cuh"

namespace nfk {

constexpr int BM = 128, BN = 64, BK = 16;

// A(m,k) = A[m*lda + k] (ta == 0) or A[k*lda + m] (ta == 1); B(k,n) = B[k*ldb + n] (tb == 0)
// or B[n*ldb + k] (tb == 1).  C[m*ldc + n] = act(sum_k A(m,k) B(k,n) + bias[n]) (+ C).
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ Bm,
                long long ldb, const float* __restrict__ bias, float* __restrict__ C,
                long long ldc, long long M, long long N, long long K, int act, int accumulate,
                long long kchunk) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 8 x 4 micro-tile
  const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // split-K (kchunk > 0): blockIdx.z owns K range [z*kchunk, (z+1)*kchunk) and ADDS its partial tile
  // into C with atomics (C zeroed by the launcher) -- weight gradients have a tiny output and a
  // batch-long K, which would otherwise run on a handful of CTAs
  const long long kb = kchunk > 0 ? (long long)blockIdx.z * kchunk : 0;
  const long long ke = kchunk > 0 ? (kb + kchunk < K ? kb + kchunk : K) : K;
  for (long long k0 = kb; k0 < ke; k0 += BK) {
    // A tile: BM x BK
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (TA) {
        m = idx % BM;
        k = idx / BM;
      } else {
        k = idx % BK;
        m = idx / BK;
      }
      const long long gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < ke) v = TA ? A[gk * lda + gm] : A[gm * lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      const int idx = tid + i * 256;
      int n, k;
      if (TB) {
        k = idx % BK;
        n = idx / BK;
      } else {
        n = idx % BN;
        k = idx / BN;
      }
      const long long gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < ke) v = TB ? Bm[gn * ldb + gk] : Bm[gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long gm = m0 + ty + 16 * i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long gn = n0 + tx + 16 * j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (act == 1) v = tanhf(v);
      float* cp = C + gm * ldc + gn;
      if (kchunk > 0)
        atomicAdd(cp, v);
      else
        *cp = accumulate ? *cp + v : v;
    }
  }
}

static int launch_gemm(const float* A, long long lda, int ta, const float* Bm, long long ldb, int tb,
                       const float* bias, float* C, long long ldc, long long M, long long N,
                       long long K, int act, int accumulate, cudaStream_t st, const char* what) {
  if (M == 0 || N == 0) return NFK_OK;
  const long long gy = (M + BM - 1) / BM, gx = (N + BN - 1) / BN;
  NFK_REQUIRE(gy <= 65535 * 1024LL, "%s: too many rows", what);
  // rows beyond the 65535 limit of gridDim.y are handled by launching in slabs
  // split-K when the output tile grid cannot fill the chip and K is long (no bias / activation then)
  long long kchunk = 0, splits = 1;
  const long long sms = sm_count();
  if (!bias && act == 0 && gx * gy * 2 <= sms && K >= 1024) {
    splits = sms / (gx * gy);
    if (splits > (K + 255) / 256) splits = (K + 255) / 256;
    if (splits > 1) {
      kchunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
      splits = (K + kchunk - 1) / kchunk;
      if (!accumulate) {
        cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st);
        if (e != cudaSuccess) {
          set_error("%s: cannot zero the split-K output: %s", what, cudaGetErrorString(e));
          return NFK_ECUDA;
        }
      }
    } else {
      splits = 1;
    }
  }
  const long long slab = 65535;
  for (long long y0 = 0; y0 < gy; y0 += slab) {
    const long long ny = (gy - y0 < slab) ? gy - y0 : slab;
    dim3 grid((unsigned)gx, (unsigned)ny, (unsigned)splits);
    const float* A2 = A + (ta ? y0 * BM : y0 * BM * lda);
    float* C2 = C + y0 * BM * ldc;
    const long long M2 = M - y0 * BM < ny * BM ? M - y0 * BM : ny * BM;
    if (ta) {
      if (tb) gemm_f32_kernel<true, true><<<grid, 256, 0, st>>>(A2, lda, Bm, ldb, bias, C2, ldc, M2, N, K, act, accumulate, kchunk);
      else    gemm_f32_kernel<true, false><<<grid, 256, 0, st>>>(A2, lda, Bm, ldb, bias, C2, ldc, M2, N, K, act, accumulate, kchunk);
    } else {
      if (tb) gemm_f32_kernel<false, true><<<grid, 256, 0, st>>>(A2, lda, Bm, ldb, bias, C2, ldc, M2, N, K, act, accumulate, kchunk);
      else    gemm_f32_kernel<false, false><<<grid, 256, 0, st>>>(A2, lda, Bm, ldb, bias, C2, ldc, M2, N, K, act, accumulate, kchunk);
    }
    count_launch();
    if (int rc = check_launch(what)) return rc;
  }
  return NFK_OK;
}

}  // namespace nfk

using namespace nfk;

extern "C" {

int nfk_linear_f32(const float* X, int64_t ldx, const float* W, const float* b, float* Y,
                   int64_t M, int K, int Nout, int act, void* stream) {
  NFK_REQUIRE(M >= 0 && K > 0 && Nout > 0, "linear_f32: bad shape M=%lld K=%d N=%d", (long long)M, K,
              Nout);
  NFK_REQUIRE(act == 0 || act == 1, "linear_f32: act must be 0 (identity) or 1 (tanh)");
  NFK_REQUIRE(ldx >= K, "linear_f32: row stride %lld < K=%d", (long long)ldx, K);
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(X && W && Y, "linear_f32: null device pointer");
  return launch_gemm(X, ldx, 0, W, K, 1, b, Y, Nout, M, Nout, K, act, 0, (cudaStream_t)stream,
                     "linear_f32");
}

int nfk_gemm_f32(const float* A, int64_t lda, int ta, const float* Bm, int64_t ldb, int tb,
                 float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate,
                 void* stream) {
  NFK_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm_f32: negative shape");
  if (M == 0 || N == 0) return NFK_OK;
  NFK_REQUIRE(A && Bm && C, "gemm_f32: null device pointer");
  return launch_gemm(A, lda, ta, Bm, ldb, tb, nullptr, C, ldc, M, N, K, 0, accumulate,
                     (cudaStream_t)stream, "gemm_f32");
}

}  // extern "C"
