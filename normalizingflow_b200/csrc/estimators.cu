// Free-energy estimators on the device: the BAR fixed point of
// applications/src/bar.py:16-67 and log-mean-exp reweighting (applications/src/test.py:66-68,
// dynamics.py:26-36) — consumers of sample()/evaluate() outputs that already live on the GPU.
// fp64 accumulation; one thread block iterates the whole fixed point, so there is no host round
// trip per iteration.
#include "nfk_common.cuh"

namespace nfk {

struct LSE {            // running log-sum-exp: sum = s * exp(m)
  double m, s;
};
__device__ __forceinline__ void lse_add(LSE& a, double v) {
  if (v > a.m) {
    a.s = a.s * exp(a.m - v) + 1.0;
    a.m = v;
  } else {
    a.s += exp(v - a.m);
  }
}
__device__ __forceinline__ void lse_merge(LSE& a, const LSE& b) {
  if (b.m > a.m) {
    a.s = a.s * exp(a.m - b.m) + b.s;
    a.m = b.m;
  } else if (b.s > 0.0) {
    a.s += b.s * exp(b.m - a.m);
  }
}
__device__ __forceinline__ LSE lse_block(LSE v, LSE* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    LSE t;
    t.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    t.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    lse_merge(v, t);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  LSE r = red[0];
  for (int i = 1; i < nw; ++i) lse_merge(r, red[i]);
  return r;
}

// log f(W) = -log(1 + exp(arg)) evaluated as in bar.py:41-43 (max-shifted)
__device__ __forceinline__ double log_fermi(double arg) {
  const double mx = arg > 0.0 ? arg : 0.0;
  return -mx - log(exp(-mx) + exp(arg - mx));
}

template <typename T>
__global__ void __launch_bounds__(1024)
bar_kernel(const T* __restrict__ wF, const T* __restrict__ wR, long long TF, long long TR, double dF, int max_iter,
           double rtol, double* __restrict__ out) {
  __shared__ LSE red[32];
  __shared__ double s_dF;
  const double M = log((double)TF / (double)TR);
  int it = 0;
  for (; it < max_iter; ++it) {
    LSE a{-INFINITY, 0.0}, b{-INFINITY, 0.0};
    for (long long i = threadIdx.x; i < TF; i += blockDim.x) lse_add(a, log_fermi(M + (double)wF[i] - dF));
    for (long long i = threadIdx.x; i < TR; i += blockDim.x) {
      const double w = (double)wR[i];
      lse_add(b, log_fermi(M - w - dF) - w);
    }
    const LSE ra = lse_block(a, red);
    const LSE rb = lse_block(b, red);
    const double log_numer = log(ra.s) + ra.m - log((double)TF);          // bar.py:44
    const double log_denom = log(rb.s) + rb.m - log((double)TR);          // bar.py:52
    const double dF_new = log_denom - log_numer;                          // -BARzero + DeltaF (bar.py:62)
    const double rel = fabs((dF_new - dF) / dF_new);
    dF = dF_new;
    if (it > 0 && rel < rtol) {                                           // bar.py:63-66
      ++it;
      break;
    }
  }
  if (threadIdx.x == 0) {
    out[0] = dF;
    out[1] = (double)it;
  }
  (void)s_dF;
}

// out[c] = logsumexp_r a[r, c] - log(rows)   (log-mean-exp over the leading dimension)
__global__ void __launch_bounds__(256)
log_mean_exp_kernel(const float* __restrict__ a, float* __restrict__ out, long long rows, long long cols) {
  __shared__ LSE red[32];
  const long long c = blockIdx.x;
  LSE v{-INFINITY, 0.0};
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) lse_add(v, (double)a[r * cols + c]);
  const LSE t = lse_block(v, red);
  if (threadIdx.x == 0) out[c] = (float)(log(t.s) + t.m - log((double)rows));
}

}  // namespace nfk

using namespace nfk;

extern "C" {

int nfk_bar(const void* w_F, const void* w_R, int64_t T_F, int64_t T_R, int is_f64, double DeltaF, int max_iter,
            double rtol, double* out, void* stream) {
  NFK_REQUIRE(T_F > 0 && T_R > 0, "bar: need at least one forward and one reverse work value");
  NFK_REQUIRE(max_iter >= 1, "bar: maximum_iterations must be positive");
  NFK_REQUIRE(w_F && w_R && out, "bar: null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (is_f64)
    bar_kernel<double><<<1, 1024, 0, st>>>(reinterpret_cast<const double*>(w_F), reinterpret_cast<const double*>(w_R),
                                           T_F, T_R, DeltaF, max_iter, rtol, out);
  else
    bar_kernel<float><<<1, 1024, 0, st>>>(reinterpret_cast<const float*>(w_F), reinterpret_cast<const float*>(w_R), T_F,
                                          T_R, DeltaF, max_iter, rtol, out);
  count_launch();
  return check_launch("bar");
}

int nfk_log_mean_exp(const float* a, float* out, int64_t rows, int64_t cols, void* stream) {
  NFK_REQUIRE(rows > 0 && cols >= 0 && cols < (1LL << 31), "log_mean_exp: bad shape");
  if (cols == 0) return NFK_OK;
  NFK_REQUIRE(a && out, "log_mean_exp: null device pointer");
  log_mean_exp_kernel<<<(unsigned)cols, 256, 0, (cudaStream_t)stream>>>(a, out, rows, cols);
  count_launch();
  return check_launch("log_mean_exp");
}

}  // extern "C"
