// Backward of the RQS coupling transform: replaces autograd through NSF_CL's op chain after the
// conditioner (reference nf/flows.py:232-239 / :246-253 + nf/utils.py:27-152).  Given
// grad_out [N, size*dim] and grad_logdet [N] it produces grad_x [N, size*dim] (direct path only;
// the path through the conditioner flows back through grad_params) and grad_params
// [N, F_t, 3K-1].  One thread per spline: the forward is recomputed (bin search on the EXACT knot
// chain so the bin equals the forward kernel's), then the adjoints of the rational-quadratic
// expression, of the cumsum/softmax/softmax chain (quirk Q1) and of the double softplus (Q2) are
// propagated by hand.
#include "rqs_bwd_math.cuh"

namespace nfk {

constexpr int BW_MAXDIM = 16;

struct BwdArgs {
  const float* x;
  const float* params;
  const float* gout;
  const float* gld;      // may be null
  float* gx;
  float* gparams;
  long long N;
  int size, dim, n_mask, n_unm, d, F_t, P, inverse;
  int mask[BW_MAXDIM];
  int unm[BW_MAXDIM];
  RqsConsts c;
};

template <int KT>
__global__ void __launch_bounds__(128) rqs_coupling_bwd_kernel(const __grid_constant__ BwdArgs a) {
  __shared__ int s_mask[BW_MAXDIM], s_unm[BW_MAXDIM];
  if (threadIdx.x < BW_MAXDIM) {
    s_mask[threadIdx.x] = a.mask[threadIdx.x];
    s_unm[threadIdx.x] = a.unm[threadIdx.x];
  }
  __syncthreads();
  constexpr int KK = KT ? KT : KMAX;
  const int K = KT ? KT : a.c.K;
  const RqsConsts& c = a.c;
  const long long total = a.N * a.F_t;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long n = e / a.F_t;
    const int f = (int)(e - n * a.F_t);
    const int s = f / a.n_unm, j = f - s * a.n_unm;
    const long long xi = n * a.d + s * a.dim + s_unm[j];
    const long long oi = n * a.d + s * a.dim + a.n_mask + j;
    const float* p = a.params + e * a.P;
    float* gp = a.gparams + e * a.P;
    const float x = a.x[xi];
    const float gy = a.gout[oi];
    const float gl = a.gld ? a.gld[n] : 0.f;
    // conditioning columns: out[:, s, i] = x[:, s, mask[i]]  (flows.py:239)
    if (j == 0)
      for (int i = 0; i < a.n_mask; ++i) a.gx[n * a.d + s * a.dim + s_mask[i]] = a.gout[n * a.d + s * a.dim + i];

    float gxv;
    float gpl[3 * KK - 1];
    rqs_element_bwd<KT>(PtrParams{p}, x, gy, gl, a.inverse != 0, c, gxv, gpl);
    a.gx[xi] = gxv;
#pragma unroll
    for (int i = 0; i < 3 * KK - 1; ++i)
      if (i < 3 * K - 1) gp[i] = gpl[i];
  }
}

// backward of rqs_elementwise_kernel: one thread per scalar
template <int KT>
__global__ void __launch_bounds__(128)
rqs_elementwise_bwd_kernel(const float* __restrict__ inputs, const float* __restrict__ params,
                           const float* __restrict__ gout, const float* __restrict__ glad,
                           float* __restrict__ gin, float* __restrict__ gparams, long long M, int inverse,
                           RqsConsts c) {
  constexpr int KK = KT ? KT : KMAX;
  const int K = KT ? KT : c.K;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M) return;
  const int P = 3 * K - 1;
  float gxv;
  float gpl[3 * KK - 1];
  rqs_element_bwd<KT>(PtrParams{params + e * P}, inputs[e], gout[e], glad ? glad[e] : 0.f, inverse != 0, c, gxv,
                      gpl);
  gin[e] = gxv;
#pragma unroll
  for (int i = 0; i < 3 * KK - 1; ++i)
    if (i < P) gparams[e * P + i] = gpl[i];
}

int fill_coupling_geometry_bwd(BwdArgs& a, int size, int dim, const int32_t* mask, int n_mask, int K) {
  NFK_REQUIRE(size > 0 && dim > 1 && dim <= BW_MAXDIM, "rqs_coupling_bwd: need size > 0 and 2 <= dim <= %d",
              BW_MAXDIM);
  NFK_REQUIRE(mask != nullptr && n_mask > 0 && n_mask < dim, "rqs_coupling_bwd: bad mask");
  NFK_REQUIRE(K >= 2 && K <= KMAX, "rqs_coupling_bwd: 2 <= K <= %d supported", KMAX);
  bool used[BW_MAXDIM] = {false};
  for (int i = 0; i < n_mask; ++i) {
    NFK_REQUIRE(mask[i] >= 0 && mask[i] < dim && !used[mask[i]], "rqs_coupling_bwd: bad mask column %d", mask[i]);
    used[mask[i]] = true;
    a.mask[i] = mask[i];
  }
  int nu = 0;
  for (int cidx = 0; cidx < dim; ++cidx)
    if (!used[cidx]) a.unm[nu++] = cidx;
  a.size = size;
  a.dim = dim;
  a.n_mask = n_mask;
  a.n_unm = nu;
  a.d = size * dim;
  a.F_t = size * nu;
  a.P = 3 * K - 1;
  return NFK_OK;
}

RqsConsts make_rqs_consts(int K, float B);

}  // namespace nfk

using namespace nfk;

extern "C" int nfk_rqs_coupling_bwd(const float* x, const float* params, const float* grad_out,
                                    const float* grad_logdet, float* grad_x, float* grad_params,
                                    int64_t N, int size, int dim, const int32_t* mask, int n_mask,
                                    int K, float B, int inverse, void* stream) {
  NFK_REQUIRE(N >= 0, "rqs_coupling_bwd: negative batch");
  BwdArgs a{};
  if (int rc = fill_coupling_geometry_bwd(a, size, dim, mask, n_mask, K)) return rc;
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && params && grad_out && grad_x && grad_params, "rqs_coupling_bwd: null device pointer");
  a.x = x;
  a.params = params;
  a.gout = grad_out;
  a.gld = grad_logdet;
  a.gx = grad_x;
  a.gparams = grad_params;
  a.N = N;
  a.inverse = inverse;
  a.c = make_rqs_consts(K, B);
  const long long total = N * a.F_t;
  long long grid = (total + 127) / 128;
  const long long cap = (long long)sm_count() * 32;
  if (grid > cap) grid = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 8)
    rqs_coupling_bwd_kernel<8><<<(unsigned)grid, 128, 0, st>>>(a);
  else
    rqs_coupling_bwd_kernel<0><<<(unsigned)grid, 128, 0, st>>>(a);
  count_launch();
  return check_launch("rqs_coupling_bwd");
}

extern "C" int nfk_rqs_elementwise_bwd(const float* inputs, const float* params, const float* grad_out,
                                       const float* grad_lad, float* grad_in, float* grad_params, int64_t M,
                                       int K, float B, int inverse, void* stream) {
  NFK_REQUIRE(M >= 0, "rqs_elementwise_bwd: negative size");
  NFK_REQUIRE(K >= 2 && K <= KMAX, "rqs_elementwise_bwd: 2 <= K <= %d supported", KMAX);
  NFK_REQUIRE(B > 0.f, "rqs_elementwise_bwd: tail bound must be positive");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(inputs && params && grad_out && grad_in && grad_params, "rqs_elementwise_bwd: null device pointer");
  const RqsConsts c = make_rqs_consts(K, B);
  const unsigned grid = (unsigned)((M + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 8)
    rqs_elementwise_bwd_kernel<8><<<grid, 128, 0, st>>>(inputs, params, grad_out, grad_lad, grad_in, grad_params, M,
                                                        inverse, c);
  else
    rqs_elementwise_bwd_kernel<0><<<grid, 128, 0, st>>>(inputs, params, grad_out, grad_lad, grad_in, grad_params, M,
                                                        inverse, c);
  count_launch();
  return check_launch("rqs_elementwise_bwd");
}
