// Backward of the RQS coupling transform: replaces autograd through NSF_CL's op chain after the
// conditioner (reference nf/flows.py:232-239 / :246-253 + nf/utils.py:27-152).  Given
// grad_out [N, size*dim] and grad_logdet [N] it produces grad_x [N, size*dim] (direct path only;
// the path through the conditioner flows back through grad_params) and grad_params
// [N, F_t, 3K-1].  One thread per spline: the forward is recomputed (bin search on the EXACT knot
// chain so the bin equals the forward kernel's), then the adjoints of the rational-quadratic
// expression, of the cumsum/softmax/softmax chain (quirk Q1) and of the double softplus (Q2) are
// propagated by hand.
#include "rqs_math.cuh"

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

__device__ __forceinline__ float sigmoidf(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float softplus_t(float v) { return v > 20.f ? v : log1pf(expf(v)); }

// softmax chain of one side with the intermediates the backward needs:
// a = softmax(raw), W1 = 2B a, b = softmax(W1); knots from w~ = 1e-3 + (1 - 1e-3 K) b
template <int KT>
__device__ __forceinline__ void side_forward(const float* raw, float* a, float* b, float* knots,
                                             const RqsConsts& c) {
  constexpr int KK = KT ? KT : KMAX;
  const int K = KT ? KT : c.K;
  float m = raw[0];
#pragma unroll
  for (int j = 1; j < KK; ++j)
    if (j < K) m = fmaxf(m, raw[j]);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < KK; ++j)
    if (j < K) {
      a[j] = expf(raw[j] - m);
      s += a[j];
    }
  float m2 = 0.f;
#pragma unroll
  for (int j = 0; j < KK; ++j)
    if (j < K) {
      a[j] /= s;
      m2 = fmaxf(m2, c.twoB * a[j]);
    }
  float s2 = 0.f;
#pragma unroll
  for (int j = 0; j < KK; ++j)
    if (j < K) {
      b[j] = expf(c.twoB * a[j] - m2);
      s2 += b[j];
    }
  float run = 0.f;
  knots[0] = c.negB;
#pragma unroll
  for (int j = 0; j < KK; ++j)
    if (j < K) {
      b[j] /= s2;
      run += c.min_bin + c.one_m * b[j];
      knots[j + 1] = c.twoB * run + c.negB;
    }
  knots[K] = c.B;
}

// adjoint of the knot chain: g_lo / g_hi = dL/d knot_k, dL/d knot_{k+1} -> dL/d raw[0..K)
template <int KT>
__device__ __forceinline__ void side_backward(const float* a, const float* b, int k, float g_lo,
                                              float g_hi, float* graw, const RqsConsts& c) {
  constexpr int KK = KT ? KT : KMAX;
  const int K = KT ? KT : c.K;
  // knot_j = -B + 2B sum_{i<j} w~_i ; knot_0 and knot_K are pinned constants
  if (k == 0) g_lo = 0.f;
  if (k + 1 == K) g_hi = 0.f;
  float gb[KK];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < KK; ++i)
    if (i < K) {
      const float gw = c.twoB * ((i < k ? g_lo : 0.f) + (i <= k ? g_hi : 0.f));
      gb[i] = c.one_m * gw;
      dot += b[i] * gb[i];
    }
  float dot2 = 0.f;
#pragma unroll
  for (int i = 0; i < KK; ++i)
    if (i < K) {
      const float gW1 = b[i] * (gb[i] - dot);        // softmax #2
      gb[i] = c.twoB * gW1;                          // W1 = 2B a
      dot2 += a[i] * gb[i];
    }
#pragma unroll
  for (int i = 0; i < KK; ++i)
    if (i < K) graw[i] = a[i] * (gb[i] - dot2);      // softmax #1
}

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

    const bool inside = (x >= c.negB) && (x <= c.B);
    if (!inside) {                                   // identity tails (utils.py:42-43)
      a.gx[xi] = gy;
      for (int i = 0; i < a.P; ++i) gp[i] = 0.f;
      continue;
    }
    float rw[KK], rh[KK], aw[KK], bw[KK], ah[KK], bh[KK], cw[KK + 1], ch[KK + 1];
#pragma unroll
    for (int i = 0; i < KK; ++i)
      if (i < K) {
        rw[i] = p[i];
        rh[i] = p[K + i];
      }
    side_forward<KT>(rw, aw, bw, cw, c);
    side_forward<KT>(rh, ah, bh, ch, c);
    // bin index from the EXACT chain of the searched side = the forward kernels' bin
    int k = 0;
    {
      float ex[KK + 1];
#pragma unroll
      for (int i = 0; i < KK; ++i)
        if (i < K) ex[i] = a.inverse ? rh[i] : rw[i];
      knot_chain<true, KT, true>(ex, c);
#pragma unroll
      for (int i = 1; i < KK; ++i)
        if (i < K) k += (x >= ex[i]) ? 1 : 0;
      k += (x >= c.Bnudge) ? 1 : 0;
      k = min(k, K - 1);
    }
    float c0 = cw[0], c1 = cw[1], e0 = ch[0], e1 = ch[1];
#pragma unroll
    for (int i = 1; i < KK; ++i)
      if (i < K && k == i) {
        c0 = cw[i];
        c1 = cw[i + 1];
        e0 = ch[i];
        e1 = ch[i + 1];
      }
    // derivatives: D2 = [c, softplus(Dr), c]; d = 1e-3 + softplus(D2)
    const int i0 = max(k - 1, 0), i1 = min(k, K - 2);
    const float dr0 = p[2 * K + i0], dr1 = p[2 * K + i1];
    const float D20 = (k == 0) ? c.edge_c : softplus_t(dr0);
    const float D21 = (k == K - 1) ? c.edge_c : softplus_t(dr1);
    const float d0 = c.min_d + softplus_t(D20), d1 = c.min_d + softplus_t(D21);

    const float wk = c1 - c0, hk = e1 - e0, delta = hk / wk;
    const float sS = d0 + d1 - 2.f * delta;
    float gxv, gc0, gc1, ge0, ge1, gd0, gd1;
    if (!a.inverse) {
      const float th = (x - c0) / wk, omt = 1.f - th, t = th * omt, th2 = th * th;
      const float A = delta * th2 + d0 * t;
      const float num = hk * A;
      const float den = delta + sS * t;
      const float Bq = d1 * th2 + 2.f * delta * t + d0 * omt * omt;
      const float dnum = delta * delta * Bq;
      // adjoints
      const float gnum = gy / den;
      float gden = -gy * num / (den * den) - 2.f * gl / den;
      ge0 = gy;
      const float gdnum = gl / dnum;
      float gdelta = gdnum * 2.f * delta * Bq;
      const float gBq = gdnum * delta * delta;
      gd1 = gBq * th2;
      float gth2 = gBq * d1;
      gdelta += gBq * 2.f * t;
      float gt = gBq * 2.f * delta;
      gd0 = gBq * omt * omt;
      float gomt = gBq * d0 * 2.f * omt;
      gdelta += gden;
      const float gs = gden * t;
      gt += gden * sS;
      gd0 += gs;
      gd1 += gs;
      gdelta -= 2.f * gs;
      float ghk = gnum * A;
      const float gA = gnum * hk;
      gdelta += gA * th2;
      gth2 += gA * delta;
      gd0 += gA * t;
      gt += gA * d0;
      float gth = gth2 * 2.f * th + gt * omt;
      gomt += gt * th;
      gth -= gomt;
      const float gu = gth / wk;
      float gwk = -gth * th / wk;
      gxv = gu;
      gc0 = -gu;
      ghk += gdelta / wk;
      gwk -= gdelta * delta / wk;
      gc1 = gwk;
      gc0 -= gwk;
      ge1 = ghk;
      ge0 -= ghk;
    } else {
      const float u = x - e0;
      const float us = u * sS;
      const float qa = us + hk * (delta - d0);
      const float qb = hk * d0 - us;
      const float qc = -delta * u;
      const float disc = fmaxf(qb * qb - 4.f * qa * qc, 0.f);
      const float sq = sqrtf(disc);
      const float Dn = -qb - sq;
      const float root = 2.f * qc / Dn;
      const float omr = 1.f - root, t = root * omr;
      const float den = delta + sS * t;
      const float Bq = d1 * root * root + 2.f * delta * t + d0 * omr * omr;
      const float dnum = delta * delta * Bq;
      // y = root*wk + c0 ; lad = -(log dnum - 2 log den)
      float groot = gy * wk;
      float gwk = gy * root;
      gc0 = gy;
      const float gdnum = -gl / dnum;
      const float gden = 2.f * gl / den;
      float gdelta = gdnum * 2.f * delta * Bq;
      const float gBq = gdnum * delta * delta;
      gd1 = gBq * root * root;
      const float gr2 = gBq * d1;
      gdelta += gBq * 2.f * t;
      float gt = gBq * 2.f * delta;
      gd0 = gBq * omr * omr;
      float gomr = gBq * d0 * 2.f * omr;
      gdelta += gden;
      float gs = gden * t;
      gt += gden * sS;
      groot += gr2 * 2.f * root + gt * omr;
      gomr += gt * root;
      groot -= gomr;
      float gqc = groot * 2.f / Dn;
      const float gDn = -groot * root / Dn;
      float gqb = -gDn;
      const float gsq = -gDn;
      const float gdisc = sq > 0.f ? gsq / (2.f * sq) : 0.f;
      gqb += gdisc * 2.f * qb;
      const float gqa = -4.f * qc * gdisc;
      gqc += -4.f * qa * gdisc;
      gdelta += -gqc * u;
      float gu = -gqc * delta;
      float ghk = gqb * d0;
      gd0 += gqb * hk;
      gu += -gqb * sS;
      gs += -gqb * u;
      gu += gqa * sS;
      gs += gqa * u;
      ghk += gqa * (delta - d0);
      gdelta += gqa * hk;
      gd0 -= gqa * hk;
      gd0 += gs;
      gd1 += gs;
      gdelta -= 2.f * gs;
      gxv = gu;
      ge0 = -gu;
      ghk += gdelta / wk;
      gwk -= gdelta * delta / wk;
      gc1 = gwk;
      gc0 -= gwk;
      ge1 = ghk;
      ge0 -= ghk;
    }
    a.gx[xi] = gxv;
    float graw[KK];
    side_backward<KT>(aw, bw, k, gc0, gc1, graw, c);
#pragma unroll
    for (int i = 0; i < KK; ++i)
      if (i < K) gp[i] = graw[i];
    side_backward<KT>(ah, bh, k, ge0, ge1, graw, c);
#pragma unroll
    for (int i = 0; i < KK; ++i)
      if (i < K) gp[K + i] = graw[i];
    // d_k = 1e-3 + softplus(D2_k), D2_k = softplus(Dr_{k-1}) for interior knots
    for (int i = 0; i < K - 1; ++i) gp[2 * K + i] = 0.f;
    if (k > 0) gp[2 * K + k - 1] += gd0 * sigmoidf(D20) * sigmoidf(dr0);
    if (k < K - 1) gp[2 * K + k] += gd1 * sigmoidf(D21) * sigmoidf(dr1);
  }
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
