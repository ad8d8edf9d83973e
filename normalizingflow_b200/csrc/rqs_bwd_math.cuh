// Per-element backward of the RQS coupling transform (adjoint of rqs_math.cuh's rqs_element with
// LAYER_NORM): shared by the stand-alone backward kernel (rqs_bwd.cu) and by the backward epilogue
// of the conditioner GEMM (gemm_ws.cu).  Replaces autograd through reference nf/flows.py:232-239 /
// :246-253 + nf/utils.py:27-152 for one transformed scalar.
#pragma once
#include "rqs_math.cuh"

namespace nfk {

__device__ __forceinline__ float sigmoidf(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float softplus_t(float v) { return v > 20.f ? v : log1pf(expf(v)); }

// FM = fast math (MUFU ex2 / lg2 / rcp, a few ulp): used where the conditioner is bf16 anyway
template <bool FM>
struct Bm {
  static __device__ __forceinline__ float exp(float v) { return FM ? ex2_approx(v * LOG2E) : expf(v); }
  static __device__ __forceinline__ float div(float a, float b) { return FM ? a * rcp_approx(b) : a / b; }
  static __device__ __forceinline__ float sigmoid(float v) {
    return FM ? rcp_approx(1.f + ex2_approx(-v * LOG2E)) : sigmoidf(v);
  }
  static __device__ __forceinline__ float softplus(float v) {
    if (!FM) return softplus_t(v);
    return v > 20.f ? v : LN2 * lg2_approx(1.f + ex2_approx(v * LOG2E));
  }
};

// softmax chain of one side with the intermediates the backward needs:
// a = softmax(raw), W1 = 2B a, b = softmax(W1); knots from w~ = 1e-3 + (1 - 1e-3 K) b
template <int KT, bool FM = false>
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
      a[j] = Bm<FM>::exp(raw[j] - m);
      s += a[j];
    }
  float m2 = 0.f;
  const float rs = FM ? rcp_approx(s) : 0.f;
#pragma unroll
  for (int j = 0; j < KK; ++j)
    if (j < K) {
      a[j] = FM ? a[j] * rs : a[j] / s;
      m2 = fmaxf(m2, c.twoB * a[j]);
    }
  float s2 = 0.f;
#pragma unroll
  for (int j = 0; j < KK; ++j)
    if (j < K) {
      b[j] = Bm<FM>::exp(c.twoB * a[j] - m2);
      s2 += b[j];
    }
  float run = 0.f;
  knots[0] = c.negB;
  const float rs2 = FM ? rcp_approx(s2) : 0.f;
#pragma unroll
  for (int j = 0; j < KK; ++j)
    if (j < K) {
      b[j] = FM ? b[j] * rs2 : b[j] / s2;
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

// raw parameters behind a pointer (stand-alone kernel)
struct PtrParams {
  const float* p;
  __device__ __forceinline__ float operator()(int i) const { return p[i]; }
  __device__ __forceinline__ float dyn(int base, int i) const { return p[base + i]; }
};

// Adjoint of the rational-quadratic segment itself (utils.py:98-152) given the bin's knots (c0, c1) x (e0, e1) and
// the derivatives d0, d1 at its ends: dL/dx (direct path) and the gradients of those six quantities.
template <bool FM>
__device__ __forceinline__ void rq_segment_bwd(bool inverse, float x, float c0, float c1, float e0, float e1, float d0,
                                               float d1, float gy, float gl, float& gxv, float& gc0, float& gc1,
                                               float& ge0, float& ge1, float& gd0, float& gd1) {
  const float wk = c1 - c0, hk = e1 - e0, delta = Bm<FM>::div(hk, wk);
  const float sS = d0 + d1 - 2.f * delta;
  if (!inverse) {
    const float th = Bm<FM>::div(x - c0, wk), omt = 1.f - th, t = th * omt, th2 = th * th;
    const float A = delta * th2 + d0 * t;
    const float num = hk * A;
    const float den = delta + sS * t;
    const float Bq = d1 * th2 + 2.f * delta * t + d0 * omt * omt;
    const float dnum = delta * delta * Bq;
    // adjoints
    const float gnum = Bm<FM>::div(gy, den);
    float gden = -Bm<FM>::div(gy * num, den * den) - Bm<FM>::div(2.f * gl, den);
    ge0 = gy;
    const float gdnum = Bm<FM>::div(gl, dnum);
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
    const float gu = Bm<FM>::div(gth, wk);
    float gwk = -Bm<FM>::div(gth * th, wk);
    gxv = gu;
    gc0 = -gu;
    ghk += Bm<FM>::div(gdelta, wk);
    gwk -= Bm<FM>::div(gdelta * delta, wk);
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
    const float root = Bm<FM>::div(2.f * qc, Dn);
    const float omr = 1.f - root, t = root * omr;
    const float den = delta + sS * t;
    const float Bq = d1 * root * root + 2.f * delta * t + d0 * omr * omr;
    const float dnum = delta * delta * Bq;
    // y = root*wk + c0 ; lad = -(log dnum - 2 log den)
    float groot = gy * wk;
    float gwk = gy * root;
    gc0 = gy;
    const float gdnum = -Bm<FM>::div(gl, dnum);
    const float gden = Bm<FM>::div(2.f * gl, den);
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
    float gqc = Bm<FM>::div(groot * 2.f, Dn);
    const float gDn = -Bm<FM>::div(groot * root, Dn);
    float gqb = -gDn;
    const float gsq = -gDn;
    const float gdisc = sq > 0.f ? Bm<FM>::div(gsq, 2.f * sq) : 0.f;
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
    ghk += Bm<FM>::div(gdelta, wk);
    gwk -= Bm<FM>::div(gdelta * delta, wk);
    gc1 = gwk;
    gc0 -= gwk;
    ge1 = ghk;
    ge0 -= ghk;
  }
}

// P: functor, P(i) = i-th of the 3K-1 raw conditioner outputs, P.dyn(base, i) = run-time index.
// In: x (transformed input), gy = dL/dy, gl = dL/dlogdet.  Out: gx_out = dL/dx (direct path),
// gp[0..3K-1) = dL/draw.
template <int KT, bool FM = false, class P>
__device__ __forceinline__ void rqs_element_bwd(const P& p, float x, float gy, float gl, bool inverse,
                                                const RqsConsts& c, float& gx_out, float* gp) {
  constexpr int KK = KT ? KT : KMAX;
  const int K = KT ? KT : c.K;
  const bool inside = (x >= c.negB) && (x <= c.B);
  if (!inside) {                                   // identity tails (utils.py:42-43)
    gx_out = gy;
#pragma unroll
    for (int i = 0; i < 3 * KK - 1; ++i)
      if (i < 3 * K - 1) gp[i] = 0.f;
    return;
  }
  float rw[KK], rh[KK], aw[KK], bw[KK], ah[KK], bh[KK], cw[KK + 1], ch[KK + 1];
#pragma unroll
  for (int i = 0; i < KK; ++i)
    if (i < K) {
      rw[i] = p(i);
      rh[i] = p(K + i);
    }
  side_forward<KT, FM>(rw, aw, bw, cw, c);
  side_forward<KT, FM>(rh, ah, bh, ch, c);
  // bin index from the EXACT chain of the searched side = the forward kernels' bin
  int k = 0;
  if (FM) {
    // fast path: search the knots just computed (a bin may differ from the forward kernel's when x
    // is within a few ulp of a knot; the transform is C1 there)
#pragma unroll
    for (int i = 1; i < KK; ++i)
      if (i < K) k += (x >= (inverse ? ch[i] : cw[i])) ? 1 : 0;
    k += (x >= c.Bnudge) ? 1 : 0;
    k = min(k, K - 1);
  } else {
    float ex[KK + 1];
#pragma unroll
    for (int i = 0; i < KK; ++i)
      if (i < K) ex[i] = inverse ? rh[i] : rw[i];
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
  const float dr0 = p.dyn(2 * K, i0), dr1 = p.dyn(2 * K, i1);
  const float D20 = (k == 0) ? c.edge_c : Bm<FM>::softplus(dr0);
  const float D21 = (k == K - 1) ? c.edge_c : Bm<FM>::softplus(dr1);
  const float d0 = c.min_d + Bm<FM>::softplus(D20), d1 = c.min_d + Bm<FM>::softplus(D21);

  float gxv, gc0, gc1, ge0, ge1, gd0, gd1;
  rq_segment_bwd<FM>(inverse, x, c0, c1, e0, e1, d0, d1, gy, gl, gxv, gc0, gc1, ge0, ge1, gd0, gd1);
  gx_out = gxv;
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
  const float gD0 = (k > 0) ? gd0 * Bm<FM>::sigmoid(D20) * Bm<FM>::sigmoid(dr0) : 0.f;
  const float gD1 = (k < K - 1) ? gd1 * Bm<FM>::sigmoid(D21) * Bm<FM>::sigmoid(dr1) : 0.f;
#pragma unroll
  for (int i = 0; i < KK - 1; ++i)
    if (i < K - 1) gp[2 * K + i] = (i == k - 1 ? gD0 : 0.f) + (i == k ? gD1 : 0.f);
}

// ---- K = 8, fast-math adjoint with the two sides PACKED (width side in the low, height side in the high half of fp32x2
// registers: FFMA2 / FADD2 / FMUL2 do both softmax chains, forward and backward, in one instruction stream), organised
// for a small live set (the one-launch layer backward, nsf_fused_bwd.cu, runs it at 96 registers per thread).  The raw
// parameters arrive interleaved: v[2j] = width logit j, v[2j+1] = height logit j, v[16+i] = derivative logit i (the
// one-launch layer backward permutes the rows of its W3 image accordingly), `b` the bias in the same order.
// Differences from rqs_element_bwd<8, true>, all inside the fast-math class: the shift of the second softmax is
// 2B / sum (= 2B max_j a_j, the argmax term being exp(0) / sum) instead of a running maximum, knots come from the
// prefix sums of the normalised bins (knot_j = q0 cum_j + kstep j - B), and the dot product of the second softmax's
// backward is taken from the two selected knots (sum_{i<k} b_i and b_k) instead of a sum over the bins.
struct PairRegParams {
  const uint32_t* v;
  const float* b;
  __device__ __forceinline__ F2 wh(int j) const {
    F2 bias;
    bias.u = *reinterpret_cast<const unsigned long long*>(b + 2 * j);
    return add2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bias);
  }
  __device__ __forceinline__ float dyn(int i) const {
    uint32_t r = v[16];
#pragma unroll
    for (int j = 1; j < 7; ++j)
      if (i == j) r = v[16 + j];
    return __uint_as_float(r) + b[16 + i];
  }
};

__device__ __forceinline__ F2 sel2(bool p, F2 a, F2 b) {
  F2 r;
  r.u = p ? a.u : b.u;
  return r;
}

// Outputs: gx_out = dL/dx (direct path), gwh[i] = (dL/d width logit i, dL/d height logit i), and the derivative-logit
// gradients as (bin, gD0, gD1): dL/d derivative logit bin-1 = gD0, dL/d derivative logit bin = gD1, all others 0.
template <bool INV>
__device__ __forceinline__ void rqs_element_bwd8_packed(const PairRegParams& p, float x, float gy, float gl,
                                                        const RqsConsts& c, float& gx_out, F2* gwh, int& bin,
                                                        float& gD0_out, float& gD1_out) {
  constexpr int K = 8;
  const F2 zero2 = pk2(0.f, 0.f);
  if (!((x >= c.negB) && (x <= c.B))) {            // identity tails (utils.py:42-43)
    gx_out = gy;
#pragma unroll
    for (int i = 0; i < K; ++i) gwh[i] = zero2;
    bin = 0;
    gD0_out = 0.f;
    gD1_out = 0.f;
    return;
  }
  F2 A[K], Bv[K];
  // ---- both double-softmax chains
  {
    F2 R[K];
    float mw, mh;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      R[j] = p.wh(j);
      float lo, hi;
      unpk2(R[j], lo, hi);
      mw = j ? fmaxf(mw, lo) : lo;
      mh = j ? fmaxf(mh, hi) : hi;
    }
    const F2 l2e = pk2(LOG2E, LOG2E);
    const F2 mm = mul2(pk2(-mw, -mh), l2e);
    F2 s = zero2;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      A[j] = ex2_2(fma2(R[j], l2e, mm));
      s = add2(s, A[j]);
    }
    float sw, sh;
    unpk2(s, sw, sh);
    const F2 rs = pk2(rcp_approx(sw), rcp_approx(sh));
    const F2 g0 = pk2(c.g0, c.g0);
    const F2 nm2 = mul2(pk2(-c.g0, -c.g0), rs);            // -2B log2(e) max_j a_j
    F2 s2 = zero2;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      A[j] = mul2(A[j], rs);
      Bv[j] = ex2_2(fma2(A[j], g0, nm2));
      s2 = add2(s2, Bv[j]);
    }
    unpk2(s2, sw, sh);
    const F2 rs2 = pk2(rcp_approx(sw), rcp_approx(sh));
#pragma unroll
    for (int j = 0; j < K; ++j) Bv[j] = mul2(Bv[j], rs2);
  }
  // ---- knots 1..7 of both sides, the bin, its two knots on each side
  int k = 0;
  F2 k0 = pk2(c.negB, c.negB), k1;
  {
    F2 kn[K];                                            // kn[j] = knot j (j = 1..7)
    const F2 q0 = pk2(c.q0, c.q0);
    F2 cum = zero2;
#pragma unroll
    for (int j = 1; j < K; ++j) {
      cum = add2(cum, Bv[j - 1]);
      const float off = fmaf(c.kstep, (float)j, c.negB);
      kn[j] = fma2(cum, q0, pk2(off, off));
      float lo, hi;
      unpk2(kn[j], lo, hi);
      k += (x >= (INV ? hi : lo)) ? 1 : 0;
    }
    k += (x >= c.Bnudge) ? 1 : 0;
    k = min(k, K - 1);
    k1 = kn[1];
#pragma unroll
    for (int i = 1; i < K; ++i)
      if (k == i) {
        k0 = kn[i];
        k1 = (i + 1 < K) ? kn[i + 1] : pk2(c.B, c.B);
      }
  }
  float c0, e0, c1, e1;
  unpk2(k0, c0, e0);
  unpk2(k1, c1, e1);
  // ---- derivatives at the bin's ends: d = 1e-3 + softplus(D2), D2 = softplus(raw) inside, the padded constant at the
  // boundary; each softplus shares its exponential with the sigmoid the backward needs
  const float dr0 = p.dyn(max(k - 1, 0)), dr1 = p.dyn(min(k, K - 2));
  auto sp_sg = [](float v, float& sp, float& sg) {      // softplus(v), sigmoid(v) from one ex2
    const float e = ex2_approx(fminf(v, 30.f) * LOG2E), r = rcp_approx(1.f + e);
    sp = v > 20.f ? v : LN2 * lg2_approx(1.f + e);
    sg = e * r;
  };
  float D20, D21, s_dr0, s_dr1, sp0, sp1, s_D0, s_D1;
  sp_sg(dr0, D20, s_dr0);
  sp_sg(dr1, D21, s_dr1);
  if (k == 0) D20 = c.edge_c;
  if (k == K - 1) D21 = c.edge_c;
  sp_sg(D20, sp0, s_D0);
  sp_sg(D21, sp1, s_D1);
  const float d0 = c.min_d + sp0, d1 = c.min_d + sp1;
  float gxv, gc0, gc1, ge0, ge1, gd0, gd1;
  rq_segment_bwd<true>(INV, x, c0, c1, e0, e1, d0, d1, gy, gl, gxv, gc0, gc1, ge0, ge1, gd0, gd1);
  gx_out = gxv;
  // ---- adjoint of both knot chains: knot_0 and knot_K are pinned constants
  {
    const bool first = (k == 0), last = (k == K - 1);
    const F2 glo = pk2(first ? 0.f : gc0, first ? 0.f : ge0);
    const F2 ghi = pk2(last ? 0.f : gc1, last ? 0.f : ge1);
    const F2 q0 = pk2(c.q0, c.q0);
    const F2 Y = mul2(q0, ghi);                          // dL/d b_i for i == k
    const F2 X = fma2(q0, glo, Y);                       //            for i <  k   (0 for i > k)
    // dot = sum_i b_i dL/db_i = X sum_{i<k} b_i + Y b_k, the sums read off the selected knots
    const float rq0 = rcp_approx(c.q0);
    const float offk = fmaf(c.kstep, (float)k, c.negB);
    const F2 cumk = mul2(add2(k0, pk2(-offk, -offk)), pk2(rq0, rq0));
    const F2 bk = mul2(add2(add2(k1, mul2(k0, pk2(-1.f, -1.f))), pk2(-c.kstep, -c.kstep)), pk2(rq0, rq0));
    const F2 dot = fma2(X, cumk, mul2(Y, bk));
    const F2 twoB = pk2(c.twoB, c.twoB);
    const F2 ndot = mul2(dot, pk2(-1.f, -1.f));
    const F2 XD = mul2(twoB, add2(X, ndot)), YD = mul2(twoB, add2(Y, ndot)), ZD = mul2(twoB, ndot);
    F2 dot2 = zero2;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      F2 t = mul2(Bv[i], ZD);                            // dL/d a_i   (second softmax, W1 = 2B a)
      if (i < k) t = mul2(Bv[i], XD);
      if (i == k) t = mul2(Bv[i], YD);
      gwh[i] = t;
      dot2 = fma2(A[i], t, dot2);
    }
    const F2 ndot2 = mul2(dot2, pk2(-1.f, -1.f));
#pragma unroll
    for (int i = 0; i < K; ++i) gwh[i] = mul2(A[i], add2(gwh[i], ndot2));      // first softmax
  }
  bin = k;
  gD0_out = (k > 0) ? gd0 * s_D0 * s_dr0 : 0.f;
  gD1_out = (k < K - 1) ? gd1 * s_D1 * s_dr1 : 0.f;
}

}  // namespace nfk
