// RQS coupling transform (K1): the part of NSF_CL.forward/inverse after the conditioner
// (reference nf/flows.py:232-239, :246-253) fused with unconstrained_RQS / RQS / searchsorted
// (nf/utils.py:20-152) into one pass over HBM.
//
// Data movement (B200): a persistent CTA walks row tiles.  The spline parameters of a tile
// (R rows x F_t features x (3K-1) fp32, one contiguous block of the [N, F_t, 3K-1] tensor) and
// the tile's activations arrive in shared memory through 1-D TMA bulk copies
// (cp.async.bulk, SASS UBLKCP) into a multi-stage mbarrier ring, so HBM reads are full-line
// and asynchronous; each thread then reads the 3K-1 words of ITS element at word stride 3K-1
// (odd => bank-conflict free), does bin search + spline + log-det entirely in registers, the
// per-row log-det is a warp-shuffle reduction, and the output tile (reference column order,
// quirk Q5) leaves through a TMA bulk store.
//
// Algorithmic HBM bytes per row: F_t*(3K-1)*4 + 2*d*4 + 4 (+4 when accumulating log-det).
#include "rqs_math.cuh"

namespace nfk {

constexpr int MAXDIM = 16;
constexpr int MAX_THREADS = 640;

struct CouplingArgs {
  const float* x;
  const float* params;
  float* out;
  float* logdet;
  int8_t* bins;
  long long n_tiles;   // tiled kernel: number of full tiles
  long long row0;      // rows kernel: first row
  int size, dim, n_mask, n_unm, d, F_t, P, R, stages, accumulate;
  int scratch;         // streaming kernel: per-warp log-det ring, a power of two >= F_t + 32 (floats)
  int mask[MAXDIM];
  int unm[MAXDIM];
  RqsConsts c;
};

struct SmemPtr {
  const float* p;
  __device__ __forceinline__ float operator()(int i) const { return p[i]; }
  __device__ __forceinline__ float dyn(int base, int i) const { return p[base + i]; }
};

// Generic geometry: one element per thread per tile (the launcher sizes the CTA to the tile), so
// every index a thread needs is loop-invariant; activations and outputs go through shared-memory
// tiles moved by TMA bulk copies.
template <int MODE, int KT, bool INVERSE>
__global__ void __launch_bounds__(MAX_THREADS)
rqs_coupling_tiled(const __grid_constant__ CouplingArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_mask[MAXDIM], s_unm[MAXDIM];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const int P = a.P, R = a.R, d = a.d, F_t = a.F_t, S = a.stages;
  const int dim = a.dim, n_mask = a.n_mask, n_unm = a.n_unm;
  const int n_el = R * F_t;
  const int ptile = n_el * P;      // floats; R % 4 == 0 keeps every tile 16-byte aligned
  const int xtile = R * d;

  float* ps = reinterpret_cast<float*>(smem_raw);
  float* xs = ps + (size_t)S * ptile;
  float* outs = xs + (size_t)S * xtile;
  float* lads = outs + 2 * xtile;
  uint64_t* full = reinterpret_cast<uint64_t*>(lads + 2 * n_el);

  if (tid < MAXDIM) {
    s_mask[tid] = a.mask[tid];
    s_unm[tid] = a.unm[tid];
  }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const unsigned first = blockIdx.x, stride = gridDim.x;
  const unsigned n_tiles = (unsigned)a.n_tiles;
  const unsigned my_tiles = (n_tiles > first) ? (n_tiles - first + stride - 1) / stride : 0;
  const uint32_t stage_bytes = (uint32_t)(ptile + xtile) * 4u;

  // loop-invariant element geometry
  const bool has_el = tid < n_el;
  int xoff = 0, ooff = 0, r_el = 0;
  if (has_el) {
    r_el = tid / F_t;
    const int f = tid - r_el * F_t;
    const int s = f / n_unm, j = f - s * n_unm;
    xoff = r_el * d + s * dim + s_unm[j];
    ooff = r_el * d + s * dim + n_mask + j;
  }
  const int per_row = a.size * n_mask;
  const int n_copy = R * per_row;

  auto issue = [&](unsigned it, int stage) {
    const size_t tile = first + (size_t)it * stride;
    mbar_expect_tx(&full[stage], stage_bytes);
    bulk_g2s(ps + (size_t)stage * ptile, a.params + tile * ptile, (uint32_t)ptile * 4u, &full[stage]);
    bulk_g2s(xs + (size_t)stage * xtile, a.x + tile * xtile, (uint32_t)xtile * 4u, &full[stage]);
  };
  if (tid == 0)
    for (int it = 0; it < S && (unsigned)it < my_tiles; ++it) issue(it, it);

  int stage = 0;
  uint32_t phase = 0;
  for (unsigned it = 0; it < my_tiles; ++it) {
    const size_t tile = first + (size_t)it * stride;
    const int buf = (int)(it & 1);
    const float* pst = ps + (size_t)stage * ptile;
    const float* xst = xs + (size_t)stage * xtile;
    float* ob = outs + buf * xtile;
    float* lb = lads + buf * n_el;
    const size_t row_base = tile * R;

    mbar_wait(&full[stage], phase);
    if (has_el) {
      const RqsOut o = rqs_element<MODE, KT, INVERSE, true, true>(SmemPtr{pst + (size_t)tid * P}, xst[xoff], a.c);
      ob[ooff] = o.y;
      lb[tid] = o.lad;
      if (a.bins) a.bins[row_base * F_t + tid] = (int8_t)o.bin;
    }
    // conditioning columns move to the front of each dim-group (flows.py:239, quirk Q5)
    for (int i = tid; i < n_copy; i += nthr) {
      const int r = i / per_row;
      const int rem = i - r * per_row;
      const int s = rem / n_mask, mi = rem - s * n_mask;
      ob[r * d + s * dim + mi] = xst[r * d + s * dim + s_mask[mi]];
    }
    fence_proxy_async();
    if (tid == 0) bulk_wait_read<0>();   // store of tile it-1 has drained outs[buf^1]
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(a.out + tile * xtile, ob, (uint32_t)xtile * 4u);
      bulk_commit();
      if (it + S < my_tiles) issue(it + S, stage);
    }
    for (int r = warp; r < R; r += nwarps) {                           // flows.py:238
      float t = 0.f;
      for (int f = lane; f < F_t; f += 32) t += lb[r * F_t + f];
      t = warp_sum(t);
      if (lane == 0) {
        float* ldp = a.logdet + row_base + r;
        *ldp = a.accumulate ? *ldp + t : t;
      }
    }
    if (++stage == S) {
      stage = 0;
      phase ^= 1;
    }
  }
  if (tid == 0) bulk_wait_all<0>();
}

// Generic geometry, K = 8, second generation: every WARP is its own pipeline, there is no CTA-wide barrier.
// A warp owns groups of G consecutive rows, G = 32 / gcd(F_t, 32), so a group is a whole number of 32-element
// blocks whose (3K-1) x 32 parameters are one contiguous, 16-byte aligned span: lane 0 streams them through the
// warp's private TMA bulk-copy ring (S stages, S - 1 blocks ahead, across group boundaries), lane l evaluates element
// 32 k + l of the group from shared memory at the conflict-free word stride 3K-1.  Activations are gathered straight
// from global memory (one block ahead in a register) and outputs stored straight to it: a block's 32 elements are
// consecutive features of at most two rows, so the 4-byte accesses of a warp fall into a handful of sectors that L1 / L2
// merge.  Per-element log-dets go through a small ring (F_t + 32 floats, rounded up to a power of two) and each row is
// summed, as soon as its last element is in, with exactly the association of the tiled kernel (lane-strided partial
// sums + butterfly), so the result does not depend on where a row sits in the batch.  64 registers and 6.5 KB of shared
// memory per warp: 32 warps per SM.
// What it fixes (profiles/ncu_tiled_r02.json): the tiled kernel has one barrier and one single-thread TMA issue per
// 128-304 elements and 31 % of the warp slots occupied; barrier stalls dominate it.
constexpr int STREAM_WARPS = 4;
constexpr int MAXCOND = 3;        // conditioning columns per dim-group the streaming kernel moves in registers

template <int MODE, int KT, bool INVERSE>
__global__ void __launch_bounds__(STREAM_WARPS * 32, 8)
rqs_coupling_stream(const __grid_constant__ CouplingArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.P, S = a.stages, G = a.R, F_t = a.F_t, d = a.d;
  const int dim = a.dim, n_mask = a.n_mask, n_unm = a.n_unm;
  const int n_el = G * F_t;                 // elements of a group: a multiple of 32
  const int n_blk = n_el >> 5;
  const uint32_t blk_floats = 32u * (uint32_t)P;
  const int cap = a.scratch, cmask = cap - 1;
  const size_t per_warp = ((size_t)S * blk_floats + (size_t)cap) * 4 + 64;
  // feature -> column tables (input column of the transformed scalar, its output column), shared by the CTA: no integer
  // division in the element loop
  short* xcol = reinterpret_cast<short*>(smem_raw + (size_t)STREAM_WARPS * per_warp);
  short* ocol = xcol + F_t;
  short* gcol = ocol + F_t;                                // first column of the feature's dim-group if it is the group's
  //                                                          first transformed feature (that lane also moves the group's
  //                                                          conditioning values), else -1
  float* ring = reinterpret_cast<float*>(smem_raw + (size_t)warp * per_warp);
  float* lads = ring + (size_t)S * blk_floats;            // ring of per-element log-dets: a row is summed as soon as it is complete
  uint64_t* full = reinterpret_cast<uint64_t*>(lads + cap);
  for (int f = tid; f < F_t; f += STREAM_WARPS * 32) {
    const int sgrp = f / n_unm, j = f - sgrp * n_unm;
    xcol[f] = (short)(sgrp * dim + a.unm[j]);
    ocol[f] = (short)(sgrp * dim + n_mask + j);
    gcol[f] = (short)(j == 0 ? sgrp * dim : -1);
  }
  if (lane == 0) {
    for (int st = 0; st < S; ++st) mbar_init(&full[st], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const long long n_groups = a.n_tiles;
  const long long wstride = (long long)gridDim.x * STREAM_WARPS;
  const long long g0 = (long long)blockIdx.x * STREAM_WARPS + warp;
  // producer cursor (lane 0): next block to request
  long long pg = g0;
  int pk = 0, pstage = 0;
  auto issue = [&]() {
    if (pg < n_groups) {
      mbar_expect_tx(&full[pstage], blk_floats * 4u);
      bulk_g2s(ring + (size_t)pstage * blk_floats, a.params + ((size_t)pg * n_el + (size_t)pk * 32) * P, blk_floats * 4u,
               &full[pstage]);
      if (++pk == n_blk) {
        pk = 0;
        pg += wstride;
      }
      if (++pstage == S) pstage = 0;
    }
  };
  if (lane == 0)
    for (int i = 0; i < S - 1; ++i) issue();

  // (row, feature) of this lane's element inside its group, advanced by 32 elements per block
  const int r_first = lane / F_t, f_first = lane - r_first * F_t;
  auto advance = [&](int& r, int& f) {
    f += 32;
    while (f >= F_t) {
      f -= F_t;
      ++r;
    }
  };
  int stage = 0;
  uint32_t phase = 0;
  float xn = 0.f;
  if (g0 < n_groups) xn = __ldg(a.x + (size_t)g0 * G * d + r_first * d + xcol[f_first]);
  for (long long g = g0; g < n_groups; g += wstride) {
    const size_t row0 = (size_t)g * G;
    const float* xg = a.x + row0 * d;
    float* og = a.out + row0 * d;
    int r = r_first, f = f_first;                                // this block's element
    int rn = r_first, fn = f_first;                              // the next block's
    int r_done = 0;                                              // rows of this group already summed
    for (int k = 0; k < n_blk; ++k) {
      const int e = (k << 5) + lane;
      const float xin = xn;
      // next block's activation: of this group, or the first block of the warp's next group
      if (k + 1 < n_blk) {
        advance(rn, fn);
        xn = __ldg(xg + rn * d + xcol[fn]);
      } else if (g + wstride < n_groups) {
        xn = __ldg(a.x + (size_t)(g + wstride) * G * d + r_first * d + xcol[f_first]);
      }
      // conditioning values of the dim-group this element opens (flows.py:239, quirk Q5: they move to the front of the
      // group): stored with the transformed value below, so every 32-byte sector of the output row is completed by
      // neighbouring lanes of one block instead of being revisited at the end of the group
      const int gc = gcol[f];
      float cv[MAXCOND];
      if (gc >= 0) {
#pragma unroll
        for (int mi = 0; mi < MAXCOND; ++mi)
          if (mi < n_mask) cv[mi] = __ldg(xg + r * d + gc + a.mask[mi]);
      }
      if (lane == 0) issue();                                    // keeps S - 1 blocks in flight
      mbar_wait(&full[stage], phase);
      const RqsOut o = rqs_element<MODE, KT, INVERSE, true, true>(SmemPtr{ring + (size_t)stage * blk_floats + (size_t)lane * P},
                                                                 xin, a.c);
      og[r * d + ocol[f]] = o.y;
      if (gc >= 0) {
#pragma unroll
        for (int mi = 0; mi < MAXCOND; ++mi)
          if (mi < n_mask) og[r * d + gc + mi] = cv[mi];
      }
      lads[e & cmask] = o.lad;
      if (a.bins) a.bins[row0 * F_t + e] = (int8_t)o.bin;
      r = rn;
      f = fn;
      __syncwarp();                                              // every lane is done with this stage's parameters
      if (++stage == S) {
        stage = 0;
        phase ^= 1;
      }
      // rows completed by this block: flows.py:238 with the tiled kernel's association (lane-strided partial sums over the
      // row's features, then the butterfly), so a row's log-det does not depend on its position in the batch
      while ((r_done + 1) * F_t <= ((k + 1) << 5)) {
        float t = 0.f;
        for (int ff = lane; ff < F_t; ff += 32) t += lads[(r_done * F_t + ff) & cmask];
        t = warp_sum(t);
        if (lane == 0) {
          float* ldp = a.logdet + row0 + r_done;
          *ldp = a.accumulate ? *ldp + t : t;
        }
        ++r_done;
      }
    }
    __syncwarp();                                                // the log-det ring is free for the next group
  }
}

// Specialised geometry of the headline workload (size = 32, dim = 2, one masked column): one warp
// owns one row, lane s owns the column pair (2s, 2s+1), so activations are one coalesced 8-byte
// load/store per thread straight from/to HBM (prefetched one tile ahead in registers), only the
// spline parameters go through the TMA ring, and the row log-det is one warp-shuffle reduction.
constexpr int PAIR_THREADS = 256;

template <int MODE, int KT, bool INVERSE>
__global__ void __launch_bounds__(PAIR_THREADS, 4)
rqs_coupling_pairs(const __grid_constant__ CouplingArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int T = PAIR_THREADS;
  const int tid = threadIdx.x, lane = tid & 31;
  const int P = a.P, S = a.stages;
  const int ptile = T * P;                       // floats per stage
  float* ps = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(ps + (size_t)S * ptile);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const unsigned first = blockIdx.x, stride = gridDim.x;
  const unsigned n_tiles = (unsigned)a.n_tiles;
  const unsigned my_tiles = (n_tiles > first) ? (n_tiles - first + stride - 1) / stride : 0;
  const bool cond_first = (a.mask[0] == 0);      // conditioning column is column 0 of the pair

  auto issue = [&](unsigned it, int stage) {
    const size_t tile = first + (size_t)it * stride;
    mbar_expect_tx(&full[stage], (uint32_t)ptile * 4u);
    bulk_g2s(ps + (size_t)stage * ptile, a.params + tile * ptile, (uint32_t)ptile * 4u, &full[stage]);
  };
  if (tid == 0)
    for (int it = 0; it < S && (unsigned)it < my_tiles; ++it) issue(it, it);

  const float2* x2 = reinterpret_cast<const float2*>(a.x);
  float2* o2 = reinterpret_cast<float2*>(a.out);
  float2 xn = make_float2(0.f, 0.f);
  if (my_tiles) xn = __ldcs(x2 + (size_t)first * T + tid);

  const float* pth = ps + (size_t)tid * P;
  int stage = 0;
  uint32_t phase = 0;
  for (unsigned it = 0; it < my_tiles; ++it) {
    const size_t tile = first + (size_t)it * stride;
    const float2 xc = xn;
    if (it + 1 < my_tiles) xn = __ldcs(x2 + (tile + stride) * T + tid);
    mbar_wait(&full[stage], phase);
    const RqsOut o = rqs_element<MODE, KT, INVERSE, true, true>(SmemPtr{pth + (size_t)stage * ptile},
                                                          cond_first ? xc.y : xc.x, a.c);
    // out pair = (conditioning value, transformed value)  (flows.py:239, quirk Q5)
    __stcs(o2 + tile * T + tid, make_float2(cond_first ? xc.x : xc.y, o.y));
    if (a.bins) a.bins[tile * T + tid] = (int8_t)o.bin;
    const float t = warp_sum(o.lad);                                   // flows.py:238
    if (lane == 0) {
      float* ldp = a.logdet + tile * (T / 32) + (tid >> 5);
      *ldp = a.accumulate ? *ldp + t : t;
    }
    __syncthreads();                     // every thread is done with this stage's parameters
    if (tid == 0 && it + S < my_tiles) issue(it + S, stage);
    if (++stage == S) {
      stage = 0;
      phase ^= 1;
    }
  }
}

struct GmemPtr {
  const float* p;
  __device__ __forceinline__ float operator()(int i) const { return __ldg(p + i); }
  __device__ __forceinline__ float dyn(int base, int i) const { return __ldg(p + base + i); }
};

// One CTA per row, plain global accesses: tail rows (N % R), tiny batches and shapes whose
// tiles do not fit shared memory.
template <int MODE, int KT, bool INVERSE>
__global__ void __launch_bounds__(128) rqs_coupling_rows(const __grid_constant__ CouplingArgs a) {
  __shared__ int s_mask[MAXDIM], s_unm[MAXDIM];
  __shared__ float red[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < MAXDIM) {
    s_mask[tid] = a.mask[tid];
    s_unm[tid] = a.unm[tid];
  }
  __syncthreads();
  const long long row = a.row0 + blockIdx.x;
  const float* xr = a.x + row * a.d;
  float* outr = a.out + row * a.d;
  float acc = 0.f;
  for (int f = tid; f < a.F_t; f += 128) {
    const int s = f / a.n_unm, j = f - s * a.n_unm;
    const float xin = xr[s * a.dim + s_unm[j]];
    const RqsOut o = rqs_element<MODE, KT, INVERSE, true, true>(
        GmemPtr{a.params + (row * a.F_t + f) * a.P}, xin, a.c);
    outr[s * a.dim + a.n_mask + j] = o.y;
    if (a.bins) a.bins[row * a.F_t + f] = (int8_t)o.bin;
    acc += o.lad;
  }
  for (int i = tid; i < a.size * a.n_mask; i += 128) {
    const int s = i / a.n_mask, mi = i - s * a.n_mask;
    outr[s * a.dim + mi] = xr[s * a.dim + s_mask[mi]];
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    const float t = (red[0] + red[1]) + (red[2] + red[3]);
    a.logdet[row] = a.accumulate ? a.logdet[row] + t : t;
  }
}

struct SplitPtr {
  const float *w, *h, *dd;
  int K;
  __device__ __forceinline__ float operator()(int i) const {
    return i < K ? __ldg(w + i) : (i < 2 * K ? __ldg(h + (i - K)) : __ldg(dd + (i - 2 * K)));
  }
  __device__ __forceinline__ float dyn(int base, int i) const { return (*this)(base + i); }
};

template <int MODE, int KT, bool INVERSE>
__global__ void __launch_bounds__(128)
unconstrained_rqs_kernel(const float* __restrict__ inputs, const float* __restrict__ W,
                         const float* __restrict__ H, const float* __restrict__ D,
                         float* __restrict__ out, float* __restrict__ lad,
                         int8_t* __restrict__ bins, long long M, RqsConsts c) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M) return;
  const int K = KT ? KT : c.K;
  const RqsOut o = rqs_element<MODE, KT, INVERSE, false, true>(
      SplitPtr{W + e * K, H + e * K, D + e * (K - 1), K}, inputs[e], c);
  out[e] = o.y;
  lad[e] = o.lad;
  if (bins) bins[e] = (int8_t)o.bin;
}

// Element-wise spline with the LAYER-side normalisation (softmax x 2B, softplus) applied to packed raw
// parameters [M, 3K-1]: the transform step of the autoregressive spline flow NSF_AR (reference
// nf/flows.py:178-190, :196-208), one thread per scalar.
struct PackedPtr {
  const float* p;
  __device__ __forceinline__ float operator()(int i) const { return __ldg(p + i); }
  __device__ __forceinline__ float dyn(int base, int i) const { return __ldg(p + base + i); }
};

template <int MODE, int KT, bool INVERSE>
__global__ void __launch_bounds__(128)
rqs_elementwise_kernel(const float* __restrict__ inputs, const float* __restrict__ params,
                       float* __restrict__ out, float* __restrict__ lad, int8_t* __restrict__ bins,
                       long long M, RqsConsts c) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M) return;
  const int K = KT ? KT : c.K;
  const RqsOut o = rqs_element<MODE, KT, INVERSE, true, true>(PackedPtr{params + e * (3 * K - 1)}, inputs[e], c);
  out[e] = o.y;
  lad[e] = o.lad;
  if (bins) bins[e] = (int8_t)o.bin;
}

template <bool EXACT, int KT, bool LAYER_NORM>
__global__ void __launch_bounds__(128)
debug_knots_kernel(const float* __restrict__ logits, float* __restrict__ knots, long long M,
                   RqsConsts c) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M) return;
  constexpr int KK = KT ? KT : KMAX;
  const int K = KT ? KT : c.K;
  float v[KK + 1];
#pragma unroll
  for (int j = 0; j < KK; ++j)
    if (j < K) v[j] = logits[e * K + j];
  knot_chain<EXACT, KT, LAYER_NORM>(v, c);
#pragma unroll
  for (int j = 0; j <= KK; ++j)
    if (j <= K) knots[e * (K + 1) + j] = v[j];
}

// ---- host side -----------------------------------------------------------------------
static RqsConsts make_consts(int K, float B) {
  RqsConsts c;
  c.B = B;
  c.twoB = (float)(2.0 * (double)B);
  c.negB = -B;
  c.Bnudge = B + 1e-6f;
  c.min_bin = 1e-3f;
  c.one_m = (float)(1.0 - 1e-3 * (double)K);
  c.min_d = 1e-3f;
  c.edge_c = (float)log(exp(1.0 - 1e-3) - 1.0);
  c.edge_d = 1e-3f + log1pf(expf(c.edge_c));
  c.g0 = c.twoB * LOG2E;
  c.q0 = c.twoB * c.one_m;
  c.kstep = c.twoB * 1e-3f;
  c.bin_eps = c.twoB * 2e-5f;
  c.K = K;
  c.scan_order = scan_order();
  return c;
}

RqsConsts make_rqs_consts(int K, float B) { return make_consts(K, B); }

// tuning / comparison hooks (nfk_set_coupling_tune): threads < 0 forces the older tiled kernel for generic geometries
static int g_tune_R = 0, g_tune_threads = 0, g_tune_stages = 0, g_tune_ctas = 0;

static int stream_group_rows(int F_t) {          // rows per group of the streaming kernel: G F_t is a multiple of 32
  int a_ = F_t, b_ = 32;
  while (b_) {
    const int t = a_ % b_;
    a_ = b_;
    b_ = t;
  }
  return 32 / a_;
}

template <int MODE, int KT, bool INVERSE>
static int launch_coupling(CouplingArgs& a, long long N, cudaStream_t st) {
  const int F_t = a.F_t;
  const bool aligned = ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.params) |
                         reinterpret_cast<uintptr_t>(a.out)) & 15) == 0;
  long long done = 0;
  if (aligned && F_t == 32 && a.dim == 2 && a.n_mask == 1 && g_tune_R >= 0) {
    // ---- headline geometry: pairs kernel, 8 rows (256 splines) per tile
    const int R = PAIR_THREADS / 32;
    const long long n_tiles = N / R;
    if (n_tiles > 0 && n_tiles < (1LL << 31)) {
      int stages = g_tune_stages > 0 ? g_tune_stages : 2;
      const size_t stage_bytes = (size_t)PAIR_THREADS * a.P * 4;
      while (stages > 2 && stages * stage_bytes + 64 > 226 * 1024) --stages;
      const size_t smem = stages * stage_bytes + 64;
      if (smem <= 226 * 1024) {
        a.R = R;
        a.stages = stages;
        a.n_tiles = n_tiles;
        auto kern = rqs_coupling_pairs<MODE, KT, INVERSE>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
          set_error("rqs_coupling: cannot set %zu B dynamic shared memory: %s", smem, cudaGetErrorString(e));
          return NFK_ECUDA;
        }
        int ctas_per_sm = g_tune_ctas > 0 ? g_tune_ctas : (int)((227 * 1024) / (smem + 1024));
        ctas_per_sm = max(1, min(ctas_per_sm, 2048 / PAIR_THREADS));
        const long long cap = (long long)sm_count() * ctas_per_sm;
        const long long grid = n_tiles < cap ? n_tiles : cap;
        kern<<<(unsigned)grid, PAIR_THREADS, smem, st>>>(a);
        count_launch();
        if (int rc = check_launch("rqs_coupling_pairs")) return rc;
        done = n_tiles * R;
      }
    }
  } else if (KT == 8 && aligned && g_tune_threads >= 0 && F_t <= 4064 && a.n_mask <= MAXCOND && a.d <= 32000 &&
             N / stream_group_rows(F_t) > 0) {
    // ---- generic geometry, K = 8: warp-autonomous streaming kernel (groups of G rows = whole 32-element blocks)
    const int G = stream_group_rows(F_t);
    const long long n_groups = N / G;
    const int stages = g_tune_stages > 0 ? max(2, g_tune_stages) : 2;
    int ring_floats = 64;
    while (ring_floats < F_t + 32) ring_floats <<= 1;
    a.scratch = ring_floats;
    const size_t per_warp = ((size_t)stages * 32 * a.P + (size_t)ring_floats) * 4 + 64;
    const size_t smem = per_warp * STREAM_WARPS + (size_t)3 * F_t * 2 + 16;
    a.R = G;
    a.stages = stages;
    a.n_tiles = n_groups;
    auto kern = rqs_coupling_stream<MODE, 8, INVERSE>;           // (K = 8 only: the branch is dead for the generic-K instantiation)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("rqs_coupling: cannot set %zu B dynamic shared memory: %s", smem, cudaGetErrorString(e));
      return NFK_ECUDA;
    }
    int ctas_per_sm = g_tune_ctas > 0 ? g_tune_ctas : (int)((227 * 1024) / (smem + 1024));
    ctas_per_sm = max(1, min(ctas_per_sm, 2048 / (STREAM_WARPS * 32)));
    const long long cap = (long long)sm_count() * ctas_per_sm;
    const long long want = (n_groups + STREAM_WARPS - 1) / STREAM_WARPS;
    const long long grid = want < cap ? want : cap;
    kern<<<(unsigned)grid, STREAM_WARPS * 32, smem, st>>>(a);
    count_launch();
    if (int rc = check_launch("rqs_coupling_stream")) return rc;
    done = n_groups * G;
  } else if (aligned) {
    // ---- generic geometry: one element per thread per tile.  Rows per tile: a multiple of 4 (every
    // tile stays 16-byte aligned), as FEW as fill a 128-thread CTA.  Small tiles mean many co-resident
    // CTAs, i.e. many independent TMA rings and barriers per SM; measured on B200 (FAST, 524,288 rows):
    // F_t = 38: R = 4 63.1 %, 8 59.6 %, 12 47.6 %, 16 54.3 % of the HBM peak; F_t = 76: R = 4 65.8 %, 8 62.3 %
    // Ring depth: a CTA keeps stages - 1 tiles in flight while it computes one.  Where registers /
    // threads already cap the co-resident CTAs at <= 5 a third stage costs at most one CTA and pays
    // (F_t = 38: 46.6 % -> 62.8 %; F_t = 76: unchanged); where 6 CTAs fit (F_t = 32, 8) the lost CTA
    // costs more than the deeper ring gains (71.5 % -> 66.7 %), so those keep two stages.
    const size_t row_floats = (size_t)F_t * a.P + a.d;
    int stages = g_tune_stages > 0 ? g_tune_stages : 2;
    int R = g_tune_R > 0 ? ((g_tune_R + 3) & ~3) : 0;
    if (R == 0) {
      for (int r = 4; r <= 64 && r * F_t <= MAX_THREADS; r += 4) {
        const size_t sm = (size_t)stages * r * row_floats * 4 + (size_t)2 * r * a.d * 4 + (size_t)2 * r * F_t * 4 + 320;
        if (sm > 226 * 1024) break;
        R = r;
        if (r * F_t >= 128) break;
      }
      if (R == 0) R = 4;
    }
    const int n_el = R * F_t;
    const int threads = (n_el + 31) & ~31;
    const size_t stage_bytes = (size_t)R * row_floats * 4;
    const size_t fixed = (size_t)2 * R * a.d * 4 + (size_t)2 * n_el * 4 + 8 * 8 + 128;
    if (g_tune_stages <= 0) {
      auto ctas_for = [&](int st) {
        const size_t sm = fixed + (size_t)st * stage_bytes;
        if (sm > 226 * 1024) return 0;
        return max(1, min((int)((227 * 1024) / (sm + 1024)), min(2048 / threads, 65536 / (threads * 80))));
      };
      const int c2 = ctas_for(2), c3 = ctas_for(3);
      if (c3 > 0 && c2 <= 5 && c3 >= c2 - 1) stages = 3;
    }
    const size_t smem = fixed + (size_t)stages * stage_bytes;
    const long long n_tiles = N / R;
    if (threads <= MAX_THREADS && smem <= 226 * 1024 && n_tiles > 0 && n_tiles < (1LL << 31)) {
      a.R = R;
      a.stages = stages;
      a.n_tiles = n_tiles;
      auto kern = rqs_coupling_tiled<MODE, KT, INVERSE>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("rqs_coupling: cannot set %zu B dynamic shared memory: %s", smem, cudaGetErrorString(e));
        return NFK_ECUDA;
      }
      int ctas_per_sm = g_tune_ctas > 0 ? g_tune_ctas : (int)((227 * 1024) / (smem + 1024));
      ctas_per_sm = max(1, min(ctas_per_sm, min(2048 / threads, 65536 / (threads * 80))));
      const long long cap = (long long)sm_count() * ctas_per_sm;
      const long long grid = n_tiles < cap ? n_tiles : cap;
      kern<<<(unsigned)grid, threads, smem, st>>>(a);
      count_launch();
      if (int rc = check_launch("rqs_coupling_tiled")) return rc;
      done = n_tiles * R;
    }
  }
  if (done < N) {
    a.row0 = done;
    rqs_coupling_rows<MODE, KT, INVERSE><<<(unsigned)(N - done), 128, 0, st>>>(a);
    count_launch();
    if (int rc = check_launch("rqs_coupling_rows")) return rc;
  }
  return NFK_OK;
}

#define NFK_DISPATCH_MODE_K_INV(FN, mode, K, inverse, ...)                                   \
  do {                                                                                        \
    const int m_ = (mode);                                                                    \
    const bool k8_ = ((K) == 8);                                                              \
    const bool inv_ = (inverse) != 0;                                                         \
    if (m_ == NFK_ARITH_EXACT) {                                                              \
      if (k8_) return inv_ ? FN<NFK_ARITH_EXACT, 8, true>(__VA_ARGS__)                        \
                           : FN<NFK_ARITH_EXACT, 8, false>(__VA_ARGS__);                      \
      return inv_ ? FN<NFK_ARITH_EXACT, 0, true>(__VA_ARGS__)                                 \
                  : FN<NFK_ARITH_EXACT, 0, false>(__VA_ARGS__);                               \
    } else if (m_ == NFK_ARITH_HYBRID) {                                                      \
      if (k8_) return inv_ ? FN<NFK_ARITH_HYBRID, 8, true>(__VA_ARGS__)                       \
                           : FN<NFK_ARITH_HYBRID, 8, false>(__VA_ARGS__);                     \
      return inv_ ? FN<NFK_ARITH_HYBRID, 0, true>(__VA_ARGS__)                                \
                  : FN<NFK_ARITH_HYBRID, 0, false>(__VA_ARGS__);                              \
    } else {                                                                                  \
      if (k8_) return inv_ ? FN<NFK_ARITH_FAST, 8, true>(__VA_ARGS__)                         \
                           : FN<NFK_ARITH_FAST, 8, false>(__VA_ARGS__);                       \
      return inv_ ? FN<NFK_ARITH_FAST, 0, true>(__VA_ARGS__)                                  \
                  : FN<NFK_ARITH_FAST, 0, false>(__VA_ARGS__);                                \
    }                                                                                         \
  } while (0)

template <int MODE, int KT, bool INVERSE>
static int launch_free(const float* inputs, const float* W, const float* H, const float* D,
                       float* out, float* lad, int8_t* bins, long long M, RqsConsts c,
                       cudaStream_t st) {
  const long long grid = (M + 127) / 128;
  unconstrained_rqs_kernel<MODE, KT, INVERSE><<<(unsigned)grid, 128, 0, st>>>(inputs, W, H, D, out,
                                                                              lad, bins, M, c);
  count_launch();
  return check_launch("unconstrained_rqs");
}

template <int MODE, int KT, bool INVERSE>
static int launch_elementwise(const float* inputs, const float* params, float* out, float* lad, int8_t* bins,
                              long long M, RqsConsts c, cudaStream_t st) {
  const long long grid = (M + 127) / 128;
  rqs_elementwise_kernel<MODE, KT, INVERSE><<<(unsigned)grid, 128, 0, st>>>(inputs, params, out, lad, bins, M, c);
  count_launch();
  return check_launch("rqs_elementwise");
}

static int dispatch_coupling(int mode, int K, int inverse, CouplingArgs& a, long long N,
                             cudaStream_t st) {
  NFK_DISPATCH_MODE_K_INV(launch_coupling, mode, K, inverse, a, N, st);
}
static int dispatch_free(int mode, int K, int inverse, const float* inputs, const float* W,
                         const float* H, const float* D, float* out, float* lad, int8_t* bins,
                         long long M, RqsConsts c, cudaStream_t st) {
  NFK_DISPATCH_MODE_K_INV(launch_free, mode, K, inverse, inputs, W, H, D, out, lad, bins, M, c, st);
}

static int dispatch_elementwise(int mode, int K, int inverse, const float* inputs, const float* params, float* out,
                                float* lad, int8_t* bins, long long M, RqsConsts c, cudaStream_t st) {
  NFK_DISPATCH_MODE_K_INV(launch_elementwise, mode, K, inverse, inputs, params, out, lad, bins, M, c, st);
}

int fill_coupling_geometry(CouplingArgs& a, int size, int dim, const int32_t* mask, int n_mask,
                           int K, float B) {
  NFK_REQUIRE(size > 0 && dim > 1 && dim <= MAXDIM, "rqs_coupling: need size > 0 and 2 <= dim <= %d",
              MAXDIM);
  NFK_REQUIRE(mask != nullptr && n_mask > 0 && n_mask < dim,
              "rqs_coupling: mask must name between 1 and dim-1 columns");
  NFK_REQUIRE(K >= 2 && K <= KMAX, "rqs_coupling: 2 <= K <= %d supported (got %d)", KMAX, K);
  // nf/utils.py:68-71
  NFK_REQUIRE(1e-3 * K <= 1.0, "Minimal bin width too large for the number of bins");
  NFK_REQUIRE(B > 0.f, "rqs_coupling: tail bound must be positive");
  bool used[MAXDIM] = {false};
  for (int i = 0; i < n_mask; ++i) {
    NFK_REQUIRE(mask[i] >= 0 && mask[i] < dim, "rqs_coupling: mask column %d outside [0, %d)",
                mask[i], dim);
    NFK_REQUIRE(!used[mask[i]], "rqs_coupling: mask column %d repeated", mask[i]);
    used[mask[i]] = true;
    a.mask[i] = mask[i];
  }
  int nu = 0;
  for (int cidx = 0; cidx < dim; ++cidx)
    if (!used[cidx]) a.unm[nu++] = cidx;                               // flows.py:225
  for (int i = n_mask; i < MAXDIM; ++i) a.mask[i] = 0;
  for (int i = nu; i < MAXDIM; ++i) a.unm[i] = 0;
  a.size = size;
  a.dim = dim;
  a.n_mask = n_mask;
  a.n_unm = nu;
  a.d = size * dim;
  a.F_t = size * nu;
  a.P = 3 * K - 1;
  a.c = make_consts(K, B);
  return NFK_OK;
}

}  // namespace nfk

using namespace nfk;

extern "C" {

int nfk_set_tuning(int rows_per_tile, int threads, int stages, int ctas_per_sm) {
  g_tune_R = rows_per_tile;
  g_tune_threads = threads;
  g_tune_stages = stages;
  g_tune_ctas = ctas_per_sm;
  return NFK_OK;
}

int nfk_rqs_coupling(const float* x, const float* params, float* out, float* logdet,
                     int8_t* bins, int64_t N, int size, int dim, const int32_t* mask,
                     int n_mask, int K, float B, int inverse, int accumulate, int arith,
                     void* stream) {
  NFK_REQUIRE(N >= 0, "rqs_coupling: negative batch");
  NFK_REQUIRE(arith >= NFK_ARITH_EXACT && arith <= NFK_ARITH_FAST, "rqs_coupling: bad arith %d",
              arith);
  CouplingArgs a{};
  if (int rc = fill_coupling_geometry(a, size, dim, mask, n_mask, K, B)) return rc;
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && params && out && logdet, "rqs_coupling: null device pointer");
  NFK_REQUIRE(x != out, "rqs_coupling: out must not alias x");
  a.x = x;
  a.params = params;
  a.out = out;
  a.logdet = logdet;
  a.bins = bins;
  a.accumulate = accumulate;
  return dispatch_coupling(arith, K, inverse, a, N, (cudaStream_t)stream);
}

int nfk_unconstrained_rqs(const float* inputs, const float* W, const float* H, const float* D,
                          float* out, float* lad, int8_t* bins, int64_t M, int K, float B,
                          int inverse, int arith, void* stream) {
  NFK_REQUIRE(M >= 0, "unconstrained_rqs: negative size");
  NFK_REQUIRE(K >= 2 && K <= KMAX, "unconstrained_rqs: 2 <= K <= %d supported (got %d)", KMAX, K);
  NFK_REQUIRE(1e-3 * K <= 1.0, "Minimal bin width too large for the number of bins");
  NFK_REQUIRE(arith >= NFK_ARITH_EXACT && arith <= NFK_ARITH_FAST, "unconstrained_rqs: bad arith");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(inputs && W && H && D && out && lad, "unconstrained_rqs: null device pointer");
  return dispatch_free(arith, K, inverse, inputs, W, H, D, out, lad, bins, M, make_consts(K, B),
                       (cudaStream_t)stream);
}

int nfk_rqs_elementwise(const float* inputs, const float* params, float* out, float* lad, int8_t* bins,
                        int64_t M, int K, float B, int inverse, int arith, void* stream) {
  NFK_REQUIRE(M >= 0, "rqs_elementwise: negative size");
  NFK_REQUIRE(K >= 2 && K <= KMAX, "rqs_elementwise: 2 <= K <= %d supported (got %d)", KMAX, K);
  NFK_REQUIRE(1e-3 * K <= 1.0, "Minimal bin width too large for the number of bins");
  NFK_REQUIRE(arith >= NFK_ARITH_EXACT && arith <= NFK_ARITH_FAST, "rqs_elementwise: bad arith");
  NFK_REQUIRE(B > 0.f, "rqs_elementwise: tail bound must be positive");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(inputs && params && out && lad, "rqs_elementwise: null device pointer");
  NFK_REQUIRE(M < (1LL << 31) * 128, "rqs_elementwise: too many elements");
  return dispatch_elementwise(arith, K, inverse, inputs, params, out, lad, bins, M, make_consts(K, B),
                              (cudaStream_t)stream);
}

int nfk_debug_knots(const float* logits, float* knots, int64_t M, int K, float B, int layer_norm,
                    int exact, void* stream) {
  NFK_REQUIRE(K >= 2 && K <= KMAX, "debug_knots: bad K");
  if (M == 0) return NFK_OK;
  const RqsConsts c = make_consts(K, B);
  const unsigned grid = (unsigned)((M + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
#define NFK_DK(EX, KT, LN) debug_knots_kernel<EX, KT, LN><<<grid, 128, 0, st>>>(logits, knots, M, c)
  if (K == 8) {
    if (exact) { if (layer_norm) NFK_DK(true, 8, true); else NFK_DK(true, 8, false); }
    else       { if (layer_norm) NFK_DK(false, 8, true); else NFK_DK(false, 8, false); }
  } else {
    if (exact) { if (layer_norm) NFK_DK(true, 0, true); else NFK_DK(true, 0, false); }
    else       { if (layer_norm) NFK_DK(false, 0, true); else NFK_DK(false, 0, false); }
  }
#undef NFK_DK
  count_launch();
  return check_launch("debug_knots");
}

}  // extern "C"
