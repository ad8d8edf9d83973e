// Wide conditioner path (hidden width > 128, e.g. the class default 800 of NSF_CL, reference
// nf/flows.py:216 / FCNN nf/flows.py:20-35): a persistent, warp-specialised tcgen05 GEMM
//     Y = act(A W^T + b)
// whose operands never need a tensor map because every matrix on this path lives in HBM in the
// *shared-memory image* the tensor core reads: 128-row x 64-column bf16 blocks (16 KB), K-major
// SWIZZLE_128B (16-byte chunk j of row r holds source chunk j ^ (r % 8)).  One 1-D TMA bulk
// copy (UBLKCP) moves a block; the epilogue writes the next layer's operand in the same image,
// so activations go HBM -> smem -> tensor core with no layout work anywhere.
//
//   A image   [m_tiles][KB][128 rows][128 B]                     (activations, 128 rows per tile)
//   W image   per N tile t (64*nb_t output columns): [KB][64*nb_t rows][128 B]
//   out image [m_tiles][sum nb][128 rows][128 B]  (bf16, tanh)  or fp32 rows [M, ldy]
//
// CTA = 10 warps, one CTA per SM, persistent over (m_tile, n_tile) work items:
//   warp 0      producer: waits empty[s], issues the two bulk copies of a K block into stage s
//   warp 1      MMA issuer: waits full[s], 4 x tcgen05.mma (M=128, N=64*nb, K=16) per K block,
//               tcgen05.commit -> empty[s]; after the last K block commit -> tmem_full[acc]
//   warps 2..9  epilogue: wait tmem_full[acc], tcgen05.ld 32 columns at a time (lane quadrant =
//               warp % 4, column half = (warp-2) / 4), bias + tanh.approx, pack to bf16, write the
//               swizzled 16 KB output block in shared memory, one thread bulk-stores it.
// The accumulator is double buffered in TMEM (2 x 256 columns) so the epilogue of tile i runs
// under the MMAs of tile i+1.  Four 48 KB operand stages + 32 KB of output staging = 224 KB.
//
// EPI_RQS (size = 32, dim = 2, K = 8 layers): the last GEMM's epilogue IS the spline transform
// (reference nf/flows.py:232-239 + nf/utils.py:20-152).  W3's 23 rows per feature are padded to
// 24, an N tile holds 8 features (192 accumulator columns), 16 epilogue warps: thread
// (row, feature) pulls its 24 raw parameters out of TMEM, evaluates bin search + spline +
// log|det| in registers (rqs_math.cuh) and writes the (conditioning, transformed) output pair;
// the [N, 32, 23] parameter tensor never exists in HBM.
#include "rqs_bwd_math.cuh"
#include "tc05.cuh"
#include <type_traits>

namespace nfk {

enum { EPI_BF16_IMG = 0, EPI_F32_ROWS = 1, EPI_RQS = 2, EPI_RQS_BWD = 3 };

constexpr int WS_M = 128;
constexpr int WS_MAX_TILES = 16;
constexpr uint32_t WS_BLK = 128 * 128;                 // one 128 x 64 bf16 block
constexpr uint32_t WS_A_BYTES = WS_BLK;
constexpr int WS_PC = 24, WS_TF = 8;    // EPI_RQS: accumulator columns per feature, features per N tile

template <int EPI>
struct WsCfg {                                         // plain GEMM epilogues
  static constexpr int STAGES = 4;
  static constexpr int EPI_WARPS = 8;
  static constexpr uint32_t B_BYTES = 256 * 128;
  static constexpr uint32_t STG_BYTES = 2 * WS_BLK;    // output staging
};
template <>
struct WsCfg<EPI_BF16_IMG> {                           // two teams of 8 epilogue warps, one staging block each
  static constexpr int STAGES = 4;
  static constexpr int EPI_WARPS = 16;
  static constexpr uint32_t B_BYTES = 256 * 128;
  static constexpr uint32_t STG_BYTES = 2 * WS_BLK;
};
template <>
struct WsCfg<EPI_RQS> {
  static constexpr int STAGES = 5;
  static constexpr int EPI_WARPS = 16;
  static constexpr uint32_t B_BYTES = WS_TF * WS_PC * 128;                       // 192 rows
  static constexpr uint32_t STG_BYTES = WS_MAX_TILES * WS_TF * WS_PC * 4 + 2 * WS_M * 4 * 4;   // b3 + log-det partials
};
template <>
struct WsCfg<EPI_RQS_BWD> {                            // spline backward as the epilogue: grad_params image out
  static constexpr int STAGES = 4;
  static constexpr int EPI_WARPS = 16;
  static constexpr uint32_t B_BYTES = WS_TF * WS_PC * 128;
  static constexpr uint32_t STG_BYTES = 3 * WS_BLK + WS_MAX_TILES * WS_TF * WS_PC * 4;   // G tile + b3
};
template <int EPI>
constexpr size_t ws_smem() {
  return (size_t)WsCfg<EPI>::STAGES * (WS_A_BYTES + WsCfg<EPI>::B_BYTES) + WsCfg<EPI>::STG_BYTES + 32 * 8 + 1024;
}
static_assert(ws_smem<EPI_BF16_IMG>() <= 227 * 1024 && ws_smem<EPI_RQS>() <= 227 * 1024 &&
                  ws_smem<EPI_RQS_BWD>() <= 227 * 1024,
              "gemm_ws exceeds the 227 KB shared-memory limit");

// grouped launch: several independent GEMMs (same N-tile plan, own operands / K depth) share one
// persistent kernel -- the dim-1 per-dimension conditioners of the autoregressive flow NSF_AR
struct WsGroup {
  const unsigned char* a_img;
  const unsigned char* w_img;
  const float* bias;
  void* out;
  int KB, kmma_last;
  int a_kb;                // K blocks per M tile of a_img (>= KB: several groups may read a prefix of one image)
  int pad_;
};

struct WsArgs {
  const WsGroup* groups;   // device array, or null
  int n_groups;
  const unsigned char* a_img;
  const unsigned char* w_img;
  const float* bias;       // [64 * sum nb], zero padded
  void* out;
  long long m_tiles;
  long long M;             // real rows (fp32 row output is bounds-checked against it)
  long long ldy;           // fp32 row stride in floats
  int KB;                  // K blocks of 64
  int kmma_last;           // tcgen05.mma (K=16) count in the last K block, 1..4
  int n_tiles;
  int nb[WS_MAX_TILES];    // 64-column blocks per N tile (1..4)
  int n_out;               // real output columns (fp32 row output)
  int act;                 // 0 identity, 1 tanh
  int fmt;                 // 16-bit format of every operand / output image of this launch: NFK_IMG_BF16 or NFK_IMG_F16
  // EPI_RQS only
  const float* x;          // [M, size*dim]
  float* logdet;           // [M]
  int dim, n_mask, n_feat; // columns per group, conditioning columns per group, transformed features
  int mask[4], unm[4];     // conditioning / transformed column indices inside a group
  int accumulate;
  float* dbg_params;       // optional test hook [M][8*n_tiles][24]: raw spline parameters (accumulator + b3)
  signed char* dbg_bins;   // optional [M][8*n_tiles]: bin used per element (-1 = identity tail)
  RqsConsts c;
  // act == 2 (bf16 image epilogue): out = acc * (1 - h^2), h = aux image laid out like out
  const unsigned char* aux;
  // EPI_RQS_BWD only
  const float* gout;       // [M, size*dim] dL/d(layer output)
  float* gin;              // [M, size*dim] dL/d(layer input), direct path
  const float* gld;        // [M] dL/dlogdet, or null: gld_const for every row
  float gld_const;
};

// work item pt -> (operands of its group, M-tile index inside the group)
__device__ __forceinline__ WsGroup ws_item(const WsArgs& a, long long pt, long long& ptile) {
  if (a.n_groups > 0) {
    ptile = pt / a.n_groups;
    return a.groups[pt % a.n_groups];
  }
  ptile = pt;
  WsGroup g;
  g.a_img = a.a_img;
  g.w_img = a.w_img;
  g.bias = a.bias;
  g.out = a.out;
  g.KB = a.KB;
  g.kmma_last = a.kmma_last;
  g.a_kb = a.KB;
  g.pad_ = 0;
  return g;
}

// two fp32 -> one 32-bit word of the image format (fp16 saturates to +-65504: operands of unknown range)
__device__ __forceinline__ uint32_t pack_img2(float lo, float hi, int fmt) {
  return fmt == NFK_IMG_F16 ? pack_f16x2_sat(lo, hi) : pack_bf16x2(lo, hi);
}

__device__ __forceinline__ bool ws_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- CTA-pair (cta_group::2) helpers: two SMs of one TPC run ONE 256-row MMA, each holding its own
// 128 A rows and HALF of the B tile, so the shared-memory port sees half the B traffic per MMA.
constexpr uint32_t WS_PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit: the even CTA's copy of an address
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of the pair -> the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// arrive on the EVEN CTA's copy of a barrier (local when executed by the even CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  // default (cta-scope) semantics on purpose: what is handed over lives in TMEM / TMA-written shared memory
  // and is ordered by the tcgen05 fences and the mbarrier itself; a cluster-scope release/acquire would make
  // ptxas emit an L1 invalidate (CCTL.IVALL) inside the spin loops
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & WS_PEER_MASK) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

template <int EPI, int MODE, bool INVERSE, bool PAIRS, bool TWO>
__global__ void __launch_bounds__((2 + WsCfg<EPI>::EPI_WARPS) * 32, 1)
gemm_ws_kernel(const __grid_constant__ WsArgs a) {
  // pair mode: a CTA stages only half of the B tile, the saved shared memory buys deeper pipelining
  constexpr uint32_t WS_STAGE_BYTES = WS_A_BYTES + (TWO ? WsCfg<EPI>::B_BYTES / 2 : WsCfg<EPI>::B_BYTES);
  constexpr int WS_STAGES = (int)((WsCfg<EPI>::STAGES * (WS_A_BYTES + WsCfg<EPI>::B_BYTES)) / WS_STAGE_BYTES);
  constexpr int WS_EPI_WARPS = WsCfg<EPI>::EPI_WARPS;
  constexpr uint32_t WS_STG_BYTES = WsCfg<EPI>::STG_BYTES;
  constexpr bool OUT_F32 = (EPI == EPI_F32_ROWS);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned char* stg = sm + WS_STAGES * WS_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + WS_STG_BYTES);
  uint64_t* full = bars;                    // [STAGES] operands landed      (1 + tx)
  uint64_t* empty = bars + WS_STAGES;       // [STAGES] stage consumed       (1, tcgen05.commit)
  uint64_t* tfull = bars + 2 * WS_STAGES;   // [2] accumulator complete      (1, tcgen05.commit)
  uint64_t* tempty = tfull + 2;             // [2] accumulator drained       (epilogue warps; pair: of both CTAs)
  uint64_t* pfull = tempty + 2;             // [STAGES] pair mode: the odd CTA's operands landed (1, remote arrive)
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = TWO ? cluster_ctarank() : 0u;
  if (warp == 0) {
    if (TWO) tmem_alloc2(&tmem_base_s, 512); else tmem_alloc(&tmem_base_s, 512);
  }
  if (tid == 32) {
    for (int i = 0; i < WS_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&pfull[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], TWO ? 2 * WS_EPI_WARPS : WS_EPI_WARPS);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();              // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // work items: M tiles (pair mode: pairs of M tiles, CTA rank r owns tile 2*pt + r)
  const long long first = TWO ? (blockIdx.x >> 1) : blockIdx.x, stride = TWO ? (gridDim.x >> 1) : gridDim.x;
  const long long n_pt = (TWO ? (a.m_tiles + 1) / 2 : a.m_tiles) * (a.n_groups > 0 ? a.n_groups : 1);

  if (warp == 0) {
    // ================================ producer ================================
    uint32_t s = 0, ph = 0;
    for (long long pt = first; pt < n_pt; pt += stride) {
      long long ptile;
      const WsGroup G = ws_item(a, pt, ptile);
      const long long mt = TWO ? 2 * ptile + crank : ptile;
      const long long mt_ld = mt < a.m_tiles ? mt : a.m_tiles - 1;      // odd tile count: the pair's spare half
      const unsigned char* ag = G.a_img + (size_t)mt_ld * G.a_kb * WS_BLK;
      const unsigned char* wg = G.w_img;
      for (int t = 0; t < a.n_tiles; ++t) {
        const uint32_t bb = (uint32_t)a.nb[t] * 64u * 128u;
        const uint32_t bl = TWO ? bb / 2 : bb;                           // pair mode: this CTA's half of the B rows
        for (int kb = 0; kb < G.KB; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          if (lane == 0) {
            unsigned char* st = sm + s * WS_STAGE_BYTES;
            mbar_expect_tx(&full[s], WS_A_BYTES + bl);
            bulk_g2s(st, ag + (size_t)kb * WS_BLK, WS_A_BYTES, &full[s]);
            bulk_g2s(st + WS_A_BYTES, wg + (size_t)kb * bb + (size_t)crank * bl, bl, &full[s]);
          }
          __syncwarp();
          if (++s == WS_STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        wg += (size_t)G.KB * bb;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    uint32_t s = 0, ph = 0, tl = 0;
    if (TWO && crank == 1) {
      // odd CTA of the pair: no MMA issue -- forward "my operands landed" to the even CTA
      for (long long pt = first; pt < n_pt; pt += stride) {
        long long ptile;
        const WsGroup G = ws_item(a, pt, ptile);
        for (int t = 0; t < a.n_tiles; ++t)
          for (int kb = 0; kb < G.KB; ++kb) {
            mbar_wait(&full[s], ph);
            if (lane == 0) mbar_arrive_leader(&pfull[s]);
            __syncwarp();
            if (++s == WS_STAGES) {
              s = 0;
              ph ^= 1;
            }
          }
      }
    } else {
      for (long long pt = first; pt < n_pt; pt += stride) {
        long long ptile;
        const WsGroup G = ws_item(a, pt, ptile);
        for (int t = 0; t < a.n_tiles; ++t, ++tl) {
          const uint32_t acc = tl & 1;
          const uint32_t idesc = a.fmt == NFK_IMG_F16 ? make_idesc_f16(TWO ? 2 * WS_M : WS_M, a.nb[t] * 64)
                                                      : make_idesc_bf16(TWO ? 2 * WS_M : WS_M, a.nb[t] * 64);
          mbar_wait(&tempty[acc], ((tl >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem + acc * 256;
          for (int kb = 0; kb < G.KB; ++kb) {
            mbar_wait(&full[s], ph);
            if (TWO) mbar_wait(&pfull[s], ph);
            tc_fence_after();
            if (ws_elect_one()) {
              const uint32_t aa = smem_u32(sm + s * WS_STAGE_BYTES), ba = aa + WS_A_BYTES;
              const int nm = (kb == G.KB - 1) ? G.kmma_last : 4;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nm) {
                  if (TWO)
                    umma2_bf16(d, make_desc_sw128(aa + k * 32), make_desc_sw128(ba + k * 32), idesc, (kb | k) ? 1u : 0u);
                  else
                    umma_bf16(d, make_desc_sw128(aa + k * 32), make_desc_sw128(ba + k * 32), idesc, (kb | k) ? 1u : 0u);
                }
              if (TWO) {
                umma2_commit(&empty[s]);                     // frees the stage in both CTAs
                if (kb == G.KB - 1) umma2_commit(&tfull[acc]);
              } else {
                umma_commit(&empty[s]);
                if (kb == G.KB - 1) umma_commit(&tfull[acc]);
              }
            }
            __syncwarp();
            if (++s == WS_STAGES) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if constexpr (EPI == EPI_RQS) {
    // ===================== epilogue: spline transform out of TMEM =====================
    const int ew = warp - 2;
    const int q = warp & 3;              // TMEM lane quadrant this warp may read
    const int slice = ew >> 2;           // features slice and slice + 4 of every 8-feature tile
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    float* sB3 = reinterpret_cast<float*>(stg);
    float* sLd = sB3 + WS_MAX_TILES * WS_TF * WS_PC;     // [2][128][4]
    for (int i = tid - 64; i < a.n_tiles * WS_TF * WS_PC; i += WS_EPI_WARPS * 32) sB3[i] = a.bias[i];
    asm volatile("bar.sync 5, %0;" ::"n"(WS_EPI_WARPS * 32) : "memory");
    const int D = a.dim, n_un = a.dim - a.n_mask, d = a.dim * (a.n_feat / n_un);
    const int unm0 = a.unm[0], unm1 = a.unm[1], unm2 = a.unm[2];
    const int mask0 = a.mask[0], mask1 = a.mask[1], mask2 = a.mask[2];
    uint32_t tl = 0, it = 0;
    for (long long pt = first; pt < n_pt; pt += stride, ++it) {
      long long ptile;
      const WsGroup G = ws_item(a, pt, ptile);
      const long long mt = TWO ? 2 * ptile + crank : ptile;
      const long long grow = mt * WS_M + row;
      const bool live = grow < a.M;
      const float* xr = a.x + grow * d;
      float* orow = reinterpret_cast<float*>(a.out) + grow * d;
      float ld_old = 0.f;
      if (slice == 0 && live && a.accumulate) ld_old = __ldg(a.logdet + grow);
      float lad_acc = 0.f;
      for (int t = 0; t < a.n_tiles; ++t, ++tl) {
        const uint32_t acc = tl & 1;
        // feature f = (group s, transformed column u): reads x[s*D + unm[u]], writes out[s*D + n_mask + u]
        float xin[2], xcond[2];
        int gbase[2], uu[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int f = t * WS_TF + slice + 4 * e;
          if (PAIRS) {                                   // dim 2, one conditioning column: one 8-byte access
            gbase[e] = live ? 2 * f : -1;
            uu[e] = 0;
            const float2 xc = live ? __ldg(reinterpret_cast<const float2*>(xr + 2 * f)) : make_float2(0.f, 0.f);
            xin[e] = unm0 ? xc.y : xc.x;
            xcond[e] = unm0 ? xc.x : xc.y;
          } else {
            const int sgrp = n_un == 1 ? f : (n_un == 2 ? (f >> 1) : f / 3);
            uu[e] = f - sgrp * n_un;
            gbase[e] = (live && f < a.n_feat) ? sgrp * D : -1;
            const int col = uu[e] == 0 ? unm0 : (uu[e] == 1 ? unm1 : unm2);
            xin[e] = gbase[e] >= 0 ? __ldg(xr + gbase[e] + col) : 0.f;
            xcond[e] = 0.f;
          }
        }
        mbar_wait(&tfull[acc], (tl >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int f = t * WS_TF + slice + 4 * e;
          uint32_t v[24];
          const uint32_t ta = tmem + acc * 256 + lane_sel + (uint32_t)((slice + 4 * e) * WS_PC);
          tmem_ld16(ta, v);
          tmem_ld8(ta + 16, v + 16);
          tmem_ld_wait();
          if (e == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (TWO) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }      // the accumulator may be overwritten
          }
          const RqsOut o = rqs_element<MODE, 8, INVERSE, true, true>(RegParams{v, sB3 + f * WS_PC}, xin[e], a.c);
          if (a.dbg_params && live) {            // test hook (uniform branch)
            const size_t ei = (size_t)grow * (a.n_tiles * WS_TF) + f;
#pragma unroll
            for (int i = 0; i < WS_PC; ++i)
              a.dbg_params[ei * WS_PC + i] = i < 23 ? __uint_as_float(v[i]) + sB3[f * WS_PC + i] : 0.f;
            if (a.dbg_bins) a.dbg_bins[ei] = (signed char)(gbase[e] >= 0 ? o.bin : -1);
          }
          if (gbase[e] >= 0) {
            // transformed columns follow the conditioning ones inside a group: quirk Q5
            if (PAIRS) {
              *reinterpret_cast<float2*>(orow + gbase[e]) = make_float2(xcond[e], o.y);
            } else {
              orow[gbase[e] + a.n_mask + uu[e]] = o.y;
              if (uu[e] == 0) {
                orow[gbase[e]] = __ldg(xr + gbase[e] + mask0);
                if (a.n_mask > 1) orow[gbase[e] + 1] = __ldg(xr + gbase[e] + mask1);
                if (a.n_mask > 2) orow[gbase[e] + 2] = __ldg(xr + gbase[e] + mask2);
              }
            }
            lad_acc += o.lad;
          }
        }
      }
      // row log-det = sum over the features (flows.py:238): 4 partial sums per row
      float* sl = sLd + (it & 1) * (WS_M * 4);
      sl[row * 4 + slice] = lad_acc;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      if (slice == 0 && live) {
        const float4 p = *reinterpret_cast<const float4*>(sl + row * 4);
        const float tsum = (p.x + p.y) + (p.z + p.w);
        a.logdet[grow] = a.accumulate ? ld_old + tsum : tsum;
      }
    }
  } else if constexpr (EPI == EPI_RQS_BWD) {
    // ============ epilogue: spline BACKWARD out of TMEM -> grad_params image + direct grad_x ============
    const int ew = warp - 2;
    const int q = warp & 3;
    constexpr int FPT = WS_TF / (WS_EPI_WARPS / 4);   // features per thread and tile
    const int slice = ew >> 2;           // features FPT*slice .. FPT*slice+FPT-1 of every 8-feature tile
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const bool issuer = (ew == 0 && lane == 0);
    unsigned char* sG = stg;                                             // [3][128][128 B] staged G tile
    float* sB3 = reinterpret_cast<float*>(stg + 3 * WS_BLK);
    for (int i = tid - 64; i < a.n_tiles * WS_TF * WS_PC; i += WS_EPI_WARPS * 32) sB3[i] = a.bias[i];
    asm volatile("bar.sync 5, %0;" ::"n"(WS_EPI_WARPS * 32) : "memory");
    const int D = a.dim, n_un = a.dim - a.n_mask, d = a.dim * (a.n_feat / n_un);
    const int unm0 = a.unm[0], unm1 = a.unm[1], unm2 = a.unm[2];
    const int mask0 = a.mask[0], mask1 = a.mask[1], mask2 = a.mask[2];
    const int ob_total = 3 * a.n_tiles;
    uint32_t tl = 0;
    for (long long pt = first; pt < n_pt; pt += stride) {
      long long ptile;
      const WsGroup G = ws_item(a, pt, ptile);
      const long long mt = TWO ? 2 * ptile + crank : ptile;
      const long long grow = mt * WS_M + row;
      const bool live = grow < a.M;
      const float* xr = a.x + grow * d;
      const float* gor = a.gout + grow * d;
      float* gir = a.gin + grow * d;
      const float gl = live ? (a.gld ? __ldg(a.gld + grow) : a.gld_const) : 0.f;
      for (int t = 0; t < a.n_tiles; ++t, ++tl) {
        const uint32_t acc = tl & 1;
        float xin[FPT], gyv[FPT];
        int gbase[FPT], uu[FPT], tcol[FPT];
#pragma unroll
        for (int e = 0; e < FPT; ++e) {
          const int f = t * WS_TF + slice * FPT + e;
          const int sgrp = n_un == 1 ? f : (n_un == 2 ? (f >> 1) : f / 3);
          uu[e] = f - sgrp * n_un;
          gbase[e] = (live && f < a.n_feat) ? sgrp * D : -1;
          tcol[e] = uu[e] == 0 ? unm0 : (uu[e] == 1 ? unm1 : unm2);
          xin[e] = gbase[e] >= 0 ? __ldg(xr + gbase[e] + tcol[e]) : 0.f;
          gyv[e] = gbase[e] >= 0 ? __ldg(gor + gbase[e] + a.n_mask + uu[e]) : 0.f;
        }
        mbar_wait(&tfull[acc], (tl >> 1) & 1);
        tc_fence_after();
        // the previous tile's bulk store must have finished reading the staging tile
        if (issuer) bulk_wait_read<0>();
        asm volatile("bar.sync 1, %0;" ::"n"(WS_EPI_WARPS * 32) : "memory");
#pragma unroll
        for (int e = 0; e < FPT; ++e) {
          const int fs = slice * FPT + e;                      // feature slot inside the tile
          const int f = t * WS_TF + fs;
          uint32_t v[24];
          const uint32_t ta = tmem + acc * 256 + lane_sel + (uint32_t)(fs * WS_PC);
          tmem_ld16(ta, v);
          tmem_ld8(ta + 16, v + 16);
          tmem_ld_wait();
          if (e == FPT - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (TWO) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
          }
          float gxv, gp[24];
          rqs_element_bwd<8, true>(RegParams{v, sB3 + f * WS_PC}, xin[e], gyv[e], gl, INVERSE, a.c, gxv, gp);
          gp[23] = 0.f;
          if (gbase[e] >= 0) {
            gir[gbase[e] + tcol[e]] = gxv;
            if (uu[e] == 0) {                                // conditioning columns: out[:, s, j] = x[:, s, mask[j]]
              gir[gbase[e] + mask0] = __ldg(gor + gbase[e]);
              if (a.n_mask > 1) gir[gbase[e] + mask1] = __ldg(gor + gbase[e] + 1);
              if (a.n_mask > 2) gir[gbase[e] + mask2] = __ldg(gor + gbase[e] + 2);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 23; ++i) gp[i] = 0.f;        // padded rows / features: keep the image finite
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            uint4 u;
            u.x = pack_bf16x2(gp[8 * j + 0], gp[8 * j + 1]);
            u.y = pack_bf16x2(gp[8 * j + 2], gp[8 * j + 3]);
            u.z = pack_bf16x2(gp[8 * j + 4], gp[8 * j + 5]);
            u.w = pack_bf16x2(gp[8 * j + 6], gp[8 * j + 7]);
            const int ch = 3 * fs + j;                       // 16-byte chunk of the 192-column tile
            *reinterpret_cast<uint4*>(sG + (ch >> 3) * WS_BLK + row * 128 + (((ch & 7) ^ (row & 7)) << 4)) = u;
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(WS_EPI_WARPS * 32) : "memory");
        if (issuer && mt < a.m_tiles) {
          unsigned char* og = reinterpret_cast<unsigned char*>(a.out) + ((size_t)mt * ob_total + 3 * t) * WS_BLK;
          bulk_s2g(og, sG, 3 * WS_BLK);
          bulk_commit();
        }
      }
    }
    if (issuer) bulk_wait_all<0>();
  } else if constexpr (EPI == EPI_BF16_IMG) {
    // ============== epilogue: bias + tanh -> bf16 image block -> TMA bulk store ==============
    // Two teams of 8 warps take alternate 64-column blocks; a team owns one 16 KB staging block.
    const int ew = warp - 2;
    const int team = ew >> 3;
    const int q = warp & 3;              // TMEM lane quadrant this warp may read
    const int h = (ew >> 2) & 1;         // 32-column half of the block
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const bool issuer = ((ew & 7) == 0 && lane == 0);
    unsigned char* sb = stg + team * WS_BLK;
    uint32_t tl = 0, bc = 0;
    int ob_total = 0;
    for (int t = 0; t < a.n_tiles; ++t) ob_total += a.nb[t];
    for (long long pt = first; pt < n_pt; pt += stride) {
      long long ptile;
      const WsGroup G = ws_item(a, pt, ptile);
      const long long mt = TWO ? 2 * ptile + crank : ptile;
      int ob0 = 0;
      for (int t = 0; t < a.n_tiles; ++t, ++tl) {
        const uint32_t acc = tl & 1;
        const int nb = a.nb[t];
        mbar_wait(&tfull[acc], (tl >> 1) & 1);
        tc_fence_after();
        const int b_first = ((bc & 1) == (uint32_t)team) ? 0 : 1;      // this team's blocks: b_first, +2, ...
        if (b_first >= nb) {                                           // nothing of this tile is ours
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (TWO) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
        }
        for (int b = b_first; b < nb; b += 2) {
          uint32_t v[32];
          tmem_ld32(tmem + acc * 256 + lane_sel + (uint32_t)(b * 64 + h * 32), v);
          tmem_ld_wait();
          if (b + 2 >= nb) {
            // this warp's last columns of the accumulator are in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (TWO) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
          }
          const int colb = (ob0 + b) * 64 + h * 32;        // first padded output column of v[]
          const float4* bp = reinterpret_cast<const float4*>(G.bias + colb);
          uint4 u[4];
          // the image format is uniform per launch: one specialised copy of the conversion loop per format
          auto convert = [&](auto is_f16) {
          constexpr int FMT = decltype(is_f16)::value ? NFK_IMG_F16 : NFK_IMG_BF16;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = __ldg(bp + 2 * j), b1 = __ldg(bp + 2 * j + 1);
            float f[8];
            f[0] = __uint_as_float(v[8 * j + 0]) + b0.x;
            f[1] = __uint_as_float(v[8 * j + 1]) + b0.y;
            f[2] = __uint_as_float(v[8 * j + 2]) + b0.z;
            f[3] = __uint_as_float(v[8 * j + 3]) + b0.w;
            f[4] = __uint_as_float(v[8 * j + 4]) + b1.x;
            f[5] = __uint_as_float(v[8 * j + 5]) + b1.y;
            f[6] = __uint_as_float(v[8 * j + 6]) + b1.z;
            f[7] = __uint_as_float(v[8 * j + 7]) + b1.w;
            if (a.act == 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = tanh_approx(f[e]);
            } else if (a.act == 2) {
              // tanh backward: multiply by 1 - h^2, h from the saved activation image (same block, same chunk)
              const uint4 hv = __ldg(reinterpret_cast<const uint4*>(
                  a.aux + ((size_t)(mt < a.m_tiles ? mt : a.m_tiles - 1) * ob_total + (ob0 + b)) * WS_BLK + row * 128 + (((h * 4 + j) ^ (row & 7)) << 4)));
              const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float h0 = __uint_as_float(hw[e] << 16), h1 = __uint_as_float(hw[e] & 0xffff0000u);
                f[2 * e] *= fmaf(-h0, h0, 1.f);
                f[2 * e + 1] *= fmaf(-h1, h1, 1.f);
              }
            }
            u[j].x = pack_img2(f[0], f[1], FMT);
            u[j].y = pack_img2(f[2], f[3], FMT);
            u[j].z = pack_img2(f[4], f[5], FMT);
            u[j].w = pack_img2(f[6], f[7], FMT);
          }
          };
          if (a.fmt == NFK_IMG_F16) convert(std::true_type{}); else convert(std::false_type{});
          // the team's previous bulk store must have finished reading the staging block
          if (issuer) bulk_wait_read<0>();
          asm volatile("bar.sync %0, 256;" ::"r"(1 + team) : "memory");
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = h * 4 + j;
            *reinterpret_cast<uint4*>(sb + row * 128 + ((ch ^ (row & 7)) << 4)) = u[j];
          }
          fence_proxy_async();
          asm volatile("bar.sync %0, 256;" ::"r"(1 + team) : "memory");
          if (issuer && mt < a.m_tiles) {
            unsigned char* og = reinterpret_cast<unsigned char*>(G.out) +
                                ((size_t)mt * ob_total + (ob0 + b)) * WS_BLK;
            bulk_s2g(og, sb, WS_BLK);
            bulk_commit();
          }
        }
        bc += nb;
        ob0 += nb;
      }
    }
    if (issuer) bulk_wait_all<0>();
  } else {
    // ================================ epilogue ================================
    const int ew = warp - 2;
    const int q = warp & 3;              // TMEM lane quadrant this warp may read
    const int h = ew >> 2;               // 32-column half of each 64-column block
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const bool issuer = (ew == 0 && lane == 0);
    uint32_t tl = 0, bc = 0;
    int ob_total = 0;
    for (int t = 0; t < a.n_tiles; ++t) ob_total += a.nb[t];
    for (long long pt = first; pt < n_pt; pt += stride) {
      long long ptile;
      const WsGroup G = ws_item(a, pt, ptile);
      const long long mt = TWO ? 2 * ptile + crank : ptile;
      int ob0 = 0;
      for (int t = 0; t < a.n_tiles; ++t, ++tl) {
        const uint32_t acc = tl & 1;
        const int nb = a.nb[t];
        mbar_wait(&tfull[acc], (tl >> 1) & 1);
        tc_fence_after();
        for (int b = 0; b < nb; ++b, ++bc) {
          uint32_t v[32];
          tmem_ld32(tmem + acc * 256 + lane_sel + (uint32_t)(b * 64 + h * 32), v);
          tmem_ld_wait();
          if (b == nb - 1) {
            // every column of this accumulator is in registers: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (TWO) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
          }
          const int colb = (ob0 + b) * 64 + h * 32;        // first padded output column of v[]
          const float4* bp = reinterpret_cast<const float4*>(G.bias + colb);
          if (!OUT_F32) {
            unsigned char* sb = stg + (bc & 1) * WS_BLK;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b0 = __ldg(bp + 2 * j), b1 = __ldg(bp + 2 * j + 1);
              float f[8];
              f[0] = __uint_as_float(v[8 * j + 0]) + b0.x;
              f[1] = __uint_as_float(v[8 * j + 1]) + b0.y;
              f[2] = __uint_as_float(v[8 * j + 2]) + b0.z;
              f[3] = __uint_as_float(v[8 * j + 3]) + b0.w;
              f[4] = __uint_as_float(v[8 * j + 4]) + b1.x;
              f[5] = __uint_as_float(v[8 * j + 5]) + b1.y;
              f[6] = __uint_as_float(v[8 * j + 6]) + b1.z;
              f[7] = __uint_as_float(v[8 * j + 7]) + b1.w;
              if (a.act == 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = tanh_approx(f[e]);
              }
              uint4 u;
              u.x = pack_img2(f[0], f[1], a.fmt);
              u.y = pack_img2(f[2], f[3], a.fmt);
              u.z = pack_img2(f[4], f[5], a.fmt);
              u.w = pack_img2(f[6], f[7], a.fmt);
              const int ch = h * 4 + j;
              *reinterpret_cast<uint4*>(sb + row * 128 + ((ch ^ (row & 7)) << 4)) = u;
            }
            fence_proxy_async();
            // the store that used the *other* buffer must have finished reading shared memory
            // before anyone passes this barrier and starts the next block in it
            if (issuer) bulk_wait_read<0>();
            asm volatile("bar.sync 1, %0;" ::"n"(WS_EPI_WARPS * 32) : "memory");
            if (issuer) {
              unsigned char* og = reinterpret_cast<unsigned char*>(G.out) +
                                  ((size_t)mt * ob_total + (ob0 + b)) * WS_BLK;
              bulk_s2g(og, sb, WS_BLK);
              bulk_commit();
            }
          } else {
            // fp32 rows: stage the 128 x 64 block (256 B per row, 16-byte chunk c of row r at
            // c ^ (r % 16)), then every warp writes 16 rows as coalesced 128-byte segments
            float* sf = reinterpret_cast<float*>(stg);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(bp + j);
              float4 o;
              o.x = __uint_as_float(v[4 * j + 0]) + bv.x;
              o.y = __uint_as_float(v[4 * j + 1]) + bv.y;
              o.z = __uint_as_float(v[4 * j + 2]) + bv.z;
              o.w = __uint_as_float(v[4 * j + 3]) + bv.w;
              if (a.act == 1) {
                o.x = tanh_approx(o.x);
                o.y = tanh_approx(o.y);
                o.z = tanh_approx(o.z);
                o.w = tanh_approx(o.w);
              }
              const int ch = h * 8 + j;
              *reinterpret_cast<float4*>(sf + row * 64 + ((ch ^ (row & 15)) << 2)) = o;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(WS_EPI_WARPS * 32) : "memory");
            const int col0 = (ob0 + b) * 64;
            float* og = reinterpret_cast<float*>(G.out);
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
              const int r = ew * 16 + rr;
              const long long gr = mt * WS_M + r;
              if (gr < a.M) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const int c = hh * 32 + lane;
                  if (col0 + c < a.n_out)
                    og[gr * a.ldy + col0 + c] = sf[r * 64 + (((c >> 2) ^ (r & 15)) << 2) + (c & 3)];
                }
              }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(WS_EPI_WARPS * 32) : "memory");
          }
        }
        ob0 += nb;
      }
    }
    if (!OUT_F32 && issuer) bulk_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();              // the peer is done with this CTA's shared memory / barriers
  if (warp == 0) {
    if (TWO) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

// fp32 x[:, :, cols] gather (nf/flows.py:230) -> bf16 A image [m_tiles][KB][128][64], zero padded.
// One thread per 16-byte chunk of the image: coalesced 16-byte stores.
__global__ void __launch_bounds__(256)
pack_a_img_kernel(const float* __restrict__ x, unsigned char* __restrict__ img, long long N, long long m_tiles,
                  int size, int dim, int n_cols, int c0, int c1, int c2, int c3, int KB, int fmt) {
  const int per_row = size * n_cols;
  const long long total = m_tiles * KB * 128 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int slot = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const long long blk = i >> 10;
    const int kb = (int)(blk % KB);
    const long long mt = blk / KB;
    const int ch = slot ^ (r & 7);
    const long long row = mt * 128 + r;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (row < N) {
      const float* xr = x + row * (long long)(size * dim);
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = kb * 64 + ch * 8 + e;
        float v = 0.f;
        if (k < per_row) {
          const int s = k / n_cols, ci = k - s * n_cols;
          const int col = ci == 0 ? c0 : ci == 1 ? c1 : ci == 2 ? c2 : c3;
          v = xr[s * dim + col];
        }
        f[e] = v;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) w[e] = pack_img2(f[2 * e], f[2 * e + 1], fmt);
    }
    *reinterpret_cast<uint4*>(img + i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// fp32 weight matrix -> bf16 W image (per N tile: [KB][64*nb][64], SWIZZLE_128B): one launch per layer
// and parameter version (training repacks after every optimizer step).  Logical element (out row o,
// input column k); `transposed` reads W[k, o] (dgrad operands), `pad_rows` / `pad_k` map a padded
// 24-per-feature index to the 23-per-feature source (spline parameter rows), everything outside the
// source is zero.  One thread per 16-byte chunk.
struct PackWArgs {
  const float* w;
  unsigned char* img;
  long long ld;
  int n_src_rows, n_src_cols;      // shape of the stored matrix
  int KB, n_tiles, transposed, pad_rows, pad_k, fmt;
  int nb[WS_MAX_TILES];
};
__device__ __forceinline__ int unpad24(int i) { return (i % 24 == 23) ? -1 : (i / 24) * 23 + (i % 24); }

__global__ void __launch_bounds__(256) pack_w_img_kernel(const __grid_constant__ PackWArgs a) {
  long long total = 0;
  for (int t = 0; t < a.n_tiles; ++t) total += (long long)a.KB * a.nb[t] * 64 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long rem = i;
    int t = 0, row0 = 0;
    for (; t < a.n_tiles; ++t) {
      const long long sz = (long long)a.KB * a.nb[t] * 64 * 8;
      if (rem < sz) break;
      rem -= sz;
      row0 += a.nb[t] * 64;
    }
    const int rows_t = a.nb[t] * 64;
    const int slot = (int)(rem & 7);
    const int r = (int)((rem >> 3) % rows_t);
    const int kb = (int)((rem >> 3) / rows_t);
    int o = row0 + r;
    if (a.pad_rows) o = unpad24(o);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int k = kb * 64 + (slot ^ (r & 7)) * 8 + e;
      if (a.pad_k) k = unpad24(k);
      float v = 0.f;
      if (o >= 0 && k >= 0) {
        const int sr = a.transposed ? k : o, sc = a.transposed ? o : k;
        if (sr < a.n_src_rows && sc < a.n_src_cols) v = a.w[(long long)sr * a.ld + sc];
      }
      f[e] = v;
    }
    uint4 u;
    u.x = pack_img2(f[0], f[1], a.fmt);
    u.y = pack_img2(f[2], f[3], a.fmt);
    u.z = pack_img2(f[4], f[5], a.fmt);
    u.w = pack_img2(f[6], f[7], a.fmt);
    *reinterpret_cast<uint4*>(a.img + i * 16) = u;
  }
}

// NSF_AR conditioner inputs (reference nf/flows.py:172-173, :186): trig_transform(x[:, :i]) for every
// dimension i is a PREFIX of one interleaved feature row [cos(pi x_0/B), sin(pi x_0/B), cos(pi x_1/B), ...]:
// one image of 2*dim columns serves all dim-1 conditioners (conditioner i reads its first ceil(2i/64)
// K blocks; its weight image is zero beyond column 2i, which masks the later dimensions).
__global__ void __launch_bounds__(256)
nsf_ar_pack_kernel(const float* __restrict__ x, unsigned char* __restrict__ img, long long N, long long m_tiles,
                   int dim, int KB, float pi_f32, float B, int fmt) {
  const long long total = m_tiles * KB * 128 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int slot = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const long long blk = i >> 10;
    const int kb = (int)(blk % KB);
    const long long row = (blk / KB) * 128 + r;
    const int k0 = kb * 64 + (slot ^ (r & 7)) * 8;            // 8 columns = 4 dimensions (cos, sin) each
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = (k0 >> 1) + e;
      float c = 0.f, sn = 0.f;
      if (row < N && j < dim) {
        const float ang = pi_f32 * x[row * dim + j] / B;      // torch.tensor(pi) * x / B
        sincosf(ang, &sn, &c);
      }
      f[2 * e] = c;
      f[2 * e + 1] = sn;
    }
    uint4 u;
    u.x = pack_img2(f[0], f[1], fmt);
    u.y = pack_img2(f[2], f[3], fmt);
    u.z = pack_img2(f[4], f[5], fmt);
    u.w = pack_img2(f[6], f[7], fmt);
    *reinterpret_cast<uint4*>(img + i * 16) = u;
  }
}

// bf16 image [m_tiles][KB][128][64] -> row-major bf16 [M, ld] (first ncols columns): hands the saved
// activations / gradient images to the weight-gradient GEMMs.  One thread per 16-byte chunk.
__global__ void __launch_bounds__(256)
unpack_img_rows_kernel(const unsigned char* __restrict__ img, unsigned char* __restrict__ rows, long long M,
                       int KB, int ncols, long long ld) {
  const long long m_tiles = (M + 127) / 128;
  const long long total = m_tiles * KB * 128 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int slot = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const long long blk = i >> 10;
    const int kb = (int)(blk % KB);
    const long long row = (blk / KB) * 128 + r;
    const int col = kb * 64 + (slot ^ (r & 7)) * 8;
    if (row < M && col + 8 <= ncols)
      *reinterpret_cast<uint4*>(rows + (row * ld + col) * 2) = *reinterpret_cast<const uint4*>(img + i * 16);
  }
}

// g[:, s*dim + mask[j]] += dxc[:, s*n_mask + j]: the conditioner-input gradient joins the direct path
__global__ void __launch_bounds__(256)
scatter_add_cols_kernel(float* __restrict__ g, const float* __restrict__ dxc, long long N, int size, int dim,
                        int n_mask, int m0, int m1, int m2, int m3) {
  const int per_row = size * n_mask;
  const long long total = N * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / per_row;
    const int rem = (int)(i - row * per_row);
    const int sgrp = rem / n_mask, j = rem - sgrp * n_mask;
    const int col = j == 0 ? m0 : j == 1 ? m1 : j == 2 ? m2 : m3;
    g[row * (long long)(size * dim) + sgrp * dim + col] += dxc[i];
  }
}

RqsConsts make_rqs_consts(int K, float B);   // rqs_coupling.cu

static int fill_rqs_geometry(WsArgs& a, int size, int dim, const int32_t* mask, int n_mask, const char* who) {
  NFK_REQUIRE(size >= 1 && dim >= 2 && dim <= 4 && mask && n_mask >= 1 && n_mask < dim,
              "%s: need 2 <= dim <= 4 and 1 <= n_mask < dim", who);
  a.dim = dim;
  a.n_mask = n_mask;
  bool used[4] = {false, false, false, false};
  for (int j = 0; j < n_mask; ++j) {
    NFK_REQUIRE(mask[j] >= 0 && mask[j] < dim && !used[mask[j]], "%s: bad mask column %d", who, mask[j]);
    used[mask[j]] = true;
    a.mask[j] = mask[j];
  }
  for (int c = 0, u = 0; c < dim; ++c)
    if (!used[c]) a.unm[u++] = c;
  a.n_feat = size * (dim - n_mask);
  a.n_tiles = (a.n_feat + WS_TF - 1) / WS_TF;
  NFK_REQUIRE(a.n_tiles <= WS_MAX_TILES, "%s: %d transformed features exceed %d", who, a.n_feat,
              WS_MAX_TILES * WS_TF);
  for (int t = 0; t < a.n_tiles; ++t) a.nb[t] = WS_TF * WS_PC / 64;
  return NFK_OK;
}

static int g_ws_last_clusters = 0;     // co-resident CTA pairs of the last pair-mode launch (diagnostic)
static int g_ws_pair_mode = -1;          // -1 automatic, 0 never, 1 always (nfk_set_gemm_ws_pair_mode)

template <int EPI, int MODE, bool INVERSE, bool PAIRS = false>
static int launch_ws(const WsArgs& a, cudaStream_t st) {
  constexpr size_t smem = ws_smem<EPI>();
  constexpr unsigned threads = (2 + WsCfg<EPI>::EPI_WARPS) * 32;
  const long long cap = sm_count();
  // CTA pairs (cta_group::2) once there are enough M tiles to keep every SM pair busy
  const long long ng = a.n_groups > 0 ? a.n_groups : 1;
  const bool two = g_ws_pair_mode < 0 ? (a.m_tiles >= 2 * cap) : (g_ws_pair_mode != 0 && a.m_tiles >= 2);
  cudaError_t e;
  if (two) {
    auto kern = gemm_ws_kernel<EPI, MODE, INVERSE, PAIRS, true>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      const long long pairs = ((a.m_tiles + 1) / 2) * ng;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)cap);
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      // persistent kernel: launch exactly as many CTA pairs as can be co-resident (a pair needs both
      // SMs of one TPC; not every SM of the chip has a free partner), never a second wave
      static int max_clusters[4] = {0, 0, 0, 0};
      if (max_clusters[EPI] == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = (int)(cap / 2);
        max_clusters[EPI] = n;
      }
      g_ws_last_clusters = max_clusters[EPI];
      const long long gp = pairs < max_clusters[EPI] ? pairs : max_clusters[EPI];
      cfg.gridDim = dim3((unsigned)(2 * gp));
      e = cudaLaunchKernelEx(&cfg, kern, a);
      if (e != cudaSuccess) {
        set_error("gemm_ws: cluster launch failed: %s", cudaGetErrorString(e));
        return NFK_ECUDA;
      }
    }
  } else {
    auto kern = gemm_ws_kernel<EPI, MODE, INVERSE, PAIRS, false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      const unsigned grid = (unsigned)(a.m_tiles * ng < cap ? a.m_tiles * ng : cap);
      kern<<<grid, threads, smem, st>>>(a);
    }
  }
  if (e != cudaSuccess) {
    set_error("gemm_ws: cannot set %zu B dynamic shared memory: %s", smem, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  count_launch();
  return check_launch("gemm_ws");
}

}  // namespace nfk

using namespace nfk;

extern "C" {

int nfk_gemm_ws_rows_per_tile(void) { return WS_M; }

int nfk_gemm_ws_last_clusters(void) { return g_ws_last_clusters; }

int nfk_set_gemm_ws_pair_mode(int mode) {
  NFK_REQUIRE(mode >= -1 && mode <= 1, "set_gemm_ws_pair_mode: -1 automatic, 0 single CTA, 1 CTA pairs");
  g_ws_pair_mode = mode;
  return NFK_OK;
}

#define NFK_REQUIRE_FMT(who) \
  NFK_REQUIRE(fmt == NFK_IMG_BF16 || fmt == NFK_IMG_F16, who ": fmt must be NFK_IMG_BF16 (0) or NFK_IMG_F16 (1)")

int nfk_pack_a_img(const float* x, void* img, int64_t N, int size, int dim, const int32_t* cols, int n_cols,
                   int KB, int fmt, void* stream) {
  NFK_REQUIRE(N >= 0 && size > 0 && dim > 0 && KB > 0, "pack_a_img: bad shape");
  NFK_REQUIRE_FMT("pack_a_img");
  NFK_REQUIRE(cols && n_cols >= 1 && n_cols <= 4, "pack_a_img: 1..4 columns supported");
  NFK_REQUIRE((long long)size * n_cols <= (long long)KB * 64, "pack_a_img: %d columns do not fit %d K blocks",
              size * n_cols, KB);
  int c[4] = {0, 0, 0, 0};
  for (int i = 0; i < n_cols; ++i) {
    NFK_REQUIRE(cols[i] >= 0 && cols[i] < dim, "pack_a_img: column %d outside [0,%d)", cols[i], dim);
    c[i] = cols[i];
  }
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && img, "pack_a_img: null device pointer");
  NFK_REQUIRE((reinterpret_cast<uintptr_t>(img) & 15) == 0, "pack_a_img: image must be 16-byte aligned");
  const long long m_tiles = (N + WS_M - 1) / WS_M;
  const long long total = m_tiles * KB * 128 * 8;
  long long grid = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  pack_a_img_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      x, reinterpret_cast<unsigned char*>(img), N, m_tiles, size, dim, n_cols, c[0], c[1], c[2], c[3], KB, fmt);
  count_launch();
  return check_launch("pack_a_img");
}

int nfk_gemm_ws(const void* a_img, const void* w_img, const float* bias, void* out, int64_t M, int KB,
                int kmma_last, const int32_t* tile_blocks, int n_tiles, int act, int out_f32, int n_out,
                int64_t ldy, const void* aux, int fmt, void* stream) {
  NFK_REQUIRE(M >= 0 && KB > 0, "gemm_ws: bad shape M=%lld KB=%d", (long long)M, KB);
  NFK_REQUIRE_FMT("gemm_ws");
  NFK_REQUIRE(act != 2 || fmt == NFK_IMG_BF16, "gemm_ws: act 2 (tanh backward) reads and writes bf16 images");
  NFK_REQUIRE(kmma_last >= 1 && kmma_last <= 4, "gemm_ws: kmma_last must be 1..4");
  NFK_REQUIRE(tile_blocks && n_tiles >= 1 && n_tiles <= WS_MAX_TILES, "gemm_ws: 1..%d N tiles", WS_MAX_TILES);
  NFK_REQUIRE(act >= 0 && act <= 2, "gemm_ws: act must be 0 (identity), 1 (tanh) or 2 (tanh backward)");
  NFK_REQUIRE(act != 2 || (aux && !out_f32 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0),
              "gemm_ws: act 2 needs the 16-byte aligned activation image `aux` and a bf16 image output");
  WsArgs a{};
  int ob = 0;
  for (int t = 0; t < n_tiles; ++t) {
    NFK_REQUIRE(tile_blocks[t] >= 1 && tile_blocks[t] <= 4, "gemm_ws: N tile of %d blocks (1..4 allowed)",
                tile_blocks[t]);
    a.nb[t] = tile_blocks[t];
    ob += tile_blocks[t];
  }
  if (out_f32) NFK_REQUIRE(n_out >= 1 && n_out <= ob * 64 && ldy >= n_out, "gemm_ws: bad fp32 output shape");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(a_img && w_img && bias && out, "gemm_ws: null device pointer");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(a_img) | reinterpret_cast<uintptr_t>(w_img) |
                reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "gemm_ws: pointers must be 16-byte aligned");
  a.a_img = reinterpret_cast<const unsigned char*>(a_img);
  a.w_img = reinterpret_cast<const unsigned char*>(w_img);
  a.bias = bias;
  a.out = out;
  a.m_tiles = (M + WS_M - 1) / WS_M;
  a.M = M;
  a.ldy = ldy;
  a.KB = KB;
  a.kmma_last = kmma_last;
  a.n_tiles = n_tiles;
  a.n_out = n_out;
  a.act = act;
  a.fmt = fmt;
  a.aux = reinterpret_cast<const unsigned char*>(aux);
  cudaStream_t st = (cudaStream_t)stream;
  return out_f32 ? launch_ws<EPI_F32_ROWS, 0, false>(a, st) : launch_ws<EPI_BF16_IMG, 0, false>(a, st);
}

int nfk_gemm_ws_grouped(const void* groups_dev, int n_groups, int64_t M, const int32_t* tile_blocks, int n_tiles,
                        int act, int out_f32, int n_out, int64_t ldy, int fmt, void* stream) {
  NFK_REQUIRE(M >= 0 && n_groups >= 1, "gemm_ws_grouped: bad shape");
  NFK_REQUIRE_FMT("gemm_ws_grouped");
  NFK_REQUIRE(tile_blocks && n_tiles >= 1 && n_tiles <= WS_MAX_TILES, "gemm_ws_grouped: 1..%d N tiles", WS_MAX_TILES);
  NFK_REQUIRE(act == 0 || act == 1, "gemm_ws_grouped: act must be 0 (identity) or 1 (tanh)");
  WsArgs a{};
  int ob = 0;
  for (int t = 0; t < n_tiles; ++t) {
    NFK_REQUIRE(tile_blocks[t] >= 1 && tile_blocks[t] <= 4, "gemm_ws_grouped: N tile of %d blocks", tile_blocks[t]);
    a.nb[t] = tile_blocks[t];
    ob += tile_blocks[t];
  }
  if (out_f32) NFK_REQUIRE(n_out >= 1 && n_out <= ob * 64 && ldy >= n_out, "gemm_ws_grouped: bad fp32 output shape");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(groups_dev && (reinterpret_cast<uintptr_t>(groups_dev) & 7) == 0, "gemm_ws_grouped: bad group table");
  a.groups = reinterpret_cast<const WsGroup*>(groups_dev);
  a.n_groups = n_groups;
  a.m_tiles = (M + WS_M - 1) / WS_M;
  a.M = M;
  a.ldy = ldy;
  a.n_tiles = n_tiles;
  a.n_out = n_out;
  a.act = act;
  a.fmt = fmt;
  a.KB = 1;
  a.kmma_last = 4;
  cudaStream_t st = (cudaStream_t)stream;
  return out_f32 ? launch_ws<EPI_F32_ROWS, 0, false>(a, st) : launch_ws<EPI_BF16_IMG, 0, false>(a, st);
}

int nfk_gemm_ws_group_bytes(void) { return (int)sizeof(WsGroup); }

int nfk_nsf_ar_pack(const float* x, void* img, int64_t N, int dim, float B, int fmt, void* stream) {
  NFK_REQUIRE(N >= 0 && dim >= 2 && B > 0.f, "nsf_ar_pack: need dim >= 2 and B > 0");
  NFK_REQUIRE_FMT("nsf_ar_pack");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && img && (reinterpret_cast<uintptr_t>(img) & 15) == 0, "nsf_ar_pack: bad pointer");
  const long long m_tiles = (N + 127) / 128;
  const int KB = (2 * dim + 63) / 64;
  const long long total = m_tiles * KB * 1024;
  long long grid = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  nsf_ar_pack_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<unsigned char*>(img), N,
                                                                      m_tiles, dim, KB, 3.14159274101257324f, B, fmt);
  count_launch();
  return check_launch("nsf_ar_pack");
}

int nfk_gemm_ws_rqs(const void* a_img, const void* w_img, const float* bias, const float* x, float* out,
                    float* logdet, int64_t M, int KB, int kmma_last, int size, int dim, const int32_t* mask,
                    int n_mask, float B, int inverse, int accumulate, int arith, int fmt, float* dbg_params,
                    int8_t* dbg_bins, void* stream) {
  NFK_REQUIRE(M >= 0 && KB > 0, "gemm_ws_rqs: bad shape M=%lld KB=%d", (long long)M, KB);
  NFK_REQUIRE_FMT("gemm_ws_rqs");
  NFK_REQUIRE(kmma_last >= 1 && kmma_last <= 4, "gemm_ws_rqs: kmma_last must be 1..4");
  NFK_REQUIRE(arith >= NFK_ARITH_EXACT && arith <= NFK_ARITH_FAST, "gemm_ws_rqs: bad arith %d", arith);
  NFK_REQUIRE(B > 0.f, "gemm_ws_rqs: tail bound must be positive");
  WsArgs a{};
  if (int rc = fill_rqs_geometry(a, size, dim, mask, n_mask, "gemm_ws_rqs")) return rc;
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(a_img && w_img && bias && x && out && logdet, "gemm_ws_rqs: null device pointer");
  NFK_REQUIRE(x != out, "gemm_ws_rqs: out must not alias x");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(a_img) | reinterpret_cast<uintptr_t>(w_img) |
                reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
              "gemm_ws_rqs: operand images must be 16-byte aligned");
  a.a_img = reinterpret_cast<const unsigned char*>(a_img);
  a.w_img = reinterpret_cast<const unsigned char*>(w_img);
  a.bias = bias;
  a.out = out;
  a.m_tiles = (M + WS_M - 1) / WS_M;
  a.M = M;
  a.KB = KB;
  a.kmma_last = kmma_last;
  a.x = x;
  a.logdet = logdet;
  a.accumulate = accumulate;
  a.fmt = fmt;
  a.dbg_params = dbg_params;
  a.dbg_bins = reinterpret_cast<signed char*>(dbg_bins);
  a.c = make_rqs_consts(8, B);
  cudaStream_t st = (cudaStream_t)stream;
  const bool inv = inverse != 0;
  const bool pairs = (dim == 2);
#define NFK_WS_RQS(MODE)                                                                             \
  do {                                                                                               \
    if (pairs)                                                                                       \
      return inv ? launch_ws<EPI_RQS, MODE, true, true>(a, st) : launch_ws<EPI_RQS, MODE, false, true>(a, st);   \
    return inv ? launch_ws<EPI_RQS, MODE, true, false>(a, st) : launch_ws<EPI_RQS, MODE, false, false>(a, st);   \
  } while (0)
  if (arith == NFK_ARITH_EXACT) NFK_WS_RQS(NFK_ARITH_EXACT);
  if (arith == NFK_ARITH_HYBRID) NFK_WS_RQS(NFK_ARITH_HYBRID);
  NFK_WS_RQS(NFK_ARITH_FAST);
#undef NFK_WS_RQS
}

int nfk_gemm_ws_rqs_bwd(const void* a_img, const void* w_img, const float* bias, const float* x,
                        const float* grad_out, const float* grad_logdet, float grad_logdet_const, float* grad_x,
                        void* grad_params_img, int64_t M, int KB, int kmma_last, int size, int dim,
                        const int32_t* mask, int n_mask, float B, int inverse, void* stream) {
  NFK_REQUIRE(M >= 0 && KB > 0, "gemm_ws_rqs_bwd: bad shape M=%lld KB=%d", (long long)M, KB);
  NFK_REQUIRE(kmma_last >= 1 && kmma_last <= 4, "gemm_ws_rqs_bwd: kmma_last must be 1..4");
  NFK_REQUIRE(B > 0.f, "gemm_ws_rqs_bwd: tail bound must be positive");
  WsArgs a{};
  if (int rc = fill_rqs_geometry(a, size, dim, mask, n_mask, "gemm_ws_rqs_bwd")) return rc;
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(a_img && w_img && bias && x && grad_out && grad_x && grad_params_img,
              "gemm_ws_rqs_bwd: null device pointer");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(a_img) | reinterpret_cast<uintptr_t>(w_img) |
                reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(grad_params_img)) & 15) == 0,
              "gemm_ws_rqs_bwd: operand images must be 16-byte aligned");
  a.a_img = reinterpret_cast<const unsigned char*>(a_img);
  a.w_img = reinterpret_cast<const unsigned char*>(w_img);
  a.bias = bias;
  a.out = grad_params_img;
  a.m_tiles = (M + WS_M - 1) / WS_M;
  a.M = M;
  a.KB = KB;
  a.kmma_last = kmma_last;
  a.x = x;
  a.gout = grad_out;
  a.gin = grad_x;
  a.gld = grad_logdet;
  a.gld_const = grad_logdet_const;
  a.c = make_rqs_consts(8, B);
  cudaStream_t st = (cudaStream_t)stream;
  return inverse ? launch_ws<EPI_RQS_BWD, 0, true>(a, st) : launch_ws<EPI_RQS_BWD, 0, false>(a, st);
}

int nfk_pack_w_img(const float* W, int64_t ld, int n_src_rows, int n_src_cols, void* img, int KB,
                   const int32_t* tile_blocks, int n_tiles, int transposed, int pad_rows, int pad_k, int fmt,
                   void* stream) {
  NFK_REQUIRE(n_src_rows > 0 && n_src_cols > 0 && ld >= n_src_cols && KB > 0, "pack_w_img: bad shape");
  NFK_REQUIRE_FMT("pack_w_img");
  NFK_REQUIRE(tile_blocks && n_tiles >= 1 && n_tiles <= WS_MAX_TILES, "pack_w_img: 1..%d N tiles", WS_MAX_TILES);
  NFK_REQUIRE(W && img && (reinterpret_cast<uintptr_t>(img) & 15) == 0, "pack_w_img: bad pointer");
  PackWArgs a{};
  a.w = W;
  a.img = reinterpret_cast<unsigned char*>(img);
  a.ld = ld;
  a.n_src_rows = n_src_rows;
  a.n_src_cols = n_src_cols;
  a.KB = KB;
  a.n_tiles = n_tiles;
  a.transposed = transposed;
  a.pad_rows = pad_rows;
  a.pad_k = pad_k;
  a.fmt = fmt;
  long long total = 0;
  for (int t = 0; t < n_tiles; ++t) {
    NFK_REQUIRE(tile_blocks[t] >= 1 && tile_blocks[t] <= 4, "pack_w_img: N tile of %d blocks", tile_blocks[t]);
    a.nb[t] = tile_blocks[t];
    total += (long long)KB * tile_blocks[t] * 64 * 8;
  }
  long long grid = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  pack_w_img_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  return check_launch("pack_w_img");
}

int nfk_unpack_img_rows(const void* img, void* rows, int64_t M, int KB, int ncols, int64_t ld, void* stream) {
  NFK_REQUIRE(M >= 0 && KB > 0 && ncols > 0 && ncols <= KB * 64 && ncols % 8 == 0 && ld >= ncols && ld % 8 == 0,
              "unpack_img_rows: need ncols, ld multiples of 8 with ncols <= 64*KB <= ... and ld >= ncols");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(img && rows, "unpack_img_rows: null device pointer");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(rows)) & 15) == 0,
              "unpack_img_rows: pointers must be 16-byte aligned");
  const long long total = ((M + 127) / 128) * KB * 128 * 8;
  long long grid = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  unpack_img_rows_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned char*>(img), reinterpret_cast<unsigned char*>(rows), M, KB, ncols, ld);
  count_launch();
  return check_launch("unpack_img_rows");
}

int nfk_scatter_add_cols(float* g, const float* dxc, int64_t N, int size, int dim, const int32_t* cols, int n_cols,
                         void* stream) {
  NFK_REQUIRE(N >= 0 && size > 0 && dim > 0, "scatter_add_cols: bad shape");
  NFK_REQUIRE(cols && n_cols >= 1 && n_cols <= 4, "scatter_add_cols: 1..4 columns supported");
  int c[4] = {0, 0, 0, 0};
  for (int i = 0; i < n_cols; ++i) {
    NFK_REQUIRE(cols[i] >= 0 && cols[i] < dim, "scatter_add_cols: column %d outside [0,%d)", cols[i], dim);
    c[i] = cols[i];
  }
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(g && dxc, "scatter_add_cols: null device pointer");
  const long long total = (long long)N * size * n_cols;
  long long grid = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  scatter_add_cols_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(g, dxc, N, size, dim, n_cols, c[0], c[1],
                                                                          c[2], c[3]);
  count_launch();
  return check_launch("scatter_add_cols");
}

}  // extern "C"
