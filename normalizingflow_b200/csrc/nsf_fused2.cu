// One NSF coupling layer in ONE kernel, second generation: the layer's three phases run on THREE sets of
// warps that overlap in time.  Replaces NSF_CL.forward/inverse (reference nf/flows.py:227-253 incl. FCNN
// nf/flows.py:26-35 and nf/utils.py:20-152) for size = 32, dim = 2, K = 8, hidden width <= 128.
//
// First generation (nsf_fused.cu): 16 epilogue warps do everything -- build the A operand, the two hidden-layer
// epilogues (tanh), then the 8 spline chunks -- so the tensor-core round trips and tanh epilogues of a tile
// (5-6 k clk of a 26 k clk tile) sit serially in front of the spline work.  Here:
//
//   warps  0..15  SPLINE   thread (row, feature-in-chunk): tcgen05.ld of its 24 raw parameters, bin search + spline
//                          + log|det| in registers (rqs_math.cuh), output pair written in place into the x tile;
//                          row log-det; lanes 0..7 of every warp move 8 rows of the tile (TMA bulk in / out)
//   warps 16..19  HIDDEN   thread = row: conditioning columns -> fp16 A operand, hidden-layer epilogues
//                          (tcgen05.ld -> + bias -> tanh -> fp16 -> next A operand), ONE TILE AHEAD of the spline warps
//   warp  20      MMA      every tcgen05.mma: GEMM1 / GEMM2 of tile i+1 slotted between the GEMM3 chunks of tile i
//   warp  21      TMA      weight ring: W2 K blocks and W3 chunks stream L2 -> shared memory (3 x 24 KB stages)
//   (warps 22, 23 only complete the last warpgroup for setmaxnreg)
//
// Registers are redistributed per warpgroup (setmaxnreg): the kernel launches at 80 per thread, the spline
// warpgroups grow to 96, the hidden warpgroup shrinks to 40 and the MMA / TMA warpgroup to 48.
// Two A buffers alternate between tiles, so GEMM3 of tile i reads h2(i) while the hidden warps build h1 / h2 of
// tile i+1; accumulators: D12 (128 TMEM columns) for the hidden GEMMs, two 96-column buffers for the chunks.
//
// SPLIT (fp32-class conditioner, conditioner="fp32x3"): every operand is kept as an fp16 pair hi + lo
// (hi = fp16(v), lo = fp16(v - hi): 22 significant bits) and every product runs as three MMAs
// hi*hi + lo*hi + hi*lo accumulated in fp32 in TMEM.  The second A buffer then holds the lo halves, so the
// hidden warps cannot run ahead (the layer's phases serialise as in the first generation).
#include <type_traits>

#include "rqs_math.cuh"
#include "tc05.cuh"

namespace nfk {

constexpr int F2_ROWS = 128;
constexpr int F2_SPLINE_WARPS = 16;
constexpr int F2_HIDDEN_WARPS = 4;
constexpr int F2_WARP_MMA = 20, F2_WARP_TMA = 21;
constexpr int F2_THREADS = 24 * 32;
constexpr int F2_HP = 128, F2_NF = 32, F2_PC = 24, F2_CF = 4;
constexpr int F2_NC = F2_CF * F2_PC;              // 96 accumulator columns per chunk
constexpr int F2_NCHUNK = F2_NF / F2_CF;          // 8
constexpr int F2_STAGES = 3;
constexpr uint32_t F2_W1_BYTES = F2_HP * 128;                  // [128 x 64] fp16 (SPLIT: columns 0-31 hi, 32-63 lo)
constexpr uint32_t F2_KB_BYTES = F2_ROWS * 128;                // one 128 x 64 fp16 K block (16 KB)
constexpr uint32_t F2_A_BYTES = 2 * F2_KB_BYTES;               // [128 x 128] fp16
constexpr uint32_t F2_W3C_BYTES = 2 * F2_NC * 128;             // one W3 chunk [96 x 128] fp16 (24 KB)
constexpr uint32_t F2_STAGE_BYTES = F2_W3C_BYTES;
constexpr int F2_XLD = 68;
constexpr uint32_t F2_XROW_BYTES = 64 * 4;
constexpr uint32_t F2_X_BYTES = F2_ROWS * F2_XLD * 4;
// setmaxnreg moves registers inside the pool the CTA was LAUNCHED with (24 warps x 32 x 80 = 61,440), so the three
// budgets must fit it: 16 x 96 + 4 x 40 + 4 x 48 = 1,888 per lane <= 24 x 80 = 1,920.
constexpr int F2_REG_LAUNCH = 80, F2_REG_SPLINE = 96, F2_REG_HIDDEN = 40, F2_REG_CTRL = 48;
static_assert(16 * F2_REG_SPLINE + 4 * F2_REG_HIDDEN + 4 * F2_REG_CTRL <= 24 * F2_REG_LAUNCH, "register budgets exceed the launch pool");

struct Fused2Args {
  const float* x;
  float* out;
  float* logdet;
  const unsigned char* w1_img;   // F2_W1_BYTES
  const unsigned char* w2_img;   // plain: [kb0][kb1] 16 KB blocks; SPLIT: [kb0 hi][kb0 lo][kb1 hi][kb1 lo]
  const unsigned char* w3_img;   // plain: 8 chunks of 24 KB; SPLIT: per chunk [hi][lo]
  const float* b1;               // [128]
  const float* b2;               // [128]
  const float* b3;               // [32*24]
  long long n_tiles;
  int cond_first;
  int accumulate;
  float* dbg_params;
  signed char* dbg_bins;
  int* flag_in;                  // [n_tiles] or null: tile t of x is complete when flag_in[t] >= flag_epoch (per-tile dependency on the producing launch)
  int* flag_out;                 // [n_tiles] or null: set to flag_epoch when tile t of out / logdet is complete
  int flag_epoch;                // >= 1; the caller zeroes the flags before the first chain and raises the epoch from chain to chain
  RqsConsts c;
};

__device__ __forceinline__ int f2_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void f2_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ bool f2_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
template <int N>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
// fp32 pair -> fp16x2 hi word and fp16x2 lo word (lo = fp16(v - float(hi)))
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split_f16x2_sat(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = fminf(fmaxf(a, -65504.f), 65504.f);          // NaN passes through (fmaxf/fminf return the other operand: keep it)
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  split_f16x2(a, b, hi, lo);
}

// tanh for the split-operand (fp32-class) configuration: MUFU.TANH carries ~2^-11 relative error, which would
// throw away what the hi + lo operands buy; 1 - 2 / (1 + e^(2x)) on ex2.approx + rcp.approx is good to ~3e-7
// absolute everywhere (saturates correctly: e^(2x) = inf -> 1, 0 -> -1) at two MUFU operations.
__device__ __forceinline__ float tanh_2mufu(float x) {
  const float t = ex2_approx(x * (2.f * LOG2E));
  return fmaf(-2.f, rcp_approx(1.f + t), 1.f);
}

// 24 raw parameters of one feature held in registers, bias already added in place (add_bias24)
struct RegVals {
  static constexpr bool in_registers = true;
  const uint32_t* v;
  __device__ __forceinline__ float operator()(int i) const { return __uint_as_float(v[i]); }
  __device__ __forceinline__ float dyn(int base, int i) const {
    uint32_t r = v[16];
#pragma unroll
    for (int j = 1; j < 7; ++j)
      if (i == j) r = v[16 + j];
    return __uint_as_float(r);
  }
};
// v[0..23] += b[0..23]: six 16-byte shared-memory loads (every lane reads the same address: broadcast) and
// twelve packed fp32x2 adds instead of 24 scalar loads + 24 scalar adds per element
__device__ __forceinline__ void add_bias24(uint32_t* v, const float* b) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 bb = *reinterpret_cast<const float4*>(b + 4 * i);
    float lo, hi;
    unpk2(add2(pk2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), pk2(bb.x, bb.y)), lo, hi);
    v[4 * i] = __float_as_uint(lo);
    v[4 * i + 1] = __float_as_uint(hi);
    unpk2(add2(pk2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), pk2(bb.z, bb.w)), lo, hi);
    v[4 * i + 2] = __float_as_uint(lo);
    v[4 * i + 3] = __float_as_uint(hi);
  }
}

// Per-tile order of the MMA warp's work and of the ring pieces (both warps walk the same list):
//   chunk 0 .. S1, [GEMM1 of the next tile], chunk S1+1 .. S2, [GEMM2 of the next tile: W2 pieces], chunk S2+1 .. 7
// plain: S1 = 2, S2 = 4 (the hidden warps run one tile ahead); SPLIT: S1 = S2 = 7 (phases serialise).
template <bool SPLIT>
struct F2Sched {
  static constexpr int S1 = SPLIT ? 7 : 2;
  static constexpr int S2 = SPLIT ? 7 : 4;
  static constexpr int NA = SPLIT ? 1 : 2;           // A buffers that alternate between tiles
  static constexpr int W2_PIECES = SPLIT ? 4 : 2;    // 16 KB ring pieces of W2 per tile
  static constexpr int W3_PIECES = SPLIT ? 2 : 1;    // 24 KB ring pieces per chunk
};

// CHAIN: the tile-flag variant (nfk_nsf_pairs_fused2_chain) is its own instantiation, so the plain kernel -- the one the
// headline times -- carries none of its code (the flag tests inside the chunk loop cost it 4 % when they were
// run-time branches).
template <int MODE, bool INVERSE, bool SPLIT, bool DBG, bool CHAIN = false>
__global__ void __launch_bounds__(F2_THREADS, 1)
nsf_fused2_kernel(const __grid_constant__ Fused2Args a) {
  using S = F2Sched<SPLIT>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned char* sW1 = sm;
  unsigned char* sA = sW1 + F2_W1_BYTES;                         // two 32 KB operand buffers
  unsigned char* sRing = sA + 2 * F2_A_BYTES;
  float* sX = reinterpret_cast<float*>(sRing + F2_STAGES * F2_STAGE_BYTES);     // 2 padded tiles
  float* sB1 = sX + 2 * F2_ROWS * F2_XLD;
  float* sB2 = sB1 + F2_HP;
  float* sB3 = sB2 + F2_HP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB3 + F2_NF * F2_PC);
  uint64_t* bar_w1 = bars;              // W1 resident                                   (1 + tx)
  uint64_t* bar_x = bars + 1;           // [2] x tile landed                             (16 spline warps + tx)
  uint64_t* bar_full = bars + 3;        // [3] ring piece landed                         (1 + tx)
  uint64_t* bar_empty = bars + 6;       // [3] ring piece consumed                       (1, tcgen05.commit)
  uint64_t* bar_aready = bars + 9;      // [2] A operand written, 3 phases per tile      (4 hidden warps)
  uint64_t* bar_afree = bars + 11;      // [2] every MMA reading this A buffer is done   (1, tcgen05.commit)
  uint64_t* bar_d12 = bars + 13;        // hidden GEMM done, 2 phases per tile           (1, tcgen05.commit)
  uint64_t* bar_d3f = bars + 14;        // [2] GEMM3 chunk done                          (1, tcgen05.commit)
  uint64_t* bar_d3e = bars + 16;        // [2] chunk accumulator drained                 (16 spline warps)
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned first = blockIdx.x, stride = gridDim.x;
  const unsigned n_tiles = (unsigned)a.n_tiles;
  const unsigned my_tiles = (n_tiles > first) ? (n_tiles - first + stride - 1) / stride : 0;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    mbar_init(bar_w1, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_x[i], F2_SPLINE_WARPS);
      mbar_init(&bar_aready[i], F2_HIDDEN_WARPS);
      mbar_init(&bar_afree[i], 1);
      mbar_init(&bar_d3f[i], 1);
      mbar_init(&bar_d3e[i], F2_SPLINE_WARPS);
    }
    for (int i = 0; i < F2_STAGES; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_d12, 1);
    fence_barrier_init();
  }
  // Programmatic dependent launch: this grid may have been scheduled while the previous kernel of the stream
  // (the previous layer) was still draining -- its CTAs start on each SM as soon as that SM's CTA exits, with
  // barrier init and TMEM allocation already done.  Nothing in global memory is touched before this wait; the
  // trigger lets the NEXT layer's grid do the same behind this one.
  // With tile flags (a chain of layer launches over the same rows) the dependency is per TILE instead: this grid does
  // not wait for the previous launch as a whole, its CTAs start on the SMs that launch has already left and take a
  // tile when its flag is set (rows are independent).  Everything else read here (weight images, biases) was complete
  // before the first launch of the chain, which waits for the stream.
  if (!CHAIN || a.flag_in == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  for (int i = tid; i < F2_HP; i += F2_THREADS) {
    sB1[i] = a.b1[i];
    sB2[i] = a.b2[i];
  }
  for (int i = tid; i < F2_NF * F2_PC; i += F2_THREADS) sB3[i] = a.b3[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD12 = tmem;                       // 128 columns; chunk buffers at +128 and +224

  if (warp >= F2_WARP_MMA) {
    // ======================= MMA issuer / TMA producer (+ two idle warps) =======================
    reg_dec<F2_REG_CTRL>();
    if (warp == F2_WARP_TMA) {
      // ---------------- ring producer: the pieces in exactly the order the MMA warp consumes them
      uint32_t s = 0, ph = 0;
      auto put = [&](const unsigned char* src, uint32_t bytes) {
        mbar_wait_idle(&bar_empty[s], ph ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&bar_full[s], bytes);
          bulk_g2s(sRing + s * F2_STAGE_BYTES, src, bytes, &bar_full[s]);
        }
        __syncwarp();
        if (++s == F2_STAGES) {
          s = 0;
          ph ^= 1;
        }
      };
      auto put_w2 = [&]() {
        for (int p = 0; p < S::W2_PIECES; ++p) put(a.w2_img + (size_t)p * F2_KB_BYTES, F2_KB_BYTES);
      };
      auto put_chunk = [&](int c) {
        for (int p = 0; p < S::W3_PIECES; ++p)
          put(a.w3_img + ((size_t)c * S::W3_PIECES + p) * F2_W3C_BYTES, F2_W3C_BYTES);
      };
      if (my_tiles) {
        if (lane == 0) {
          mbar_expect_tx(bar_w1, F2_W1_BYTES);
          bulk_g2s(sW1, a.w1_img, F2_W1_BYTES, bar_w1);
        }
        __syncwarp();
        put_w2();                                                  // GEMM2 of the first tile
      }
      for (unsigned it = 0; it < my_tiles; ++it) {
        const bool next = it + 1 < my_tiles;
        for (int c = 0; c < F2_NCHUNK; ++c) {
          put_chunk(c);
          if (c == S::S2 && next) put_w2();
        }
      }
    } else if (warp == F2_WARP_MMA) {
      // ---------------- MMA issuer
      const uint32_t idesc12 = make_idesc_f16(F2_ROWS, F2_HP);
      const uint32_t idesc3 = make_idesc_f16(F2_ROWS, F2_NC);
      const uint32_t aA = smem_u32(sA), aW1 = smem_u32(sW1), aRing = smem_u32(sRing);
      uint32_t s = 0, ph = 0;            // ring consumer state
      uint32_t n_ar0 = 0, n_ar1 = 0;     // phases consumed on bar_aready[0 / 1]
      auto ar_parity = [&](uint32_t ab) -> uint32_t { return (ab ? n_ar1++ : n_ar0++) & 1u; };
      uint32_t g = 0;                    // running chunk counter
      auto ring_next = [&]() {
        if (++s == F2_STAGES) {
          s = 0;
          ph ^= 1;
        }
      };
      // GEMM1 + GEMM2 issue of hidden tile `t`: the two waits on bar_aready are what the hidden warps signal
      auto gemm1 = [&](unsigned t) {
        const uint32_t ab = SPLIT ? 0u : (t & 1u);
        const uint32_t abase = aA + ab * F2_A_BYTES;
        mbar_wait_idle(&bar_aready[ab], ar_parity(ab));
        tc_fence_after();
        if (f2_elect_one()) {
          if (SPLIT) {
            // A1 / W1: K columns 0-31 = hi, 32-63 = lo (16-column MMA slices 0,1 / 2,3)
            const int ka[6] = {0, 1, 2, 3, 0, 1}, kw[6] = {0, 1, 0, 1, 2, 3};
#pragma unroll
            for (int i = 0; i < 6; ++i)
              umma_bf16(tD12, make_desc_sw128(abase + ka[i] * 32), make_desc_sw128(aW1 + kw[i] * 32), idesc12, i ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < 2; ++k)      // 32 real conditioning columns: K slices 0 and 1
              umma_bf16(tD12, make_desc_sw128(abase + k * 32), make_desc_sw128(aW1 + k * 32), idesc12, k ? 1u : 0u);
          }
          umma_commit(bar_d12);
        }
        __syncwarp();
      };
      auto gemm2 = [&](unsigned t) {
        const uint32_t ab = SPLIT ? 0u : (t & 1u);
        const uint32_t abase = aA + ab * F2_A_BYTES;
        mbar_wait_idle(&bar_aready[ab], ar_parity(ab));
        tc_fence_after();
        for (int p = 0; p < S::W2_PIECES; ++p) {
          const int kb = SPLIT ? (p >> 1) : p;
          const bool lo_piece = SPLIT && (p & 1);
          mbar_wait_idle(&bar_full[s], ph);
          tc_fence_after();
          if (f2_elect_one()) {
            const uint32_t bb = aRing + s * F2_STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tD12, make_desc_sw128(abase + kb * F2_KB_BYTES + k * 32), make_desc_sw128(bb + k * 32), idesc12,
                        (p | k) ? 1u : 0u);
            if (SPLIT && !lo_piece) {
#pragma unroll
              for (int k = 0; k < 4; ++k)     // lo(A) * hi(W)
                umma_bf16(tD12, make_desc_sw128(abase + F2_A_BYTES + kb * F2_KB_BYTES + k * 32),
                          make_desc_sw128(bb + k * 32), idesc12, 1u);
            }
            umma_commit(&bar_empty[s]);
            if (p == S::W2_PIECES - 1) umma_commit(bar_d12);
          }
          __syncwarp();
          ring_next();
        }
      };
      if (my_tiles) {
        mbar_wait_idle(bar_w1, 0);
        gemm1(0);
        gemm2(0);
      }
      for (unsigned it = 0; it < my_tiles; ++it) {
        const bool next = it + 1 < my_tiles;
        const uint32_t ab = SPLIT ? 0u : (it & 1u);
        const uint32_t abase = aA + ab * F2_A_BYTES;
        mbar_wait_idle(&bar_aready[ab], ar_parity(ab));          // h2 of this tile is in the A buffer
        for (int c = 0; c < F2_NCHUNK; ++c, ++g) {
          if (g >= 2) mbar_wait_idle(&bar_d3e[g & 1], ((g >> 1) + 1) & 1);   // chunk g-2 left this accumulator
          const uint32_t d = tmem + 128 + (g & 1) * F2_NC;
          for (int p = 0; p < S::W3_PIECES; ++p) {
            mbar_wait_idle(&bar_full[s], ph);
            tc_fence_after();
            if (f2_elect_one()) {
              const uint32_t bb = aRing + s * F2_STAGE_BYTES;
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(d, make_desc_sw128(abase + kb * F2_KB_BYTES + k * 32),
                            make_desc_sw128(bb + kb * (F2_NC * 128) + k * 32), idesc3, (p | kb | k) ? 1u : 0u);
              if (SPLIT && p == 0) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                  for (int k = 0; k < 4; ++k)   // lo(A) * hi(W)
                    umma_bf16(d, make_desc_sw128(abase + F2_A_BYTES + kb * F2_KB_BYTES + k * 32),
                              make_desc_sw128(bb + kb * (F2_NC * 128) + k * 32), idesc3, 1u);
              }
              umma_commit(&bar_empty[s]);
              if (p == S::W3_PIECES - 1) {
                umma_commit(&bar_d3f[g & 1]);
                if (c == F2_NCHUNK - 1) umma_commit(&bar_afree[ab]);     // nothing reads this A buffer any more
              }
            }
            __syncwarp();
            ring_next();
          }
          if (next && c == S::S1) gemm1(it + 1);
          if (next && c == S::S2) gemm2(it + 1);
        }
      }
    }
  } else if (warp >= F2_SPLINE_WARPS) {
    // =============================== hidden warps (one row per thread) ===============================
    reg_dec<F2_REG_HIDDEN>();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t n_d12 = 0;
    auto hidden_epilogue = [&](const float* bias, unsigned char* dstA) {
#pragma unroll 1
      for (int part = 0; part < 8; ++part) {           // 16 accumulator columns at a time (56-register budget)
        uint32_t v[16];
        tmem_ld16(tD12 + lane_sel + part * 16, v);
        tmem_ld_wait();
        unsigned char* dst = dstA + (part >> 2) * F2_KB_BYTES + row * 128;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float pre = __uint_as_float(v[t * 8 + j]) + bias[part * 16 + t * 8 + j];
            f[j] = SPLIT ? tanh_2mufu(pre) : tanh_approx(pre);
          }
          const int ch = (part & 3) * 2 + t;
          if (SPLIT) {
            uint4 uh, ul;
            split_f16x2(f[0], f[1], uh.x, ul.x);
            split_f16x2(f[2], f[3], uh.y, ul.y);
            split_f16x2(f[4], f[5], uh.z, ul.z);
            split_f16x2(f[6], f[7], uh.w, ul.w);
            *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = uh;
            *reinterpret_cast<uint4*>(dst + F2_A_BYTES + ((ch ^ (row & 7)) << 4)) = ul;
          } else {
            uint4 u;
            u.x = pack_f16x2(f[0], f[1]);
            u.y = pack_f16x2(f[2], f[3]);
            u.z = pack_f16x2(f[4], f[5]);
            u.w = pack_f16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = u;
          }
        }
      }
    };
    for (unsigned t = 0; t < my_tiles; ++t) {
      const uint32_t ab = SPLIT ? 0u : (t & 1u);
      unsigned char* dstA = sA + ab * F2_A_BYTES;
      const float* xs = sX + (t & 1) * F2_ROWS * F2_XLD;
      // the A buffer: every MMA of the tile that used it last (tile t - NA) has completed
      if (t >= (unsigned)S::NA) mbar_wait_idle(&bar_afree[ab], ((t / S::NA) + 1) & 1);
      mbar_wait_idle(&bar_x[t & 1], (t >> 1) & 1);
      // ---- P1: conditioning columns of this row -> A1 (K block 0: 4 chunks hi; SPLIT: 4 more chunks lo)
      {
        const float* xr = xs + row * F2_XLD + (a.cond_first ? 0 : 1);
        unsigned char* dst = dstA + row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = xr[2 * (ch * 8 + j)];
          if (SPLIT) {
            uint4 uh, ul;
            split_f16x2_sat(f[0], f[1], uh.x, ul.x);
            split_f16x2_sat(f[2], f[3], uh.y, ul.y);
            split_f16x2_sat(f[4], f[5], uh.z, ul.z);
            split_f16x2_sat(f[6], f[7], uh.w, ul.w);
            *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = uh;
            *reinterpret_cast<uint4*>(dst + (((ch + 4) ^ (row & 7)) << 4)) = ul;
          } else {
            uint4 u;
            u.x = pack_f16x2_sat(f[0], f[1]);
            u.y = pack_f16x2_sat(f[2], f[3]);
            u.z = pack_f16x2_sat(f[4], f[5]);
            u.w = pack_f16x2_sat(f[6], f[7]);
            *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = u;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_aready[ab]);
      // ---- hidden layer 1
      mbar_wait_idle(bar_d12, n_d12++ & 1);
      tc_fence_after();
      hidden_epilogue(sB1, dstA);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_aready[ab]);
      // ---- hidden layer 2
      mbar_wait_idle(bar_d12, n_d12++ & 1);
      tc_fence_after();
      hidden_epilogue(sB2, dstA);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_aready[ab]);
    }
  } else {
    // =============================== spline warps ===============================
    reg_inc<F2_REG_SPLINE>();
    const int q = warp & 3;            // TMEM lane quadrant
    const int slice = warp >> 2;       // feature-in-chunk
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    // lanes 0..7 of every spline warp move 8 rows of the x / output tile each (256-byte TMA bulk copies into the
    // padded rows); a lane reloads exactly the rows it stored, so buffer reuse needs only its own bulk-group wait
    const int myrow = q * 32 + slice * 8 + (lane & 7);
    auto load_rows = [&](unsigned it2) {
      const size_t tile2 = first + (size_t)it2 * stride;
      const int s = it2 & 1;
      if (CHAIN && a.flag_in != nullptr) {                               // the producing launch has finished this tile
        if (lane == 0)
          while (f2_ld_acquire(a.flag_in + tile2) < a.flag_epoch) __nanosleep(64);
        __syncwarp();
        // ... and the bulk copy below (async proxy) sees what the flag publishes, also when the producer wrote it with
        // ordinary stores (generic proxy): a proxy fence over GLOBAL memory, not just shared
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      if (lane == 0) mbar_expect_tx(&bar_x[s], 8 * F2_XROW_BYTES);       // arrive (1 of 16) + expect
      __syncwarp();
      if (lane < 8)
        bulk_g2s(sX + (s * F2_ROWS + myrow) * F2_XLD, a.x + (tile2 * F2_ROWS + myrow) * 64, F2_XROW_BYTES, &bar_x[s]);
    };
    // tile `t_` of out / logdet is complete: the bulk stores of the 128 row-moving lanes have been performed, every
    // thread's writes are fenced, then one release store of the flag (tile-flag chains only)
    auto publish = [&](size_t t_) {
      if (lane < 8) bulk_wait_all<0>();
      asm volatile("fence.proxy.async;" ::: "memory");       // the rows went out through the async proxy
      __threadfence();
      asm volatile("bar.sync 5, 512;" ::: "memory");
      if (tid == 0) f2_st_release(a.flag_out + t_, a.flag_epoch);
    };
    if (my_tiles) load_rows(0);
    if (my_tiles > 1) load_rows(1);
    uint32_t g = 0;
    for (unsigned it = 0; it < my_tiles; ++it) {
      const size_t tile = first + (size_t)it * stride;
      float* xs = sX + (it & 1) * F2_ROWS * F2_XLD;
      float ld_old = 0.f;
      if (lane < 8 && a.accumulate) ld_old = CHAIN ? __ldcg(a.logdet + tile * F2_ROWS + myrow) : __ldg(a.logdet + tile * F2_ROWS + myrow);
      mbar_wait(&bar_x[it & 1], (it >> 1) & 1);
      float lad_acc = 0.f;
#pragma unroll 1
      for (int c = 0; c < F2_NCHUNK; ++c, ++g) {
        // the previous tile's rows have long left shared memory by now: publish it (deferred so that nobody waits on
        // the bulk stores)
        if (CHAIN && c == 1 && it > 0 && a.flag_out != nullptr) publish(tile - stride);
        mbar_wait_idle(&bar_d3f[g & 1], (g >> 1) & 1, 200);   // suspended probe: a waiting warp leaves the issue port to the other three
        tc_fence_after();
        const int f = c * F2_CF + slice;
        uint32_t v[24];
        const uint32_t t = tmem + 128 + (g & 1) * F2_NC + lane_sel + slice * F2_PC;
        tmem_ld16(t, v);
        tmem_ld8(t + 16, v + 16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_d3e[g & 1]);           // the accumulator may be overwritten
        float2* pr = reinterpret_cast<float2*>(xs + row * F2_XLD + 2 * f);
        const float2 xc = *pr;
        add_bias24(v, sB3 + f * F2_PC);
        const RqsOut o = rqs_element<MODE, 8, INVERSE, true, true>(RegVals{v}, a.cond_first ? xc.y : xc.x, a.c);
        if constexpr (DBG) {
          const size_t e = ((size_t)tile * F2_ROWS + row) * F2_NF + f;
#pragma unroll
          for (int i = 0; i < F2_PC; ++i)
            a.dbg_params[e * F2_PC + i] = i < 23 ? __uint_as_float(v[i]) : 0.f;
          a.dbg_bins[e] = (signed char)o.bin;
        }
        *pr = make_float2(a.cond_first ? xc.x : xc.y, o.y);  // (conditioning, transformed): Q5
        lad_acc += o.lad;
      }
      // ---- row log-det: the four warps of a lane quadrant exchange partial sums through the row padding
      xs[row * F2_XLD + 64 + slice] = lad_acc;
      fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      if (lane < 8) {
        const float* pr = xs + myrow * F2_XLD + 64;
        const float t = (pr[0] + pr[1]) + (pr[2] + pr[3]);                   // flows.py:238
        a.logdet[tile * F2_ROWS + myrow] = a.accumulate ? ld_old + t : t;
        bulk_s2g(a.out + (tile * F2_ROWS + myrow) * 64, xs + myrow * F2_XLD, F2_XROW_BYTES);
        bulk_commit();
      }
      // refill this buffer with tile it + 2 as soon as the store has read it
      if (it + 2 < my_tiles) {
        if (lane < 8) bulk_wait_read<0>();
        __syncwarp();
        load_rows(it + 2);
      }
    }
    if (lane < 8) bulk_wait_all<0>();
    if (CHAIN && my_tiles && a.flag_out != nullptr) publish(first + (size_t)(my_tiles - 1) * stride);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

constexpr size_t F2_SMEM = F2_W1_BYTES + 2 * F2_A_BYTES + F2_STAGES * F2_STAGE_BYTES + 2 * F2_X_BYTES +
                           (2 * F2_HP + F2_NF * F2_PC) * 4 + 32 * 8 + 1024;
static_assert(F2_SMEM <= 227 * 1024, "fused layer kernel (v2) exceeds the 227 KB shared-memory limit");

RqsConsts make_rqs_consts(int K, float B);   // rqs_coupling.cu

static int g_fused2_pdl = 1;      // programmatic dependent launch of consecutive layers (nfk_set_fused2_pdl; tuning / tests)

template <int MODE, bool INVERSE, bool SPLIT, bool DBG, bool CHAIN = false>
static int launch_fused2(const Fused2Args& a, cudaStream_t st) {
  auto kern = nsf_fused2_kernel<MODE, INVERSE, SPLIT, DBG, CHAIN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F2_SMEM);
  if (e != cudaSuccess) {
    set_error("nsf_fused2: cannot set %zu B dynamic shared memory: %s", F2_SMEM, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  const long long cap = sm_count();
  const long long grid = a.n_tiles < cap ? a.n_tiles : cap;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(F2_THREADS);
  cfg.dynamicSmemBytes = F2_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see griddepcontrol.wait in the kernel
  attr[0].val.programmaticStreamSerializationAllowed = g_fused2_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) {
    set_error("nsf_fused2: launch failed: %s", cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  count_launch();
  return check_launch("nsf_fused2");
}

template <bool SPLIT, bool DBG, bool CHAIN = false>
static int dispatch_fused2(const Fused2Args& a, int arith, bool inv, cudaStream_t st) {
  if (arith == NFK_ARITH_EXACT)
    return inv ? launch_fused2<NFK_ARITH_EXACT, true, SPLIT, DBG, CHAIN>(a, st) : launch_fused2<NFK_ARITH_EXACT, false, SPLIT, DBG, CHAIN>(a, st);
  if (arith == NFK_ARITH_HYBRID)
    return inv ? launch_fused2<NFK_ARITH_HYBRID, true, SPLIT, DBG, CHAIN>(a, st) : launch_fused2<NFK_ARITH_HYBRID, false, SPLIT, DBG, CHAIN>(a, st);
  return inv ? launch_fused2<NFK_ARITH_FAST, true, SPLIT, DBG, CHAIN>(a, st) : launch_fused2<NFK_ARITH_FAST, false, SPLIT, DBG, CHAIN>(a, st);
}

}  // namespace nfk

using namespace nfk;

extern "C" int nfk_set_fused2_pdl(int on) {
  g_fused2_pdl = on ? 1 : 0;
  return NFK_OK;
}

extern "C" int nfk_nsf_pairs_fused2(const float* x, float* out, float* logdet, const void* w1_img, const void* w2_img,
                                    const void* w3_img, const float* b1, const float* b2, const float* b3, int64_t N,
                                    int mask_col, float B, int inverse, int accumulate, int arith, int split,
                                    float* dbg_params, int8_t* dbg_bins, void* stream) {
  return nfk_nsf_pairs_fused2_chain(x, out, logdet, w1_img, w2_img, w3_img, b1, b2, b3, N, mask_col, B, inverse, accumulate,
                                    arith, split, dbg_params, dbg_bins, nullptr, nullptr, 1, stream);
}

extern "C" int nfk_nsf_pairs_fused2_chain(const float* x, float* out, float* logdet, const void* w1_img, const void* w2_img,
                                          const void* w3_img, const float* b1, const float* b2, const float* b3, int64_t N,
                                          int mask_col, float B, int inverse, int accumulate, int arith, int split,
                                          float* dbg_params, int8_t* dbg_bins, int32_t* tile_flags_in,
                                          int32_t* tile_flags_out, int tile_flag_epoch, void* stream) {
  NFK_REQUIRE(N >= 0 && N % F2_ROWS == 0, "nsf_pairs_fused2: N must be a multiple of %d (got %lld)", F2_ROWS, (long long)N);
  NFK_REQUIRE(mask_col == 0 || mask_col == 1, "nsf_pairs_fused2: mask column must be 0 or 1");
  NFK_REQUIRE(arith >= NFK_ARITH_EXACT && arith <= NFK_ARITH_FAST, "nsf_pairs_fused2: bad arith %d", arith);
  NFK_REQUIRE(B > 0.f, "nsf_pairs_fused2: tail bound must be positive");
  NFK_REQUIRE((dbg_params == nullptr) == (dbg_bins == nullptr), "nsf_pairs_fused2: dbg_params and dbg_bins go together");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && out && logdet && w1_img && w2_img && w3_img && b1 && b2 && b3, "nsf_pairs_fused2: null device pointer");
  NFK_REQUIRE(x != out, "nsf_pairs_fused2: out must not alias x");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w1_img) |
                reinterpret_cast<uintptr_t>(w2_img) | reinterpret_cast<uintptr_t>(w3_img)) & 15) == 0,
              "nsf_pairs_fused2: pointers must be 16-byte aligned");
  Fused2Args a{};
  a.x = x;
  a.out = out;
  a.logdet = logdet;
  a.w1_img = reinterpret_cast<const unsigned char*>(w1_img);
  a.w2_img = reinterpret_cast<const unsigned char*>(w2_img);
  a.w3_img = reinterpret_cast<const unsigned char*>(w3_img);
  a.b1 = b1;
  a.b2 = b2;
  a.b3 = b3;
  a.n_tiles = N / F2_ROWS;
  a.cond_first = (mask_col == 0);
  a.accumulate = accumulate;
  a.dbg_params = dbg_params;
  a.dbg_bins = reinterpret_cast<signed char*>(dbg_bins);
  a.flag_in = tile_flags_in;
  a.flag_out = tile_flags_out;
  a.flag_epoch = tile_flag_epoch;
  a.c = make_rqs_consts(8, B);
  cudaStream_t st = (cudaStream_t)stream;
  const bool inv = inverse != 0;
  if (tile_flags_in != nullptr || tile_flags_out != nullptr) {
    NFK_REQUIRE(!split && dbg_params == nullptr, "nsf_pairs_fused2_chain: tile flags go with the plain 16-bit kernel only (split = 0, no debug outputs)");
    NFK_REQUIRE(tile_flag_epoch >= 1, "nsf_pairs_fused2_chain: the flag epoch starts at 1");
    return dispatch_fused2<false, false, true>(a, arith, inv, st);
  }
  if (dbg_params) return split ? dispatch_fused2<true, true>(a, arith, inv, st) : dispatch_fused2<false, true>(a, arith, inv, st);
  return split ? dispatch_fused2<true, false>(a, arith, inv, st) : dispatch_fused2<false, false>(a, arith, inv, st);
}
