// One NSF coupling layer in ONE kernel: conditioner MLP on the tensor cores + RQS transform as
// the epilogue of its last GEMM.  Replaces NSF_CL.forward/inverse entirely (reference
// nf/flows.py:227-253 incl. FCNN nf/flows.py:26-35 and nf/utils.py:20-152) for the headline
// geometry: size = 32, dim = 2 (32 conditioning + 32 transformed columns), K = 8 bins,
// hidden width <= 128.  The [N, 32, 23] spline-parameter tensor (2,944 B/row each way) never
// exists in HBM: per row the kernel reads 256 B of x and writes 256 B of z plus the log-det.
//
// Persistent CTA (16 warps), 128 rows per tile:
// Operands are IEEE fp16, not bf16: the activations are tanh outputs in (-1, 1) and the weights O(1), so
// fp16's range is ample (the conditioning inputs are clamped to +-65504) and its 11-bit significand carries
// 8x less quantisation noise than bf16's 8 bits at the same tensor-core rate (measured on the timed
// configuration, per layer vs the fp32 oracle: spline parameters 3.5e-3 -> see DESIGN.md section 6).
//   P1  x tile (TMA bulk, double buffered) -> conditioning columns -> fp16 -> A operand in the
//       K-major SWIZZLE_128B layout (shared memory)
//   P2  GEMM1  D12[128x128] = A1 W1^T      (tcgen05.mma, M=128 N=128 K=16 x4, accumulators in TMEM)
//   P3  epilogue 1: tcgen05.ld -> +b1 -> tanh -> fp16 -> A operand (same buffer)
//   P4  GEMM2  D12 = A2 W2^T ; P5 epilogue 2 -> A3
//   P6  for each of 8 chunks of 4 features (4 x 24 = 96 accumulator columns, the 23 parameters
//       of a feature padded to 24): GEMM3 chunk into one of two TMEM buffers while the previous
//       chunk's epilogue runs: thread (row, feature) pulls its 24 raw parameters out of TMEM,
//       adds b3 and evaluates bin search + spline + log|det| in registers (rqs_math.cuh), writing
//       the output pair in place into the x tile.  W3 chunks stream L2 -> shared memory through a
//       3-stage TMA bulk ring (weights are pre-swizzled into their shared-memory image once).
//   P7  row log-det = sum over the 32 features (4 partial sums per row), output tile -> TMA
//       bulk store.
#include "rqs_math.cuh"
#include "tc05.cuh"

namespace nfk {

constexpr int FU_ROWS = 128;
constexpr int FU_EPI_WARPS = 16;
constexpr int FU_THREADS = (FU_EPI_WARPS + 1) * 32;   // 16 epilogue warps + 1 control warp
constexpr int FU_HP = 128;        // padded hidden width
constexpr int FU_K1P = 64;        // padded conditioner input width
constexpr int FU_NF = 32;         // transformed features
constexpr int FU_PC = 24;         // accumulator columns per feature (23 used)
constexpr int FU_CF = 4;          // features per GEMM3 chunk
constexpr int FU_NC = FU_CF * FU_PC;      // 96 columns per chunk
constexpr int FU_NCHUNK = FU_NF / FU_CF;  // 8
constexpr int FU_W3STAGES = 3;

constexpr uint32_t FU_W1_BYTES = FU_HP * 128;                    // [128 x 64] fp16
constexpr uint32_t FU_W2_BYTES = 2 * FU_HP * 128;                // 2 K blocks of [128 x 64]
constexpr uint32_t FU_A_BYTES = 2 * FU_ROWS * 128;               // 2 K blocks of [128 x 64]
constexpr uint32_t FU_W3C_BYTES = 2 * FU_NC * 128;               // 2 K blocks of [96 x 64]
constexpr int FU_XLD = 68;        // padded x-tile row stride in floats (272 B): the 32 rows a warp
                                  // touches at one column then fall on 8 bank groups, not 1
constexpr uint32_t FU_XROW_BYTES = 64 * 4;
constexpr uint32_t FU_X_BYTES = FU_ROWS * FU_XLD * 4;            // padded fp32 [128 x 64] tile

struct FusedArgs {
  const float* x;
  float* out;
  float* logdet;
  const void* w1_img;     // FU_W1_BYTES, pre-swizzled
  const void* w2_img;     // FU_W2_BYTES
  const void* w3_img;     // FU_NCHUNK * FU_W3C_BYTES
  const float* b1;        // [128]
  const float* b2;        // [128]
  const float* b3;        // [32*24]
  long long n_tiles;
  int cond_first;         // 1: conditioning column is column 0 of each pair (mask = [0])
  int accumulate;
  float* dbg_params;      // optional [N][32][24]: the raw spline parameters (accumulator + b3) each element saw
  signed char* dbg_bins;  // optional [N][32]: the bin each element used (-1 = identity tail)
  long long* trace;       // optional [17][32] clock64 stamps of CTA 0, tile 2: control warp, then the 16 epilogue warps (tools/trace_fused.py)
  RqsConsts c;
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp roles: warps 0..15 are epilogue warps (TMEM lane quadrant q = warp % 4, column slice /
// feature-in-chunk = warp / 4); warp 16 is the control warp: it issues every tcgen05.mma and
// every TMA bulk copy and never waits on arithmetic, so tensor work and weight streaming run
// ahead of the epilogues.  All hand-offs are mbarriers; there is no CTA-wide barrier in the loop.
// DBG: test-hook instantiation that also writes the raw parameters and bins of every element (the production
// instantiation carries none of that code: even a never-taken uniform branch cost 1-2 % of the launch).
template <int MODE, bool INVERSE, bool DBG = false>
__global__ void __launch_bounds__(FU_THREADS, 1)
nsf_pairs_fused_kernel(const __grid_constant__ FusedArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned char* sW1 = sm;
  unsigned char* sW2 = sW1 + FU_W1_BYTES;
  unsigned char* sA = sW2 + FU_W2_BYTES;
  unsigned char* sW3 = sA + FU_A_BYTES;
  float* sX = reinterpret_cast<float*>(sW3 + FU_W3STAGES * FU_W3C_BYTES);      // 2 padded tiles
  float* sB1 = sX + 2 * FU_ROWS * FU_XLD;
  float* sB2 = sB1 + FU_HP;
  float* sB3 = sB2 + FU_HP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB3 + FU_NF * FU_PC);
  uint64_t* bar_w = bars;            // W1 + W2 resident                 (1)
  uint64_t* bar_x = bars + 1;        // [2] x tile landed                (16 warps + tx)
  uint64_t* bar_w3 = bars + 3;       // [3] W3 chunk landed              (1 + tx)
  uint64_t* bar_mma = bars + 6;      // GEMM1 / GEMM2 done               (1, tcgen05.commit)
  uint64_t* bar_d3f = bars + 7;      // [2] GEMM3 chunk done             (1, tcgen05.commit)
  uint64_t* bar_d3e = bars + 9;      // [2] D3 buffer drained            (16 warps)
  uint64_t* bar_a = bars + 11;       // A operand written                (16 warps)
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned first = blockIdx.x, stride = gridDim.x;
  const unsigned n_tiles = (unsigned)a.n_tiles;
  const unsigned my_tiles = (n_tiles > first) ? (n_tiles - first + stride - 1) / stride : 0;
  const unsigned total_chunks = my_tiles * FU_NCHUNK;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(&bars[i], (i == 1 || i == 2) ? FU_EPI_WARPS : 1);
    mbar_init(&bar_d3e[0], FU_EPI_WARPS);
    mbar_init(&bar_d3e[1], FU_EPI_WARPS);
    mbar_init(bar_a, FU_EPI_WARPS);
    fence_barrier_init();
  }
  for (int i = tid; i < FU_HP; i += FU_THREADS) {
    sB1[i] = a.b1[i];
    sB2[i] = a.b2[i];
  }
  for (int i = tid; i < FU_NF * FU_PC; i += FU_THREADS) sB3[i] = a.b3[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD12 = tmem;                       // 128 columns

  if (warp == FU_EPI_WARPS) {
    // =============================== control warp ===============================
    const uint32_t idesc12 = make_idesc_f16(FU_ROWS, FU_HP);
    const uint32_t idesc3 = make_idesc_f16(FU_ROWS, FU_NC);
    const uint32_t aA = smem_u32(sA), aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
    auto issue_w3 = [&](unsigned g) {        // lane 0
      const int s = g % FU_W3STAGES;
      const int c = g % FU_NCHUNK;
      mbar_expect_tx(&bar_w3[s], FU_W3C_BYTES);
      bulk_g2s(sW3 + s * FU_W3C_BYTES,
               reinterpret_cast<const unsigned char*>(a.w3_img) + (size_t)c * FU_W3C_BYTES, FU_W3C_BYTES,
               &bar_w3[s]);
    };
    if (my_tiles) {
      if (lane == 0) {
        mbar_expect_tx(bar_w, FU_W1_BYTES + FU_W2_BYTES);
        bulk_g2s(sW1, a.w1_img, FU_W1_BYTES, bar_w);
        bulk_g2s(sW2, a.w2_img, FU_W2_BYTES, bar_w);
        for (unsigned g = 0; g < FU_W3STAGES && g < total_chunks; ++g) issue_w3(g);
      }
      if (lane == 0) mbar_wait_idle(bar_w, 0);
      __syncwarp();
    }
    unsigned g = 0, na = 0;          // running GEMM3 chunk counter, running bar_a phase counter
    // every lane runs the same control flow (waits included); only the tcgen05 / TMA issue
    // instructions sit under elect.sync, so the compiler keeps their operands uniform
    for (unsigned it = 0; it < my_tiles; ++it) {
      const bool tr = a.trace && blockIdx.x == 0 && it == 2 && lane == 0;
      int ts = 0;
#define NFK_STAMP(base) do { if (tr) a.trace[(base) + ts++] = clock64(); } while (0)
      // GEMM1: D12 = A1 W1^T
      NFK_STAMP(0);
      mbar_wait_idle(bar_a, na++ & 1);
      NFK_STAMP(0);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FU_K1P / 16; ++k)
          umma_bf16(tD12, make_desc_sw128(aA + k * 32), make_desc_sw128(aW1 + k * 32), idesc12, k ? 1u : 0u);
        umma_commit(bar_mma);
      }
      __syncwarp();
      NFK_STAMP(0);
      // GEMM2: D12 = A2 W2^T
      NFK_STAMP(0);
      mbar_wait_idle(bar_a, na++ & 1);
      NFK_STAMP(0);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tD12, make_desc_sw128(aA + kb * (FU_ROWS * 128) + k * 32),
                      make_desc_sw128(aW2 + kb * (FU_HP * 128) + k * 32), idesc12, (kb | k) ? 1u : 0u);
        umma_commit(bar_mma);
      }
      __syncwarp();
      NFK_STAMP(0);
      // GEMM3 chunks: D3[g & 1] = A3 W3chunk^T
      mbar_wait_idle(bar_a, na++ & 1);
      NFK_STAMP(0);
      for (int c = 0; c < FU_NCHUNK; ++c, ++g) {
        const int s = g % FU_W3STAGES;
        mbar_wait_idle(&bar_w3[s], (g / FU_W3STAGES) & 1);
        if (g >= 2) mbar_wait_idle(&bar_d3e[g & 1], ((g >> 1) + 1) & 1);   // chunk g-2 left this buffer
        tc_fence_after();
        const uint32_t d = tmem + 128 + (g & 1) * FU_NC;
        const uint32_t bbase = aW3 + s * FU_W3C_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d, make_desc_sw128(aA + kb * (FU_ROWS * 128) + k * 32),
                        make_desc_sw128(bbase + kb * (FU_NC * 128) + k * 32), idesc3, (kb | k) ? 1u : 0u);
          umma_commit(&bar_d3f[g & 1]);
        }
        __syncwarp();
        NFK_STAMP(0);
        if (g >= 1 && g + 2 < total_chunks) {
          // chunk g-1 has been consumed by the tensor core: refill its ring slot with chunk g+2
          mbar_wait_idle(&bar_d3f[(g - 1) & 1], ((g - 1) >> 1) & 1);
          if (lane == 0) issue_w3(g + 2);
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int q = warp & 3;            // TMEM lane quadrant
    const int slice = warp >> 2;       // column slice (epilogues 1/2) / feature-in-chunk (epilogue 3)
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;

    auto hidden_epilogue = [&](const float* bias) {
      uint32_t v[32];
      tmem_ld32(tD12 + lane_sel + slice * 32, v);
      tmem_ld_wait();
      unsigned char* dst = sA + (slice >> 1) * (FU_ROWS * 128) + row * 128;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#ifdef NFK_ABLATE_NOTANH   // timing experiment only
          f[j] = 0.5f * (__uint_as_float(v[t * 8 + j]) + bias[slice * 32 + t * 8 + j]);
#else
          f[j] = tanh_approx(__uint_as_float(v[t * 8 + j]) + bias[slice * 32 + t * 8 + j]);
#endif
        uint4 u;
        u.x = pack_f16x2(f[0], f[1]);
        u.y = pack_f16x2(f[2], f[3]);
        u.z = pack_f16x2(f[4], f[5]);
        u.w = pack_f16x2(f[6], f[7]);
        const int ch = (slice & 1) * 4 + t;
        *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = u;
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(bar_a);
    };

    // lanes 0..7 of every epilogue warp move 8 rows of the x / output tile each (256-byte TMA bulk
    // copies into the padded rows): the copy work is spread over 16 warps instead of serialising
    // on the control warp, and a lane reloads exactly the row it stored, so buffer reuse needs
    // no cross-thread hand-off beyond its own bulk-group wait.
    const int myrow = q * 32 + slice * 8 + (lane & 7);
    auto load_rows = [&](unsigned it2) {
      const size_t tile2 = first + (size_t)it2 * stride;
      const int s = it2 & 1;
      if (lane == 0) mbar_expect_tx(&bar_x[s], 8 * FU_XROW_BYTES);       // arrive (1 of 16) + expect
      __syncwarp();
      if (lane < 8)
        bulk_g2s(sX + (s * FU_ROWS + myrow) * FU_XLD, a.x + (tile2 * FU_ROWS + myrow) * 64, FU_XROW_BYTES,
                 &bar_x[s]);
    };
    if (my_tiles) load_rows(0);
    if (my_tiles > 1) load_rows(1);

    for (unsigned it = 0; it < my_tiles; ++it) {
      const size_t tile = first + (size_t)it * stride;
      float* xs = sX + (it & 1) * FU_ROWS * FU_XLD;
      float ld_old = 0.f;
      if (lane < 8 && a.accumulate) ld_old = __ldg(a.logdet + tile * FU_ROWS + myrow);   // consumed in P7
      const bool tr = a.trace && blockIdx.x == 0 && it == 2 && lane == 0;     // every epilogue warp: [32 + 32 * warp ...]
      int ts = 0;
      // ---- P1: conditioning columns -> A1 (K block 0; columns 32..63 are zero padding)
      NFK_STAMP(32 + 32 * warp);
      mbar_wait(&bar_x[it & 1], (it >> 1) & 1);
      NFK_STAMP(32 + 32 * warp);
      for (int i = tid; i < FU_ROWS * 8; i += FU_EPI_WARPS * 32) {
        const int r = i >> 3, ch = i & 7;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (ch < 4) {
          const float* xr = xs + r * FU_XLD + (a.cond_first ? 0 : 1);
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = xr[2 * (ch * 8 + j)];
          u.x = pack_f16x2_sat(f[0], f[1]);                  // conditioning inputs: saturate to the fp16 range
          u.y = pack_f16x2_sat(f[2], f[3]);
          u.z = pack_f16x2_sat(f[4], f[5]);
          u.w = pack_f16x2_sat(f[6], f[7]);
        }
        *reinterpret_cast<uint4*>(sA + r * 128 + ((ch ^ (r & 7)) << 4)) = u;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(bar_a);
      NFK_STAMP(32 + 32 * warp);
      // prefetch the next tile into the other buffer (this lane stored that row one tile ago)
      if (it >= 1 && it + 1 < my_tiles) {
        if (lane < 8) bulk_wait_read<0>();
        __syncwarp();
        load_rows(it + 1);
      }
      // ---- P3 / P5: hidden-layer epilogues
      mbar_wait(bar_mma, 0);
      NFK_STAMP(32 + 32 * warp);
      tc_fence_after();
      hidden_epilogue(sB1);
      NFK_STAMP(32 + 32 * warp);
      mbar_wait(bar_mma, 1);
      NFK_STAMP(32 + 32 * warp);
      tc_fence_after();
      hidden_epilogue(sB2);
      NFK_STAMP(32 + 32 * warp);
      // ---- P6: spline transform as the epilogue of the GEMM3 chunks
      float lad_acc = 0.f;
#pragma unroll 1
      for (int c = 0; c < FU_NCHUNK; ++c) {
        mbar_wait(&bar_d3f[c & 1], (c >> 1) & 1);
        NFK_STAMP(32 + 32 * warp);
        tc_fence_after();
        const int f = c * FU_CF + slice;                     // feature of this thread
        uint32_t v[24];
        const uint32_t t = tmem + 128 + (c & 1) * FU_NC + lane_sel + slice * FU_PC;
        tmem_ld16(t, v);
        tmem_ld8(t + 16, v + 16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cnt(&bar_d3e[c & 1]);      // the buffer may be overwritten
        float2* pr = reinterpret_cast<float2*>(xs + row * FU_XLD + 2 * f);
        const float2 xc = *pr;
#ifdef NFK_ABLATE_NOELEM   // timing experiment only (tools/ubench/README.md): everything but the spline math
        RqsOut o;
        o.y = (a.cond_first ? xc.y : xc.x) + __uint_as_float(v[0] ^ v[7] ^ v[15] ^ v[22]);
        o.lad = __uint_as_float(v[1] ^ v[9] ^ v[17]);
#else
        // FAST decides the bin lazily on the exact chain next to a knot (FIXBINS), so in every arithmetic the
        // bin is the one nf/utils.py:20-25 finds on the parameters this kernel computed
#ifdef NFK_ABLATE_NOFIX       // timing experiment only: without the lazy exact-bin decision
        const RqsOut o = rqs_element<MODE, 8, INVERSE, true, false>(RegParams{v, sB3 + f * FU_PC},
                                                                     a.cond_first ? xc.y : xc.x, a.c);
#else
        const RqsOut o = rqs_element<MODE, 8, INVERSE, true, true>(RegParams{v, sB3 + f * FU_PC},
                                                                    a.cond_first ? xc.y : xc.x, a.c);
#endif
#ifndef NFK_ABLATE_NODBG
        if constexpr (DBG) {           // test hook: what this element computed from
          const size_t e = ((size_t)tile * FU_ROWS + row) * FU_NF + f;
#pragma unroll
          for (int i = 0; i < FU_PC; ++i)
            a.dbg_params[e * FU_PC + i] = i < 23 ? __uint_as_float(v[i]) + sB3[f * FU_PC + i] : 0.f;
          a.dbg_bins[e] = (signed char)o.bin;
        }
#endif
#endif
        *pr = make_float2(a.cond_first ? xc.x : xc.y, o.y);  // (conditioning, transformed): Q5
        lad_acc += o.lad;
        NFK_STAMP(32 + 32 * warp);
      }
      // ---- P7: the four warps of this lane quadrant own these 32 rows: exchange the partial
      // log-dets through the row padding, then lanes 0..7 finish 8 rows each
      xs[row * FU_XLD + 64 + slice] = lad_acc;
      fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      if (lane < 8) {
        const float* pr = xs + myrow * FU_XLD + 64;
        const float t = (pr[0] + pr[1]) + (pr[2] + pr[3]);                   // flows.py:238
        a.logdet[tile * FU_ROWS + myrow] = a.accumulate ? ld_old + t : t;
        bulk_s2g(a.out + (tile * FU_ROWS + myrow) * 64, xs + myrow * FU_XLD, FU_XROW_BYTES);
        bulk_commit();
      }
    }
    if (lane < 8) bulk_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

constexpr size_t FU_SMEM = FU_W1_BYTES + FU_W2_BYTES + FU_A_BYTES + FU_W3STAGES * FU_W3C_BYTES + 2 * FU_X_BYTES +
                           (2 * FU_HP + FU_NF * FU_PC) * 4 + 16 * 8 + 1024;
static_assert(FU_SMEM <= 227 * 1024, "fused layer kernel exceeds the 227 KB shared-memory limit");

RqsConsts make_rqs_consts(int K, float B);   // rqs_coupling.cu

template <int MODE, bool INVERSE, bool DBG = false>
static int launch_fused(const FusedArgs& a, cudaStream_t st) {
  auto kern = nsf_pairs_fused_kernel<MODE, INVERSE, DBG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM);
  if (e != cudaSuccess) {
    set_error("nsf_fused: cannot set %zu B dynamic shared memory: %s", FU_SMEM, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  const long long cap = sm_count();
  const long long grid = a.n_tiles < cap ? a.n_tiles : cap;
  kern<<<(unsigned)grid, FU_THREADS, FU_SMEM, st>>>(a);
  count_launch();
  return check_launch("nsf_pairs_fused");
}

}  // namespace nfk

using namespace nfk;

static long long* g_fused_trace = nullptr;

extern "C" {

/* test hook: device buffer of 17 x 32 int64 that receives clock64 stamps of CTA 0 / tile 2 */
int nfk_set_fused_trace(void* dev_buf) {
  g_fused_trace = reinterpret_cast<long long*>(dev_buf);
  return NFK_OK;
}

int nfk_nsf_fused_rows_per_tile(void) { return FU_ROWS; }

int nfk_nsf_pairs_fused(const float* x, float* out, float* logdet, const void* w1_img, const void* w2_img,
                        const void* w3_img, const float* b1, const float* b2, const float* b3, int64_t N,
                        int mask_col, float B, int inverse, int accumulate, int arith, float* dbg_params,
                        int8_t* dbg_bins, void* stream) {
  NFK_REQUIRE(N >= 0 && N % FU_ROWS == 0, "nsf_pairs_fused: N must be a multiple of %d (got %lld)", FU_ROWS,
              (long long)N);
  NFK_REQUIRE(mask_col == 0 || mask_col == 1, "nsf_pairs_fused: mask column must be 0 or 1");
  NFK_REQUIRE(arith >= NFK_ARITH_EXACT && arith <= NFK_ARITH_FAST, "nsf_pairs_fused: bad arith %d", arith);
  NFK_REQUIRE(B > 0.f, "nsf_pairs_fused: tail bound must be positive");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && out && logdet && w1_img && w2_img && w3_img && b1 && b2 && b3,
              "nsf_pairs_fused: null device pointer");
  NFK_REQUIRE(x != out, "nsf_pairs_fused: out must not alias x");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                reinterpret_cast<uintptr_t>(w1_img) | reinterpret_cast<uintptr_t>(w2_img) |
                reinterpret_cast<uintptr_t>(w3_img)) & 15) == 0,
              "nsf_pairs_fused: pointers must be 16-byte aligned");
  FusedArgs a{};
  a.x = x;
  a.out = out;
  a.logdet = logdet;
  a.w1_img = w1_img;
  a.w2_img = w2_img;
  a.w3_img = w3_img;
  a.b1 = b1;
  a.b2 = b2;
  a.b3 = b3;
  a.n_tiles = N / FU_ROWS;
  a.cond_first = (mask_col == 0);
  a.accumulate = accumulate;
  a.dbg_params = dbg_params;
  a.dbg_bins = reinterpret_cast<signed char*>(dbg_bins);
  a.trace = g_fused_trace;
  a.c = make_rqs_consts(8, B);
  cudaStream_t st = (cudaStream_t)stream;
  const bool inv = inverse != 0;
  if (dbg_params || dbg_bins) {
    NFK_REQUIRE(dbg_params && dbg_bins, "nsf_pairs_fused: dbg_params and dbg_bins go together");
    if (arith == NFK_ARITH_EXACT)
      return inv ? launch_fused<NFK_ARITH_EXACT, true, true>(a, st) : launch_fused<NFK_ARITH_EXACT, false, true>(a, st);
    if (arith == NFK_ARITH_HYBRID)
      return inv ? launch_fused<NFK_ARITH_HYBRID, true, true>(a, st) : launch_fused<NFK_ARITH_HYBRID, false, true>(a, st);
    return inv ? launch_fused<NFK_ARITH_FAST, true, true>(a, st) : launch_fused<NFK_ARITH_FAST, false, true>(a, st);
  }
  if (arith == NFK_ARITH_EXACT)
    return inv ? launch_fused<NFK_ARITH_EXACT, true>(a, st) : launch_fused<NFK_ARITH_EXACT, false>(a, st);
  if (arith == NFK_ARITH_HYBRID)
    return inv ? launch_fused<NFK_ARITH_HYBRID, true>(a, st) : launch_fused<NFK_ARITH_HYBRID, false>(a, st);
  return inv ? launch_fused<NFK_ARITH_FAST, true>(a, st) : launch_fused<NFK_ARITH_FAST, false>(a, st);
}

}  // extern "C"
