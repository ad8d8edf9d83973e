// Gradient of one NSF coupling layer w.r.t. its INPUT in ONE kernel (hidden width <= 128, size = 32, dim = 2,
// K = 8): what the flow-preconditioned HMC leapfrog needs at every step (reference nf/hmc.py:34-41 drives
// simulation.integration_step; applications/src/systems.py:308-311, :331-336 compute the force by autograd
// through NSF_CL.forward, nf/flows.py:227-239 + nf/utils.py:27-152).  Given the layer input x, dL/d(out) and
// dL/d(log_det), it returns dL/dx with no weight gradients and nothing saved by the forward:
//
//   recompute   A1 = fp16(x[:, cond]) -> GEMM1 -> h1 = tanh(. + b1) -> GEMM2 -> h2 = tanh(. + b2)        (kept in shared memory)
//   per chunk   GEMM3 chunk (4 features x 24 parameters) -> TMEM -> thread (row, feature): spline ADJOINT in
//               registers (rqs_bwd_math.cuh, both knot chains in packed fp32x2): dL/dx of the transformed column and the
//               pass-through gradient of the conditioning column -> HBM (one 8-byte pair), dL/dparams (24 values) ->
//               bf16 A operand G in shared memory (8 features = 192 columns = 3 K blocks per pair of chunks)
//   per pair    dH2 += G W3_pair        (B operand: W3 transposed, streamed through the ring; accumulates in TMEM)
//   then        dZ2 = dH2 (1 - h2^2) -> GEMM with W2^T -> dH1 ;  dZ1 = dH1 (1 - h1^2) -> GEMM with W1^T -> dXc
//               dL/dx[:, cond] += dXc      (red.global.add after a 32 x 32 transpose per lane quadrant through the idle G
//               buffer; or, with the leapfrog fold, the thread that completes a pair also advances momentum / position)
//
// Chains: with tile flags (flag_in / flag_out / flag_epoch) consecutive launches depend on each other per 128-row tile
// instead of per launch -- see include/nfk.h and the comment at griddepcontrol below.
//
// The forward GEMMs use exactly the fp16 operands and MMA order of nsf_fused2_kernel (the rows of W3 permuted so
// that a feature's width and height logits arrive interleaved), so the recomputed parameters are the forward
// kernel's, bit for bit; the backward GEMMs use bf16 operands (gradient range).
// Warps 0..15 WORK (thread = (row, slice): the adjoint of feature 4 c + slice in chunk c, and a quarter of every
// hidden-layer epilogue: 32 of the 128 columns of its row), warp 16 issues every MMA, warp 17 runs the weight ring.
// The first version of this kernel gave the hidden-layer epilogues to four dedicated warps as nsf_fused2.cu does:
// with nothing to overlap them with (a tile's backward is one dependent chain) they cost 29 k of a tile's 69 k
// clocks; spread over the 16 work warps they cost a quarter of that.
// The activations x / dL/d(out) are read straight from global memory by the thread that needs them (8-byte pairs
// one chunk ahead, the next tile's rows prefetched into L2 by the ring warp), so shared memory holds only
// operands: W1 16 KB + h1 32 + h2 32 + G 48 + ring 72.  TMEM: D12 128 + two chunk buffers 192 + dH2 128 + dXc 32
// = 480 of 512 columns.
#include "rqs_bwd_math.cuh"
#include "tc05.cuh"

namespace nfk {

constexpr int FB_ROWS = 128;
constexpr int FB_WORK_WARPS = 16, FB_WARP_MMA = 16, FB_WARP_TMA = 17;
constexpr int FB_THREADS = 18 * 32;
constexpr int FB_HP = 128, FB_NF = 32, FB_PC = 24, FB_CF = 4;
constexpr int FB_NC = FB_CF * FB_PC;             // 96
constexpr int FB_NCHUNK = FB_NF / FB_CF;         // 8
constexpr int FB_NPAIR = FB_NCHUNK / 2;          // 4 pairs of chunks = 4 x 192 gradient columns
constexpr int FB_STAGES = 3;
constexpr uint32_t FB_KB_BYTES = FB_ROWS * 128;                 // one 128 x 64 16-bit K block
constexpr uint32_t FB_W1_BYTES = FB_KB_BYTES;
constexpr uint32_t FB_A_BYTES = 2 * FB_KB_BYTES;
constexpr uint32_t FB_G_BYTES = 3 * FB_KB_BYTES;
constexpr uint32_t FB_W3C_BYTES = 2 * FB_NC * 128;              // 24 KB
constexpr uint32_t FB_STAGE_BYTES = FB_W3C_BYTES;
constexpr uint32_t FB_W1T_BYTES = 32 * 128;                     // one K block of W1^T: [32 rows x 64]
constexpr uint32_t FB_T_D12 = 0, FB_T_D3 = 128, FB_T_DH2 = 320, FB_T_DX = 448;

struct FusedBwdArgs {
  const float* x;         // [N, 64] layer input
  const float* gout;      // [N, 64] dL/d(layer output), reference column order (conditioning, transformed)
  const float* gld;       // [N] dL/dlogdet, or null: gld_const for every row
  float gld_const;
  float gout_scale;       // dL/d(out) = gout_scale * gout (the prior's -1/var when gout is z itself)
  float* gin;             // [N, 64] dL/d(layer input)
  const unsigned char* w1_img;    // forward images as nfk_nsf_pairs_fused2 (fp16) ...
  const unsigned char* w2_img;
  const unsigned char* w3_img;    // ... with every feature's 24 rows in the order (w0, h0, w1, h1, ..., w7, h7, d0..d6, 0)
  const unsigned char* w3t_img;   // [4 pairs][3 K blocks][128 x 64] bf16: (n = hidden unit, k = parameter index in the pair, same order)
  const unsigned char* w2t_img;   // [2 K blocks][128 x 64] bf16: (n = input unit, k = output unit)
  const unsigned char* w1t_img;   // [2 K blocks][32 x 64] bf16: (n = conditioning feature, k = hidden unit)
  const float* b1;
  const float* b2;
  const float* b3;                // [32][24] in the order of w3_img's rows
  long long n_tiles;
  int cond_first;
  int* flag_in;           // [n_tiles] or null: tile t may start when flag_in[t] >= flag_epoch (set by the launch that produces gout)
  int* flag_out;          // [n_tiles] or null: set to flag_epoch when tile t of gin (and of lf_p / lf_q) is complete
  int flag_epoch;         // >= 1: flags only grow inside a sequence of chains, so a launch of a later chain that is already
  //                         resident cannot mistake an earlier chain's flag for its own
  int flag_out_epoch;     // what this launch stores: flag_epoch, or the NEXT evaluation's epoch when its first launch hangs on this one
  float* lf_p;            // leapfrog fold (null: off): p += lf_kick * gin ; lf_q += lf_drift * p  (lf_q may be x itself)
  float* lf_q;
  float lf_kick, lf_drift;
  RqsConsts c;
};

__device__ __forceinline__ int fb_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fb_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ bool fb_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// L2 prefetch of a contiguous span (the next tile's rows of x and dL/d(out)): the per-thread 8-byte loads that follow
// then miss L1 only
__device__ __forceinline__ void fb_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fb_red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

#ifdef FB_TRACE   // tools/ubench timing experiment: clock64 of the phases of CTA 0's second tile
__device__ long long fb_trace[256];
#define FB_T(on, slot) do { if ((on) && (threadIdx.x & 31) == 0) fb_trace[slot] = clock64(); } while (0)
#else
#define FB_T(on, slot) do { } while (0)
#endif

template <bool INV>
__global__ void __launch_bounds__(FB_THREADS, 1)
nsf_fused_bwd_kernel(const __grid_constant__ FusedBwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned char* sW1 = sm;
  unsigned char* sA1 = sW1 + FB_W1_BYTES;          // A1 -> h1 -> dZ1
  unsigned char* sA2 = sA1 + FB_A_BYTES;           // h2 -> dZ2
  unsigned char* sG = sA2 + FB_A_BYTES;            // dL/dparams of one pair of chunks (bf16, 3 K blocks)
  unsigned char* sRing = sG + FB_G_BYTES;
  float* sB1 = reinterpret_cast<float*>(sRing + FB_STAGES * FB_STAGE_BYTES);
  float* sB2 = sB1 + FB_HP;
  float* sB3 = sB2 + FB_HP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB3 + FB_NF * FB_PC);
  uint64_t* bar_w1 = bars;             // W1 resident                                        (1 + tx)
  uint64_t* bar_full = bars + 1;       // [3] ring piece landed                              (1 + tx)
  uint64_t* bar_empty = bars + 4;      // [3] ring piece consumed                            (1, commit)
  uint64_t* bar_a = bars + 7;          // operand written by the work warps, 5 phases/tile   (16 warps)
  uint64_t* bar_d12 = bars + 8;        // GEMM1 / GEMM2 / dH1 GEMM done, 3 phases/tile       (1, commit)
  uint64_t* bar_d3f = bars + 9;        // [2] GEMM3 chunk done                               (1, commit)
  uint64_t* bar_d3e = bars + 11;       // [2] chunk accumulator drained                      (16 warps)
  uint64_t* bar_gready = bars + 13;    // G of a pair of chunks written                      (32 = 16 warps x 2 chunks)
  uint64_t* bar_gfree = bars + 14;     // the dH2 GEMM has read G                            (1, commit)
  uint64_t* bar_dh2 = bars + 15;       // dH2 complete                                       (1, commit)
  uint64_t* bar_dx = bars + 16;        // dXc complete                                       (1, commit)
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned first = blockIdx.x, stride = gridDim.x;
  const unsigned n_tiles = (unsigned)a.n_tiles;
  const unsigned my_tiles = (n_tiles > first) ? (n_tiles - first + stride - 1) / stride : 0;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    mbar_init(bar_w1, 1);
    for (int i = 0; i < FB_STAGES; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_a, FB_WORK_WARPS);
    mbar_init(bar_d12, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_d3f[i], 1);
      mbar_init(&bar_d3e[i], FB_WORK_WARPS);
    }
    mbar_init(bar_gready, 2 * FB_WORK_WARPS);
    mbar_init(bar_gfree, 1);
    mbar_init(bar_dh2, 1);
    mbar_init(bar_dx, 1);
    fence_barrier_init();
  }
  // Programmatic dependent launch (as nsf_fused2.cu): this grid may be scheduled while the previous kernel of the stream
  // drains; nothing in global memory is touched before the wait, and the trigger lets the next launch do the same.
  // With tile flags the dependency on the previous launch is per TILE (rows are independent: tile t of this layer needs
  // only tile t of dL/d(out)), so this grid's CTAs start on the SMs the previous launch has already left -- at 65,536
  // rows (3.46 tiles per SM) that turns the four waves of every launch into 3.5.  Everything else this kernel reads
  // (x, the weight images) was complete before the first launch of the chain, which waits for the whole stream.
  if (a.flag_in == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  for (int i = tid; i < FB_HP; i += FB_THREADS) {
    sB1[i] = a.b1[i];
    sB2[i] = a.b2[i];
  }
  for (int i = tid; i < FB_NF * FB_PC; i += FB_THREADS) sB3[i] = a.b3[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == FB_WARP_TMA) {
    // ---------------- ring producer: pieces in the order the MMA warp consumes them
    uint32_t s = 0, ph = 0;
    auto put = [&](const unsigned char* src, uint32_t bytes) {
      mbar_wait_idle(&bar_empty[s], ph ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&bar_full[s], bytes);
        bulk_g2s(sRing + s * FB_STAGE_BYTES, src, bytes, &bar_full[s]);
      }
      __syncwarp();
      if (++s == FB_STAGES) {
        s = 0;
        ph ^= 1;
      }
    };
    if (my_tiles && lane == 0) {
      mbar_expect_tx(bar_w1, FB_W1_BYTES);
      bulk_g2s(sW1, a.w1_img, FB_W1_BYTES, bar_w1);
    }
    __syncwarp();
    for (unsigned it = 0; it < my_tiles; ++it) {
      if (it + 1 < my_tiles && lane == 0 && a.flag_in == nullptr) {
        const size_t nrow = (first + (size_t)(it + 1) * stride) * FB_ROWS;
        fb_prefetch_l2(a.x + nrow * 64, FB_ROWS * 256);
        fb_prefetch_l2(a.gout + nrow * 64, FB_ROWS * 256);
      }
      put(a.w2_img, FB_KB_BYTES);
      put(a.w2_img + FB_KB_BYTES, FB_KB_BYTES);
#pragma unroll 1
      for (int c = 0; c < FB_NCHUNK; ++c) {
        put(a.w3_img + (size_t)c * FB_W3C_BYTES, FB_W3C_BYTES);
        // the dH2 GEMM of pair p is issued after chunk 2p+2 (after the last chunk for the last pair)
        const int p = (c >= 2 && (c & 1) == 0) ? (c >> 1) - 1 : (c == FB_NCHUNK - 1 ? FB_NPAIR - 1 : -1);
        if (p >= 0)
          for (int kb = 0; kb < 3; ++kb) put(a.w3t_img + ((size_t)p * 3 + kb) * FB_KB_BYTES, FB_KB_BYTES);
      }
      put(a.w2t_img, FB_KB_BYTES);
      put(a.w2t_img + FB_KB_BYTES, FB_KB_BYTES);
      put(a.w1t_img, FB_W1T_BYTES);
      put(a.w1t_img + FB_W1T_BYTES, FB_W1T_BYTES);
    }
  } else if (warp == FB_WARP_MMA) {
    // ---------------- MMA issuer
    const uint32_t id_f16_128 = make_idesc_f16(FB_ROWS, FB_HP);
    const uint32_t id_f16_96 = make_idesc_f16(FB_ROWS, FB_NC);
    const uint32_t id_bf16_128 = make_idesc_bf16(FB_ROWS, FB_HP);
    const uint32_t id_bf16_32 = make_idesc_bf16(FB_ROWS, 32);
    const uint32_t aA1 = smem_u32(sA1), aA2 = smem_u32(sA2), aG = smem_u32(sG), aW1 = smem_u32(sW1), aRing = smem_u32(sRing);
    uint32_t s = 0, ph = 0, n_a = 0, g = 0, n_pair = 0;
    auto ring_next = [&]() {
      if (++s == FB_STAGES) {
        s = 0;
        ph ^= 1;
      }
    };
    // D (+)= A[K block kb] * ring piece, 4 K slices of 16; `first` clears the accumulator on the first slice
    auto mma_block = [&](uint32_t d, uint32_t a_base, uint32_t idesc, bool first) {
      const uint32_t bb = aRing + s * FB_STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(d, make_desc_sw128(a_base + k * 32), make_desc_sw128(bb + k * 32), idesc, (first && k == 0) ? 0u : 1u);
    };
    auto dh2_gemm = [&](int p) {          // dH2 (+)= G(pair p) * W3^T(pair p): 3 K blocks of 64 gradient columns
      mbar_wait_idle(bar_gready, n_pair & 1);
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < 3; ++kb) {
        mbar_wait_idle(&bar_full[s], ph);
        tc_fence_after();
        if (fb_elect_one()) {
          mma_block(tmem + FB_T_DH2, aG + kb * FB_KB_BYTES, id_bf16_128, p == 0 && kb == 0);
          umma_commit(&bar_empty[s]);
          if (kb == 2) {
            umma_commit(bar_gfree);
            if (p == FB_NPAIR - 1) umma_commit(bar_dh2);
          }
        }
        __syncwarp();
        ring_next();
      }
      ++n_pair;
    };
    if (my_tiles) mbar_wait_idle(bar_w1, 0);
    for (unsigned it = 0; it < my_tiles; ++it) {
      const bool tr = blockIdx.x == 0 && it == 1;
      (void)tr;
      // ---- GEMM1: D12 = A1 W1^T (32 real conditioning columns: K slices 0, 1)
      mbar_wait_idle(bar_a, n_a++ & 1);
      tc_fence_after();
      if (fb_elect_one()) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16(tmem + FB_T_D12, make_desc_sw128(aA1 + k * 32), make_desc_sw128(aW1 + k * 32), id_f16_128, k ? 1u : 0u);
        umma_commit(bar_d12);
      }
      __syncwarp();
      FB_T(tr, 0);
      // ---- GEMM2: D12 = h1 W2^T
      mbar_wait_idle(bar_a, n_a++ & 1);
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        mbar_wait_idle(&bar_full[s], ph);
        tc_fence_after();
        if (fb_elect_one()) {
          mma_block(tmem + FB_T_D12, aA1 + kb * FB_KB_BYTES, id_f16_128, kb == 0);
          umma_commit(&bar_empty[s]);
          if (kb == 1) umma_commit(bar_d12);
        }
        __syncwarp();
        ring_next();
      }
      FB_T(tr, 1);
      // ---- GEMM3 chunks interleaved with the dH2 GEMMs
      mbar_wait_idle(bar_a, n_a++ & 1);                       // h2 written
#pragma unroll 1
      for (int c = 0; c < FB_NCHUNK; ++c, ++g) {
        if (g >= 2) mbar_wait_idle(&bar_d3e[g & 1], ((g >> 1) + 1) & 1);
        mbar_wait_idle(&bar_full[s], ph);
        tc_fence_after();
        if (fb_elect_one()) {
          const uint32_t d = tmem + FB_T_D3 + (g & 1) * FB_NC;
          const uint32_t bb = aRing + s * FB_STAGE_BYTES;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d, make_desc_sw128(aA2 + kb * FB_KB_BYTES + k * 32), make_desc_sw128(bb + kb * (FB_NC * 128) + k * 32),
                        id_f16_96, (kb | k) ? 1u : 0u);
          umma_commit(&bar_empty[s]);
          umma_commit(&bar_d3f[g & 1]);
        }
        __syncwarp();
        ring_next();
        FB_T(tr, 2 + c);
        // pair p's dH2 GEMM goes after chunk 2p+2 has been issued (after the last chunk for the last pair)
        const int p = (c >= 2 && (c & 1) == 0) ? (c >> 1) - 1 : (c == FB_NCHUNK - 1 ? FB_NPAIR - 1 : -1);
        if (p >= 0) {
          dh2_gemm(p);
          FB_T(tr, 10 + p);
        }
      }
      // ---- dH1 = dZ2 W2 (B operand W2^T) -> D12
      mbar_wait_idle(bar_a, n_a++ & 1);                       // dZ2 written over h2
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        mbar_wait_idle(&bar_full[s], ph);
        tc_fence_after();
        if (fb_elect_one()) {
          mma_block(tmem + FB_T_D12, aA2 + kb * FB_KB_BYTES, id_bf16_128, kb == 0);
          umma_commit(&bar_empty[s]);
          if (kb == 1) umma_commit(bar_d12);
        }
        __syncwarp();
        ring_next();
      }
      FB_T(tr, 14);
      // ---- dXc = dZ1 W1 (B operand W1^T, 32 output columns)
      mbar_wait_idle(bar_a, n_a++ & 1);                       // dZ1 written over h1
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        mbar_wait_idle(&bar_full[s], ph);
        tc_fence_after();
        if (fb_elect_one()) {
          mma_block(tmem + FB_T_DX, aA1 + kb * FB_KB_BYTES, id_bf16_32, kb == 0);
          umma_commit(&bar_empty[s]);
          if (kb == 1) umma_commit(bar_dx);
        }
        __syncwarp();
        ring_next();
      }
      FB_T(tr, 15);
    }
  } else {
    // =============================== work warps ===============================
    const int q = warp & 3;
    const int slice = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t g = 0, n_pair = 0, n_d12 = 0;
    auto signal = [&]() {                      // this warp's part of an MMA operand is in shared memory
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a);
    };
    // columns [32 slice, 32 slice + 32) of this row sit in K block slice / 2, 16-byte chunks 4 (slice % 2) + 0..3
    const uint32_t blk_off = (uint32_t)(slice >> 1) * FB_KB_BYTES + (uint32_t)row * 128;
    const int ch0 = (slice & 1) * 4;
    // h = tanh(D12 + bias) -> fp16 operand
    auto tanh_epilogue = [&](const float* bias, unsigned char* dstA) {
      uint32_t v[32];
      tmem_ld32(tmem + FB_T_D12 + lane_sel + slice * 32, v);
      tmem_ld_wait();
      unsigned char* dst = dstA + blk_off;
      const float* bs = bias + slice * 32;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = tanh_approx(__uint_as_float(v[t * 8 + j]) + bs[t * 8 + j]);
        uint4 u;
        u.x = pack_f16x2(f[0], f[1]);
        u.y = pack_f16x2(f[2], f[3]);
        u.z = pack_f16x2(f[4], f[5]);
        u.w = pack_f16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(dst + (((ch0 + t) ^ (row & 7)) << 4)) = u;
      }
    };
    // dZ = D[tcol] * (1 - h^2), h read from the fp16 operand `buf`, dZ written over it as bf16
    auto tanh_backward = [&](uint32_t tcol, unsigned char* buf) {
      uint32_t v[32];
      tmem_ld32(tmem + tcol + lane_sel + slice * 32, v);
      tmem_ld_wait();
      unsigned char* dst = buf + blk_off;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint4* p = reinterpret_cast<uint4*>(dst + (((ch0 + t) ^ (row & 7)) << 4));
        const uint4 hv = *p;
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
          const float d0 = __uint_as_float(v[t * 8 + 2 * e]) * fmaf(-h.x, h.x, 1.f);
          const float d1 = __uint_as_float(v[t * 8 + 2 * e + 1]) * fmaf(-h.y, h.y, 1.f);
          o[e] = pack_bf16x2(d0, d1);
        }
        *p = make_uint4(o[0], o[1], o[2], o[3]);
      }
    };
    // x / dL/d(out) pairs are fetched one chunk ahead (across tiles too), so their latency hides behind an adjoint
    // In a tile-flag chain x itself may be produced inside the chain (the forward launches of the same evaluation), so a
    // tile's x is only touched once its flag has been seen: the cross-tile prefetches below are taken when a probe finds
    // the next tile's flag already set (the steady state) and otherwise made up for after the blocking wait.  All loads
    // of chain data go to L2 (ld.global.cg): an L1 line could predate the producing launch.
    float2 xn = make_float2(0.f, 0.f), gn = xn;
    // the 8 pairs whose conditioning columns this thread turns into the A1 operand, fetched before the previous tile's
    // tail (the waits on its last three GEMMs hide the latency)
    float4 a1n[4];
    bool have_next = false;                    // xn / a1n of the coming tile are already in registers
    if (my_tiles && a.flag_in == nullptr) {
      xn = __ldcg(reinterpret_cast<const float2*>(a.x + ((size_t)first * FB_ROWS + row) * 64) + slice);
      const float4* x4 = reinterpret_cast<const float4*>(a.x + ((size_t)first * FB_ROWS + row) * 64 + slice * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) a1n[j] = __ldcg(x4 + j);
      have_next = true;
    }
    for (unsigned it = 0; it < my_tiles; ++it) {
      const size_t tile = first + (size_t)it * stride;
      const size_t grow = tile * FB_ROWS + row;
      const float2* xr = reinterpret_cast<const float2*>(a.x + grow * 64);
      const float2* gor = reinterpret_cast<const float2*>(a.gout + grow * 64);
      float2* gir = reinterpret_cast<float2*>(a.gin + grow * 64);
      const bool tr = blockIdx.x == 0 && it == 1 && q == 0;
      (void)tr;
      FB_T(tr && slice == 0, 32);
      // ---- dL/d(out) of this tile is complete (per-tile dependency on the producing launch), first pair of it
      if (a.flag_in != nullptr) {
        if (lane == 0)
          while (fb_ld_acquire(a.flag_in + tile) < a.flag_epoch) __nanosleep(64);
        __syncwarp();
      }
      gn = __ldcg(gor + slice);
      if (!have_next) {
        xn = __ldcg(xr + slice);
        const float4* x4 = reinterpret_cast<const float4*>(xr + slice * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) a1n[j] = __ldcg(x4 + j);
      }
      have_next = false;
      // ---- A1: conditioning columns of features 8 slice .. 8 slice + 7 of this row (fp16, K block 0, chunk `slice`)
      {
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 pr = a1n[j];
          f[2 * j] = a.cond_first ? pr.x : pr.y;
          f[2 * j + 1] = a.cond_first ? pr.z : pr.w;
        }
        uint4 u;
        u.x = pack_f16x2_sat(f[0], f[1]);
        u.y = pack_f16x2_sat(f[2], f[3]);
        u.z = pack_f16x2_sat(f[4], f[5]);
        u.w = pack_f16x2_sat(f[6], f[7]);
        *reinterpret_cast<uint4*>(sA1 + row * 128 + ((slice ^ (row & 7)) << 4)) = u;
      }
      signal();
      FB_T(tr && slice == 0, 33);
      const float gl = a.gld ? __ldg(a.gld + grow) : a.gld_const;
      mbar_wait(bar_d12, n_d12++ & 1);
      tc_fence_after();
      FB_T(tr && slice == 0, 34);
      tanh_epilogue(sB1, sA1);                                 // h1 (kept for the tanh backward)
      signal();
      FB_T(tr && slice == 0, 35);
      mbar_wait(bar_d12, n_d12++ & 1);
      tc_fence_after();
      FB_T(tr && slice == 0, 36);
      tanh_epilogue(sB2, sA2);                                 // h2
      signal();
      FB_T(tr && slice == 0, 37);
      // ---- the 8 chunks: adjoint of feature 4 c + slice
#pragma unroll 1
      for (int c = 0; c < FB_NCHUNK; ++c, ++g) {
        const int f = c * FB_CF + slice;
        const float2 xc = xn, gc = gn;
        if (c + 1 < FB_NCHUNK) {
          xn = __ldcg(xr + f + FB_CF);
          gn = __ldcg(gor + f + FB_CF);
        } else if (it + 1 < my_tiles) {
          // next tile: prefetch only if its flag is already up (non-blocking probe by lane 0)
          int up = 1;
          if (a.flag_in != nullptr) {
            up = (lane == 0) ? (fb_ld_acquire(a.flag_in + tile + stride) >= a.flag_epoch ? 1 : 0) : 0;
            __syncwarp();
            up = __shfl_sync(0xffffffffu, up, 0);
          }
          have_next = up != 0;
          if (have_next) xn = __ldcg(reinterpret_cast<const float2*>(a.x + ((tile + stride) * FB_ROWS + row) * 64) + slice);
        }
        FB_T(tr, 64 + slice * 32 + c * 3);
        mbar_wait(&bar_d3f[g & 1], (g >> 1) & 1);
        tc_fence_after();
        FB_T(tr, 64 + slice * 32 + c * 3 + 1);
        uint32_t v[24];
        const uint32_t ta = tmem + FB_T_D3 + (g & 1) * FB_NC + lane_sel + slice * FB_PC;
        tmem_ld16(ta, v);
        tmem_ld8(ta + 16, v + 16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_d3e[g & 1]);
        float gxv, gD0, gD1;
        int bin;
        F2 gwh[8];
        rqs_element_bwd8_packed<INV>(PairRegParams{v, sB3 + f * FB_PC}, a.cond_first ? xc.y : xc.x, a.gout_scale * gc.y, gl,
                                     a.c, gxv, gwh, bin, gD0, gD1);
        // input-order pair (the conditioning column's pass-through gradient now, its path through the conditioner is
        // added at the end of the tile)
        const float gcond = a.gout_scale * gc.x;
        gir[f] = a.cond_first ? make_float2(gcond, gxv) : make_float2(gxv, gcond);
        // G of this pair of chunks: feature slot fs = (c & 1) * 4 + slice, 24 columns each, bf16, swizzled K-major.
        // The first chunk of a pair must not overwrite G before the previous pair's dH2 GEMM has read it.
        if ((c & 1) == 0) {
          if (n_pair >= 1) mbar_wait(bar_gfree, (n_pair - 1) & 1);
          ++n_pair;
        }
        const int fs = (c & 1) * FB_CF + slice;
        auto chunk_ptr = [&](int j) {
          const int ch = 3 * fs + j;
          return sG + (ch >> 3) * FB_KB_BYTES + row * 128 + (((ch & 7) ^ (row & 7)) << 4);
        };
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float l0, h0, l1, h1, l2, h2, l3, h3;
          unpk2(gwh[4 * j + 0], l0, h0);
          unpk2(gwh[4 * j + 1], l1, h1);
          unpk2(gwh[4 * j + 2], l2, h2);
          unpk2(gwh[4 * j + 3], l3, h3);
          uint4 u;
          u.x = pack_bf16x2(l0, h0);
          u.y = pack_bf16x2(l1, h1);
          u.z = pack_bf16x2(l2, h2);
          u.w = pack_bf16x2(l3, h3);
          *reinterpret_cast<uint4*>(chunk_ptr(j)) = u;
        }
        {
          // derivative logits: only slots bin - 1 and bin are non-zero -- clear the 16-byte chunk, then drop the two
          // bf16 values at their run-time positions (same thread, same addresses: program order holds)
          unsigned char* dch = chunk_ptr(2);
          *reinterpret_cast<uint4*>(dch) = make_uint4(0u, 0u, 0u, 0u);
          const uint32_t pr = pack_bf16x2(gD0, gD1);
          if (bin > 0) *reinterpret_cast<unsigned short*>(dch + 2 * (bin - 1)) = (unsigned short)(pr & 0xffffu);
          if (bin < 7) *reinterpret_cast<unsigned short*>(dch + 2 * bin) = (unsigned short)(pr >> 16);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_gready);
        FB_T(tr, 64 + slice * 32 + c * 3 + 2);
      }
      if (have_next) {
        const float4* x4 = reinterpret_cast<const float4*>(a.x + ((tile + stride) * FB_ROWS + row) * 64 + slice * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) a1n[j] = __ldcg(x4 + j);
      }
      // ---- dZ2 = dH2 (1 - h2^2), in place over h2 (every chunk GEMM has completed: bar_dh2 follows them)
      mbar_wait(bar_dh2, it & 1);
      tc_fence_after();
      FB_T(tr && slice == 0, 38);
      tanh_backward(FB_T_DH2, sA2);
      signal();
      FB_T(tr && slice == 0, 39);
      // ---- dZ1 = dH1 (1 - h1^2), in place over h1
      mbar_wait(bar_d12, n_d12++ & 1);
      tc_fence_after();
      FB_T(tr && slice == 0, 40);
      tanh_backward(FB_T_D12, sA1);
      signal();
      FB_T(tr && slice == 0, 41);
      // ---- the conditioning columns' path through the conditioner: features 8 slice .. 8 slice + 7 of this row, added
      // to the pass-through gradients the adjoint threads stored (the mbarrier round trips order those stores first)
      mbar_wait(bar_dx, it & 1);
      tc_fence_after();
      FB_T(tr && slice == 0, 42);
      {
        // dXc arrives row-per-lane; a 32 x 32 transpose per TMEM lane quadrant through the (now idle) G buffer turns the
        // adds into feature-per-lane accesses: 8 sectors per instruction instead of 32
        uint32_t v[8];
        tmem_ld8(tmem + FB_T_DX + lane_sel + slice * 8, v);
        tmem_ld_wait();
        tc_fence_before();
        float* sT = reinterpret_cast<float*>(sG) + q * 1024;            // [32 rows][32 features], column ^ row swizzle
#pragma unroll
        for (int j = 0; j < 8; ++j) sT[lane * 32 + ((slice * 8 + j) ^ lane)] = __uint_as_float(v[j]);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
        if (a.lf_p == nullptr) {
          float* gq = a.gin + (tile * FB_ROWS + q * 32) * 64 + 2 * lane + (a.cond_first ? 0 : 1);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = slice * 8 + j;
            fb_red_add(gq + (size_t)r * 64, sT[r * 32 + (lane ^ r)]);
          }
        } else {
          // Leapfrog folded into the launch that completes the force (the layer nearest the data): the thread that owns
          // (row, feature pair) finishes dL/dx of the pair -- 256 contiguous bytes per row and warp -- and advances the
          // momentum and the position of the same pair (velocity Verlet, applications/src/systems.py:331-336).  The
          // position may be this launch's own input: every read of this tile's x happened before this point.
          const size_t pair0 = ((tile * FB_ROWS + q * 32) * 64) / 2 + lane;
          float2* g2 = reinterpret_cast<float2*>(a.gin) + pair0;
          float2* p2 = reinterpret_cast<float2*>(a.lf_p) + pair0;
          float2* q2 = reinterpret_cast<float2*>(a.lf_q) + pair0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = slice * 8 + j;
            float2 gv = __ldcg(g2 + (size_t)r * 32);
            const float dxc = sT[r * 32 + (lane ^ r)];
            if (a.cond_first) gv.x += dxc; else gv.y += dxc;
            g2[(size_t)r * 32] = gv;
            float2 pv = p2[(size_t)r * 32];
            pv.x = fmaf(a.lf_kick, gv.x, pv.x);
            pv.y = fmaf(a.lf_kick, gv.y, pv.y);
            p2[(size_t)r * 32] = pv;
            if (a.lf_drift != 0.f) {
              float2 qv = q2[(size_t)r * 32];
              qv.x = fmaf(a.lf_drift, pv.x, qv.x);
              qv.y = fmaf(a.lf_drift, pv.y, qv.y);
              q2[(size_t)r * 32] = qv;
            }
          }
        }
      }
      FB_T(tr && slice == 0, 43);
      if (a.flag_out != nullptr) {
        // every work thread's stores and reductions of this tile are ordered before the flag (fence, CTA barrier of the
        // work warps, release store)
        if (a.lf_p != nullptr) asm volatile("fence.proxy.async;" ::: "memory");   // the position may be read by the next launch's TMA
        __threadfence();
        asm volatile("bar.sync 5, 512;" ::: "memory");
        if (tid == 0) fb_st_release(a.flag_out + tile, a.flag_out_epoch);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

constexpr size_t FB_SMEM = FB_W1_BYTES + 2 * FB_A_BYTES + FB_G_BYTES + FB_STAGES * FB_STAGE_BYTES +
                           (2 * FB_HP + FB_NF * FB_PC) * 4 + 32 * 8 + 1024;
static_assert(FB_SMEM <= 227 * 1024, "fused layer backward exceeds the 227 KB shared-memory limit");

RqsConsts make_rqs_consts(int K, float B);   // rqs_coupling.cu

}  // namespace nfk

using namespace nfk;

#ifdef FB_TRACE
extern "C" int nfk_fused_bwd_trace_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, fb_trace, sizeof(long long) * 256);
}
#endif

extern "C" int nfk_nsf_pairs_fused_bwd_leapfrog(const float* x, const float* grad_out, float grad_out_scale,
                                                const float* grad_logdet, float grad_logdet_const, float* grad_x,
                                                const void* w1_img, const void* w2_img, const void* w3_img,
                                                const void* w3t_img, const void* w2t_img, const void* w1t_img,
                                                const float* b1, const float* b2, const float* b3, int64_t N, int mask_col,
                                                float B, int inverse, int32_t* tile_flags_in, int32_t* tile_flags_out,
                                                int tile_flag_epoch, int tile_flag_out_epoch, float* momentum, float* position,
                                                float kick, float drift, void* stream);

extern "C" int nfk_nsf_pairs_fused_bwd(const float* x, const float* grad_out, float grad_out_scale, const float* grad_logdet,
                                       float grad_logdet_const, float* grad_x, const void* w1_img, const void* w2_img,
                                       const void* w3_img, const void* w3t_img, const void* w2t_img, const void* w1t_img,
                                       const float* b1, const float* b2, const float* b3, int64_t N, int mask_col, float B,
                                       int inverse, int32_t* tile_flags_in, int32_t* tile_flags_out, int tile_flag_epoch,
                                       void* stream) {
  return nfk_nsf_pairs_fused_bwd_leapfrog(x, grad_out, grad_out_scale, grad_logdet, grad_logdet_const, grad_x, w1_img, w2_img,
                                          w3_img, w3t_img, w2t_img, w1t_img, b1, b2, b3, N, mask_col, B, inverse,
                                          tile_flags_in, tile_flags_out, tile_flag_epoch, tile_flag_epoch, nullptr, nullptr, 0.f,
                                          0.f, stream);
}

extern "C" int nfk_nsf_pairs_fused_bwd_leapfrog(const float* x, const float* grad_out, float grad_out_scale, const float* grad_logdet,
                                       float grad_logdet_const, float* grad_x, const void* w1_img, const void* w2_img, const void* w3_img,
                                       const void* w3t_img, const void* w2t_img, const void* w1t_img, const float* b1,
                                       const float* b2, const float* b3, int64_t N, int mask_col, float B, int inverse,
                                       int32_t* tile_flags_in, int32_t* tile_flags_out, int tile_flag_epoch,
                                       int tile_flag_out_epoch, float* momentum, float* position, float kick, float drift,
                                       void* stream) {
  NFK_REQUIRE((tile_flags_in == nullptr && tile_flags_out == nullptr) || tile_flag_epoch >= 1,
              "nsf_pairs_fused_bwd: the flag epoch starts at 1");
  NFK_REQUIRE((momentum == nullptr) == (position == nullptr), "nsf_pairs_fused_bwd: momentum and position go together");
  NFK_REQUIRE(N >= 0 && N % FB_ROWS == 0, "nsf_pairs_fused_bwd: N must be a multiple of %d (got %lld)", FB_ROWS, (long long)N);
  NFK_REQUIRE(mask_col == 0 || mask_col == 1, "nsf_pairs_fused_bwd: mask column must be 0 or 1");
  NFK_REQUIRE(B > 0.f, "nsf_pairs_fused_bwd: tail bound must be positive");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && grad_out && grad_x && w1_img && w2_img && w3_img && w3t_img && w2t_img && w1t_img && b1 && b2 && b3,
              "nsf_pairs_fused_bwd: null device pointer");
  NFK_REQUIRE(grad_x != grad_out && grad_x != x, "nsf_pairs_fused_bwd: grad_x must not alias its inputs");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_x) |
                reinterpret_cast<uintptr_t>(w1_img) | reinterpret_cast<uintptr_t>(w2_img) | reinterpret_cast<uintptr_t>(w3_img) |
                reinterpret_cast<uintptr_t>(w3t_img) | reinterpret_cast<uintptr_t>(w2t_img) | reinterpret_cast<uintptr_t>(w1t_img)) & 15) == 0,
              "nsf_pairs_fused_bwd: pointers must be 16-byte aligned");
  FusedBwdArgs a{};
  a.x = x;
  a.gout = grad_out;
  a.gld = grad_logdet;
  a.gld_const = grad_logdet_const;
  a.gout_scale = grad_out_scale;
  a.gin = grad_x;
  a.w1_img = reinterpret_cast<const unsigned char*>(w1_img);
  a.w2_img = reinterpret_cast<const unsigned char*>(w2_img);
  a.w3_img = reinterpret_cast<const unsigned char*>(w3_img);
  a.w3t_img = reinterpret_cast<const unsigned char*>(w3t_img);
  a.w2t_img = reinterpret_cast<const unsigned char*>(w2t_img);
  a.w1t_img = reinterpret_cast<const unsigned char*>(w1t_img);
  a.b1 = b1;
  a.b2 = b2;
  a.b3 = b3;
  a.n_tiles = N / FB_ROWS;
  a.cond_first = (mask_col == 0);
  a.flag_in = tile_flags_in;
  a.flag_out = tile_flags_out;
  a.flag_epoch = tile_flag_epoch;
  a.flag_out_epoch = tile_flag_out_epoch >= 1 ? tile_flag_out_epoch : tile_flag_epoch;
  a.lf_p = momentum;
  a.lf_q = position;
  a.lf_kick = kick;
  a.lf_drift = drift;
  a.c = make_rqs_consts(8, B);
  auto kern = inverse ? nsf_fused_bwd_kernel<true> : nsf_fused_bwd_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FB_SMEM);
  if (e != cudaSuccess) {
    set_error("nsf_pairs_fused_bwd: cannot set %zu B dynamic shared memory: %s", FB_SMEM, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  const long long cap = sm_count();
  const long long grid = a.n_tiles < cap ? a.n_tiles : cap;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(FB_THREADS);
  cfg.dynamicSmemBytes = FB_SMEM;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see griddepcontrol.wait in the kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) {
    set_error("nsf_pairs_fused_bwd: launch failed: %s", cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  count_launch();
  return check_launch("nsf_pairs_fused_bwd");
}
