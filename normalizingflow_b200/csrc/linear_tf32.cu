// fp32-accurate conditioner GEMM on the tensor cores: Y = act(X W^T + b) with fp32 inputs and outputs,
// one nn.Linear (+Tanh) of FCNN (reference nf/flows.py:26-35) for the parity mode (z / log_det within
// 1e-5 of the reference), replacing the CUDA-core fp32 kernel of linear_f32.cu wherever the shape allows.
//
// 3xTF32: every fp32 operand is split as x = hi + lo with hi = x truncated to TF32 (top 19 bits) and
// lo = x - hi (exact in fp32; its own truncation to TF32 loses at most 2^-22 |x|), and the product is
// accumulated in fp32 in TMEM as  hi*hi + hi*lo + lo*hi  — three tcgen05.mma.kind::tf32 per K step.
// The dropped lo*lo term is 2^-22 relative, so the result carries fp32-class error (measured
// <= 4e-7 of the output scale against an fp64 product) at 6x the cost of a bf16 MMA, which is still
// several times the fp32 CUDA-core rate.
//
// The tensor core's fp32 accumulation truncates when it aligns addends (measured: ~2e-7 of the output scale per
// MMA, biased), so the error of one long accumulation chain grows linearly with K.  The K blocks are therefore
// (k-steps of 8) are spread round-robin over LT_REGIONS separate TMEM accumulators and the partial sums are added with
// round-to-nearest fp32 adds in the epilogue: a chain is K / (8 * LT_REGIONS) k-steps long.
//
// One CTA computes a [128 x BN] tile (BN <= 64).  Per 32-wide K block (32 fp32 = 128 bytes = one
// SWIZZLE_128B row) every thread loads 16-byte pieces of X and W into registers, splits them, and writes
// the hi and lo images (K-major SWIZZLE_128B, same layout as the bf16 kernels: the swizzle is on bytes);
// the next block's global loads are in flight while the elected thread issues the 12 MMAs of this one;
// two shared-memory stages, completion by tcgen05.commit on an mbarrier; epilogue as linear_bf16.cu
// (tcgen05.ld, bias, accurate tanhf, warp-transposed coalesced fp32 row segments).
#include "tc05.cuh"

namespace nfk {

constexpr int LT_M = 128;
constexpr int LT_KF = 32;          // fp32 per K block
constexpr int LT_THREADS = 256;
// Two tile configurations (template parameters LT_BN = widest N tile, LT_REGIONS = TMEM accumulators used
// round-robin over the k-steps, LT_STAGES = shared-memory operand stages), both two CTAs per SM (one CTA's
// epilogue runs under the other's main loop):
//   K <= 256: <128, 2, 1>  64 KB of operands, 2 x 128 TMEM columns; X is split half as often as with 64-wide
//             tiles (2^20 x 128 -> 736: 2.16 ms against 3.07 ms); a chain is at most 16 k-steps
//   K >  256: < 64, 4, 2>  96 KB, 4 x 64 TMEM columns: four accumulators keep the chains short where K is long
//             (K = 800: 6.1e-6 of the output scale against 1.2e-5 with two accumulators)

// kind::tf32 instruction descriptor: D fp32 (bits 4-5 = 1), A/B TF32 (format 2 in bits 7-9, 10-12),
// both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Piece {
  uint4 v;
};

template <int LT_BN, int LT_REGIONS, int LT_STAGES>
__global__ void __launch_bounds__(LT_THREADS)
linear_tf32x3_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ W, long long ldw,
                     const float* __restrict__ bias, float* __restrict__ Y, long long ldy, long long M, int K,
                     int N, int BN, int n_tiles, int act, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t mma_done[LT_STAGES];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[LT_BN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // consecutive CTAs share the X tile (N tile index fastest): its re-reads hit L2
  const long long m0 = (long long)(blockIdx.x / n_tiles) * LT_M;
  const int n0 = (int)(blockIdx.x % n_tiles) * BN;
  const int bn = min(BN, ((N - n0) + 15) & ~15);       // this tile's MMA N (multiple of 16)
  const uint32_t a_bytes = LT_M * 128, b_bytes = (uint32_t)BN * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;        // A_hi | A_lo | B_hi | B_lo
  unsigned char* sbase = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);

  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < LT_STAGES; ++s) mbar_init(&mma_done[s], 1);
    fence_barrier_init();
  }
  for (int i = tid; i < bn; i += LT_THREADS) sbias[i] = (bias && n0 + i < N) ? bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc_tf32(LT_M, bn);
  const int region_cols = (int)tmem_cols / LT_REGIONS;

  const int KB = (K + LT_KF - 1) / LT_KF;
  constexpr int A_PER = LT_M * 8 / LT_THREADS;          // 16-byte pieces of the A tile per thread (4)
  constexpr int B_PER = LT_BN * 8 / LT_THREADS;         // ... of the widest B tile (4)
  Piece pa[A_PER], pb[B_PER];
  // piece j of this thread is always the same (row, 16-byte column) of the tile: pointers hoisted out of the K loop
  const int pc = (tid & 7) * 4;                          // first fp32 column of this thread's pieces within a K block
  const float* ap[A_PER];
  const float* bp[B_PER];
#pragma unroll
  for (int j = 0; j < A_PER; ++j) {
    const long long gr = m0 + ((tid + j * LT_THREADS) >> 3);
    ap[j] = gr < M ? X + gr * ldx + pc : nullptr;
  }
#pragma unroll
  for (int j = 0; j < B_PER; ++j) {
    const int r = (tid + j * LT_THREADS) >> 3;
    bp[j] = (r < bn && n0 + r < N) ? W + (long long)(n0 + r) * ldw + pc : nullptr;
  }
  auto fetch = [&](int kb) {
    const int k0 = kb * LT_KF;
    const bool in_k = k0 + pc < K;
#pragma unroll
    for (int j = 0; j < A_PER; ++j)
      pa[j].v = (ap[j] && in_k) ? __ldg(reinterpret_cast<const uint4*>(ap[j] + k0)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < B_PER; ++j)
      pb[j].v = (bp[j] && in_k) ? __ldg(reinterpret_cast<const uint4*>(bp[j] + k0)) : make_uint4(0u, 0u, 0u, 0u);
  };
  auto split_store = [&](unsigned char* hi_tile, unsigned char* lo_tile, int i, const uint4& v) {
    const int r = i >> 3, c = i & 7;
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    uint4 h, l;
    h.x = v.x & 0xFFFFE000u;
    h.y = v.y & 0xFFFFE000u;
    h.z = v.z & 0xFFFFE000u;
    h.w = v.w & 0xFFFFE000u;
    l.x = __float_as_uint(__fsub_rn(__uint_as_float(v.x), __uint_as_float(h.x)));
    l.y = __float_as_uint(__fsub_rn(__uint_as_float(v.y), __uint_as_float(h.y)));
    l.z = __float_as_uint(__fsub_rn(__uint_as_float(v.z), __uint_as_float(h.z)));
    l.w = __float_as_uint(__fsub_rn(__uint_as_float(v.w), __uint_as_float(h.w)));
    *reinterpret_cast<uint4*>(hi_tile + off) = h;
    *reinterpret_cast<uint4*>(lo_tile + off) = l;
  };

  fetch(0);
  for (int kb = 0; kb < KB; ++kb) {
    const int s = kb % LT_STAGES;
    unsigned char* st = sbase + s * stage_bytes;
    if (kb >= LT_STAGES) mbar_wait(&mma_done[s], ((kb - LT_STAGES) / LT_STAGES) & 1);   // block kb-2 has left this stage
#pragma unroll
    for (int j = 0; j < A_PER; ++j) split_store(st, st + a_bytes, tid + j * LT_THREADS, pa[j].v);
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      const int i = tid + j * LT_THREADS;
      if ((i >> 3) < BN) split_store(st + 2 * a_bytes, st + 2 * a_bytes + b_bytes, i, pb[j].v);
    }
    if (kb + 1 < KB) fetch(kb + 1);          // in flight while this block's MMAs are issued and run
    fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
#pragma unroll
      for (int k = 0; k < LT_KF / 8; ++k) {
        const uint64_t dah = make_desc_sw128(a_hi + k * 32), dal = make_desc_sw128(a_lo + k * 32);
        const uint64_t dbh = make_desc_sw128(b_hi + k * 32), dbl = make_desc_sw128(b_lo + k * 32);
        const int step = kb * (LT_KF / 8) + k;                                  // k-step index over the whole K
        const uint32_t d = tmem_d + (uint32_t)((step % LT_REGIONS) * region_cols);
        umma_tf32(d, dal, dbh, idesc, step >= LT_REGIONS ? 1u : 0u);            // small terms first
        umma_tf32(d, dah, dbl, idesc, 1u);
        umma_tf32(d, dah, dbh, idesc, 1u);
      }
      umma_commit(&mma_done[s]);
    }
  }
  mbar_wait(&mma_done[(KB - 1) % LT_STAGES], ((KB - 1) / LT_STAGES) & 1);
  tc_fence_after();
  __syncthreads();                           // the operand stages are free: reused as the transpose staging

  // ---- epilogue: thread = row (TMEM lane); a warp's 32 x 32 block is transposed through padded shared
  // memory so that every global store writes one contiguous 128-byte row segment
  float* stg = reinterpret_cast<float*>(sbase) + warp * (32 * 33);
  const int q = warp & 3;
  const long long row_w0 = m0 + q * 32;
  const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(Y) & 15) == 0) && (ldy % 4 == 0) && (n0 % 4 == 0);
  for (int c0 = (warp >> 2) * 32; c0 < bn; c0 += 64) {
    uint32_t v[32];
    float acc[32];
    const int used = min((K + 7) / 8, LT_REGIONS);          // accumulators that received at least one k-step
    tmem_ld32(lane_addr + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
    for (int rg = 1; rg < used; ++rg) {                       // partial sums of the other accumulators (fp32, RN)
      tmem_ld32(lane_addr + (uint32_t)(rg * region_cols + c0), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
    }
    const int ncol = min(min(32, bn - c0), N - (n0 + c0));
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float t = acc[j];
      if (j < ncol) {
        t += sbias[c0 + j];
        if (act == 1) t = tanhf(t);
      }
      stg[lane * 33 + j] = t;
    }
    __syncwarp();
    float* yb = Y + n0 + c0;
    if (vec_ok && ncol == 32) {
      // 8 lanes x 16 bytes cover one 128-byte row segment: four rows per store instruction
      const int rr = lane >> 3, cc = (lane & 7) * 4;
      float* yrow = yb + (row_w0 + rr) * ldy + cc;
#pragma unroll
      for (int r = 0; r < 32; r += 4) {
        if (row_w0 + r + rr < M) {
          const float* sp = stg + (r + rr) * 33 + cc;
          *reinterpret_cast<float4*>(yrow) = make_float4(sp[0], sp[1], sp[2], sp[3]);
        }
        yrow += 4 * ldy;
      }
    } else if (lane < ncol) {
      for (int r = 0; r < 32; ++r)
        if (row_w0 + r < M) yb[(row_w0 + r) * ldy + lane] = stg[r * 33 + lane];
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

}  // namespace nfk

using namespace nfk;

extern "C" {

int nfk_linear_tf32x3(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* b, float* Y,
                      int64_t ldy, int64_t M, int K, int Nout, int act, void* stream) {
  NFK_REQUIRE(M >= 0 && K > 0 && Nout > 0, "linear_tf32x3: bad shape M=%lld K=%d N=%d", (long long)M, K, Nout);
  NFK_REQUIRE(act == 0 || act == 1, "linear_tf32x3: act must be 0 (identity) or 1 (tanh)");
  NFK_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0,
              "linear_tf32x3: K and the row strides must be multiples of 4 floats (16 bytes)");
  NFK_REQUIRE(ldx >= K && ldw >= K && ldy >= Nout, "linear_tf32x3: row stride smaller than the row");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(X && W && Y, "linear_tf32x3: null device pointer");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15) == 0,
              "linear_tf32x3: X and W must be 16-byte aligned");
  const bool wide = K <= 256;
  const int bn_max = wide ? 128 : 64, regions = wide ? 2 : 4, stages = wide ? 1 : 2;
  const int n_tiles = (Nout + bn_max - 1) / bn_max;
  int BN = (((Nout + n_tiles - 1) / n_tiles) + 15) & ~15;
  if (BN < 16) BN = 16;
  uint32_t cols = 32;
  while ((int)cols < BN) cols <<= 1;
  cols *= regions;                           // `regions` accumulators of a power-of-two width each (<= 256 columns)
  const size_t smem = stages * (2 * (size_t)LT_M * 128 + 2 * (size_t)BN * 128) + 1024;
  const long long gm = (M + LT_M - 1) / LT_M;
  NFK_REQUIRE(gm * n_tiles < (1LL << 31), "linear_tf32x3: too many tiles");
  auto kern = wide ? linear_tf32x3_kernel<128, 2, 1> : linear_tf32x3_kernel<64, 4, 2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("linear_tf32x3: cannot set %zu B dynamic shared memory: %s", smem, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  kern<<<(unsigned)(gm * n_tiles), LT_THREADS, smem, (cudaStream_t)stream>>>(X, ldx, W, ldw, b, Y, ldy, M, K, Nout, BN,
                                                                             n_tiles, act, cols);
  count_launch();
  return check_launch("linear_tf32x3");
}

}  // extern "C"
