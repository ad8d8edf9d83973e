// Conditioner GEMM on the 5th-generation tensor cores: Y = act(X W^T + b), one nn.Linear (+Tanh)
// of FCNN (reference nf/flows.py:26-35) with bf16 operands and fp32 accumulation in TMEM.
//
// One CTA computes a [128 x BN] output tile (BN <= 256).  Per 64-wide K block the 128 A rows
// and BN W rows are written to shared memory in the canonical K-major SWIZZLE_128B layout
// (8-row x 128-byte atoms, 16-byte chunk index XOR row%8), one elected thread issues four
// tcgen05.mma (M=128, N=BN, K=16) per block, completion is tracked with tcgen05.commit on an
// mbarrier (two shared-memory stages), and the epilogue reads the accumulator with tcgen05.ld
// (one TMEM lane = one output row per thread), adds the bias, applies tanh (MUFU.TANH) and
// writes bf16 (next layer's operand) or fp32 (spline parameters) rows.
#include "tc05.cuh"

namespace nfk {

constexpr int LB_M = 128;      // rows per CTA = TMEM lanes
constexpr int LB_K = 64;       // K block: 64 bf16 = 128 bytes = one swizzle row
constexpr int LB_THREADS = 256;   // 8 warps: all load operands; epilogue = 4 lane quadrants x 2 column halves

// rows x 64 bf16 tile of a row-major matrix -> swizzled smem (zero fill outside the matrix)
__device__ __forceinline__ void load_tile_sw128(unsigned char* smem_tile, const __nv_bfloat16* g,
                                                long long ld, long long row0, long long n_rows,
                                                int k0, int K, int rows, int tid) {
  // 8 chunks of 16 bytes per row
  for (int i = tid; i < rows * 8; i += LB_THREADS) {
    const int r = i >> 3, c = i & 7;
    const long long gr = row0 + r;
    const int gk = k0 + c * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (gr < n_rows && gk < K) v = __ldg(reinterpret_cast<const uint4*>(g + gr * ld + gk));
    *reinterpret_cast<uint4*>(smem_tile + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
}

// same tile, but with cp.async (LDGSTS): 16-byte chunks go global -> shared without passing through
// registers, so several K blocks can be in flight per thread (src-size 0 zero-fills)
__device__ __forceinline__ void load_tile_sw128_async(unsigned char* smem_tile, const __nv_bfloat16* g,
                                                      long long ld, long long row0, long long n_rows,
                                                      int k0, int K, int rows, int tid) {
  for (int i = tid; i < rows * 8; i += LB_THREADS) {
    const int r = i >> 3, c = i & 7;
    const long long gr = row0 + r;
    const int gk = k0 + c * 8;
    const bool ok = gr < n_rows && gk < K;
    const __nv_bfloat16* src = ok ? g + gr * ld + gk : g;
    const uint32_t dst = smem_u32(smem_tile + r * 128 + ((c ^ (r & 7)) << 4));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <bool OUT_F32, int LB_STAGES>
__global__ void __launch_bounds__(LB_THREADS)
linear_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long ldx,
                   const __nv_bfloat16* __restrict__ W, long long ldw,
                   const float* __restrict__ bias, void* __restrict__ Y, long long ldy,
                   long long M, int K, int N, int BN, int act, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t mma_done[LB_STAGES];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * LB_M;
  const int n0 = blockIdx.y * BN;
  const int bn = min(BN, ((N - n0) + 15) & ~15);       // this tile's MMA N (multiple of 16)
  // stage s: A at s*stage_bytes, B right after
  const uint32_t a_bytes = LB_M * 128, b_bytes = (uint32_t)BN * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  unsigned char* sbase = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);

  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < LB_STAGES; ++s) mbar_init(&mma_done[s], 1);
    fence_barrier_init();
  }
  for (int i = tid; i < bn; i += LB_THREADS) sbias[i] = (bias && n0 + i < N) ? bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(LB_M, bn);

  // LB_STAGES-deep cp.async ring over the K blocks; the tensor core drains a stage (tracked by
  // tcgen05.commit on mma_done[stage]) before it is refilled
  const int KB = (K + LB_K - 1) / LB_K;
  auto issue_loads = [&](int kb) {
    unsigned char* sa = sbase + (kb % LB_STAGES) * stage_bytes;
    load_tile_sw128_async(sa, X, ldx, m0, M, kb * LB_K, K, LB_M, tid);
    load_tile_sw128_async(sa + a_bytes, W, ldw, n0, N, kb * LB_K, K, bn, tid);
  };
  for (int kb = 0; kb < LB_STAGES - 1; ++kb) {
    if (kb < KB) issue_loads(kb);
    cp_async_commit();
  }
  for (int kb = 0; kb < KB; ++kb) {
    const int s = kb % LB_STAGES;
    // refill the stage K block kb-1 used (its MMAs must have completed) with block kb+STAGES-1
    const int nxt = kb + LB_STAGES - 1;
    if (nxt < KB) {
      if (kb >= 1) mbar_wait(&mma_done[(kb - 1) % LB_STAGES], ((kb - 1) / LB_STAGES) & 1);
      issue_loads(nxt);
    }
    cp_async_commit();
    cp_async_wait<LB_STAGES - 1>();      // this thread's part of block kb has landed
    fence_proxy_async();                 // ... and is visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sbase + s * stage_bytes), b_addr = a_addr + a_bytes;
#pragma unroll
      for (int k = 0; k < LB_K / 16; ++k)
        umma_bf16(tmem_d, make_desc_sw128(a_addr + k * 32), make_desc_sw128(b_addr + k * 32), idesc,
                  (kb | k) ? 1u : 0u);
      umma_commit(&mma_done[s]);         // implies tcgen05.fence::before_thread_sync
    }
  }
  // MMAs complete in order: the last block's commit covers all of them
  mbar_wait(&mma_done[(KB - 1) % LB_STAGES], ((KB - 1) / LB_STAGES) & 1);
  tc_fence_after();
  __syncthreads();                       // every thread sees the operand stages as free (staging reuse)

  // ---- epilogue: thread = row (TMEM lane), 32 columns per tcgen05.ld.  The 32x32 block of a warp
  // is transposed through a padded shared-memory tile (the operand stages are idle now) so that
  // every global store instruction writes one contiguous row segment (128 B fp32 / 64 B bf16)
  // instead of 32 scattered 16-byte pieces.
  float* stg = reinterpret_cast<float*>(sbase) + warp * (32 * 33);
  const int q = warp & 3;                                  // TMEM lane quadrant of this warp
  const long long row_w0 = m0 + q * 32;                    // first row of this warp
  const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
  const bool last_tile = (n0 + bn >= N);
  for (int c0 = (warp >> 2) * 32; c0 < bn; c0 += 64) {
    uint32_t v[32];
    tmem_ld32(lane_addr + (uint32_t)c0, v);
    tmem_ld_wait();
    const int ncol = min(min(32, bn - c0), N - (n0 + c0));   // columns this tile owns in the chunk
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float t = __uint_as_float(v[j]);
      if (j < ncol) {
        t += sbias[c0 + j];
        if (act == 1) t = tanh_approx(t);
      } else {
        t = 0.f;
      }
      stg[lane * 33 + j] = t;
    }
    __syncwarp();
    if (OUT_F32) {
      float* yb = reinterpret_cast<float*>(Y) + n0 + c0;
      if (lane < ncol)
        for (int r = 0; r < 32; ++r)
          if (row_w0 + r < M) yb[(row_w0 + r) * ldy + lane] = stg[r * 33 + lane];
    } else {
      // the last N tile also zero-fills the pad columns [N, ldy) so the next layer reads K = ldy
      const long long left = ldy - (n0 + c0);
      const int nwrite = (last_tile && c0 + 32 >= bn) ? (left < 32 ? (int)left : 32) : ncol;
      __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(Y) + n0 + c0;
      // two rows per instruction: lanes 0..15 -> row r, lanes 16..31 -> row r+1, 2 columns each
      const int half = lane >> 4, cp = (lane & 15) * 2;
      for (int r = 0; r < 32; r += 2) {
        const long long gr = row_w0 + r + half;
        if (gr < M && cp < nwrite) {
          const float f0 = stg[(r + half) * 33 + cp], f1 = stg[(r + half) * 33 + cp + 1];
          __nv_bfloat16* dst = yb + gr * ldy + cp;
          if (cp + 1 < nwrite && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
            *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(f0, f1);
          } else {
            dst[0] = __float2bfloat16_rn(f0);
            if (cp + 1 < nwrite) dst[1] = __float2bfloat16_rn(f1);
          }
        }
      }
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

int nfk_linear_bf16(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* b,
                    void* Y, int64_t ldy, int64_t M, int K, int Nout, int act, int out_f32,
                    void* stream) {
  NFK_REQUIRE(M >= 0 && K > 0 && Nout > 0, "linear_bf16: bad shape M=%lld K=%d N=%d", (long long)M, K,
              Nout);
  NFK_REQUIRE(act == 0 || act == 1, "linear_bf16: act must be 0 (identity) or 1 (tanh)");
  NFK_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0,
              "linear_bf16: K and the row strides must be multiples of 8 bf16 (16 bytes); pad with zeros");
  NFK_REQUIRE(ldx >= K && ldw >= K && ldy >= Nout, "linear_bf16: row stride smaller than the row");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(X && W && Y, "linear_bf16: null device pointer");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15) == 0,
              "linear_bf16: X and W must be 16-byte aligned");
  // N tile: as wide as possible (<= 256) with the fewest tiles
  const int n_tiles = (Nout + 255) / 256;
  int BN = (((Nout + n_tiles - 1) / n_tiles) + 15) & ~15;
  if (BN < 16) BN = 16;
  uint32_t cols = 32;
  while ((int)cols < BN) cols <<= 1;
  const int KBh = (K + LB_K - 1) / LB_K;
  const int stages = KBh <= 2 ? 2 : 4;
  const size_t smem = stages * ((size_t)LB_M * 128 + (size_t)BN * 128) + 1024;
  const long long gm = (M + LB_M - 1) / LB_M;
  NFK_REQUIRE(gm < (1LL << 31), "linear_bf16: too many rows");
  dim3 grid((unsigned)gm, (unsigned)n_tiles);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* Xb = reinterpret_cast<const __nv_bfloat16*>(X);
  const __nv_bfloat16* Wb = reinterpret_cast<const __nv_bfloat16*>(W);
  cudaError_t e = cudaSuccess;
#define NFK_LB(F32, ST)                                                                                   \
  do {                                                                                                    \
    e = cudaFuncSetAttribute(linear_bf16_kernel<F32, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                             (int)smem);                                                                  \
    if (e == cudaSuccess)                                                                                 \
      linear_bf16_kernel<F32, ST><<<grid, LB_THREADS, smem, st>>>(Xb, ldx, Wb, ldw, b, Y, ldy, M, K, Nout, \
                                                                  BN, act, cols);                         \
  } while (0)
  if (out_f32) {
    if (stages == 2) NFK_LB(true, 2); else NFK_LB(true, 4);
  } else {
    if (stages == 2) NFK_LB(false, 2); else NFK_LB(false, 4);
  }
#undef NFK_LB
  if (e != cudaSuccess) {
    set_error("linear_bf16: cannot set %zu B dynamic shared memory: %s", smem, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  count_launch();
  return check_launch("linear_bf16");
}

}  // extern "C"
