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
constexpr int LB_THREADS = 128;

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

template <bool OUT_F32>
__global__ void __launch_bounds__(LB_THREADS)
linear_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long ldx,
                   const __nv_bfloat16* __restrict__ W, long long ldw,
                   const float* __restrict__ bias, void* __restrict__ Y, long long ldy,
                   long long M, int K, int N, int BN, int act, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t mma_done[2];
  __shared__ uint32_t tmem_base_s;
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
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(LB_M, bn);

  const int KB = (K + LB_K - 1) / LB_K;
  uint32_t ph[2] = {0u, 0u};
  for (int kb = 0; kb < KB; ++kb) {
    const int s = kb & 1;
    if (kb >= 2) {                       // the MMAs that read this stage two blocks ago are done
      mbar_wait(&mma_done[s], ph[s]);
      ph[s] ^= 1u;
    }
    unsigned char* sa = sbase + s * stage_bytes;
    unsigned char* sb = sa + a_bytes;
    load_tile_sw128(sa, X, ldx, m0, M, kb * LB_K, K, LB_M, tid);
    load_tile_sw128(sb, W, ldw, n0, N, kb * LB_K, K, bn, tid);
    fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
#pragma unroll
      for (int k = 0; k < LB_K / 16; ++k) {
        const uint64_t da = make_desc_sw128(a_addr + k * 32);
        const uint64_t db = make_desc_sw128(b_addr + k * 32);
        umma_bf16(tmem_d, da, db, idesc, (kb | k) ? 1u : 0u);
      }
      umma_commit(&mma_done[s]);         // implies tcgen05.fence::before_thread_sync
    }
  }
  // wait for the last commit of each stage that is still outstanding
  {
    const int last = (KB - 1) & 1;
    if (KB >= 2) {
      mbar_wait(&mma_done[last ^ 1], ph[last ^ 1]);
    }
    mbar_wait(&mma_done[last], ph[last]);
  }
  tc_fence_after();

  // ---- epilogue: thread = row (TMEM lane), 32 columns per tcgen05.ld
  const long long row = m0 + warp * 32 + lane;
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < bn; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(lane_addr + (uint32_t)c0, v);
    tmem_ld_wait();
    if (row < M) {
      const int ncol = min(min(32, bn - c0), N - (n0 + c0));   // columns this tile owns in the chunk
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float t = __uint_as_float(v[j]);
        if (j < ncol) {
          if (bias) t += __ldg(bias + n0 + c0 + j);
          if (act == 1) t = tanh_approx(t);
        } else {
          t = 0.f;
        }
        f[j] = t;
      }
      if (OUT_F32) {
        float* yp = reinterpret_cast<float*>(Y) + row * ldy + n0 + c0;
        if (ncol == 32 && ((reinterpret_cast<uintptr_t>(yp) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(yp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
          for (int j = 0; j < ncol; ++j) yp[j] = f[j];
        }
      } else {
        __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(Y) + row * ldy + n0 + c0;
        // pad columns up to ldy are written as zeros so the next layer can read K = ldy
        // the last N tile also zero-fills the pad columns [N, ldy) so the next layer reads K = ldy
        const long long left = ldy - (n0 + c0);
        const int nwrite = (n0 + bn >= N && c0 + 32 >= bn) ? (left < 32 ? (int)left : 32) : ncol;
        if (nwrite == 32 && ((reinterpret_cast<uintptr_t>(yp) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0);
            u.y = *reinterpret_cast<uint32_t*>(&p1);
            u.z = *reinterpret_cast<uint32_t*>(&p2);
            u.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(yp + j) = u;
          }
        } else {
          for (int j = 0; j < nwrite; ++j) yp[j] = __float2bfloat16_rn(f[j]);
        }
      }
    }
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
  const size_t smem = 2 * ((size_t)LB_M * 128 + (size_t)BN * 128) + 1024;
  const long long gm = (M + LB_M - 1) / LB_M;
  NFK_REQUIRE(gm < (1LL << 31), "linear_bf16: too many rows");
  dim3 grid((unsigned)gm, (unsigned)n_tiles);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* Xb = reinterpret_cast<const __nv_bfloat16*>(X);
  const __nv_bfloat16* Wb = reinterpret_cast<const __nv_bfloat16*>(W);
  cudaError_t e;
  if (out_f32) {
    e = cudaFuncSetAttribute(linear_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      linear_bf16_kernel<true><<<grid, LB_THREADS, smem, st>>>(Xb, ldx, Wb, ldw, b, Y, ldy, M, K, Nout, BN, act, cols);
  } else {
    e = cudaFuncSetAttribute(linear_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      linear_bf16_kernel<false><<<grid, LB_THREADS, smem, st>>>(Xb, ldx, Wb, ldw, b, Y, ldy, M, K, Nout, BN, act, cols);
  }
  if (e != cudaSuccess) {
    set_error("linear_bf16: cannot set %zu B dynamic shared memory: %s", smem, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  count_launch();
  return check_launch("linear_bf16");
}

}  // extern "C"
