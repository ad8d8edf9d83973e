// Weight-gradient GEMM of the conditioner (autograd through nn.Linear, reference nf/flows.py:26-35;
// training step applications/src/train.py:22-29):   C[p, q] += sum_n A[n, p] * B[n, q]
// contracting the BATCH dimension of two bf16 activation / gradient images
// ([m_tiles][KB][128 rows][64 cols], SWIZZLE_128B).  The images are read exactly as the forward and
// dgrad GEMMs wrote them: a 128-row x 64-column block is, for a contraction over rows, the canonical
// MN-major SWIZZLE_128B operand of tcgen05.mma (8-row x 128-byte atoms, K = row index), so the
// tensor core transposes on the fly -- no re-layout, no row-major copies.
//
// One CTA = one 128 x (64*nbq) output tile and one slice of the batch (split-K); per 64-row half
// block: 4 x tcgen05.mma (M=128, N=64*nbq, K=16), operands through a 4-stage TMA bulk ring;
// fp32 accumulator in TMEM; epilogue adds the partial tile into C with red.global.add.f32.
#include "tc05.cuh"

namespace nfk {

constexpr int WG_STAGES = 4;
constexpr uint32_t WG_HALF = 64 * 128;                 // 64 rows of a 128 x 64 bf16 block
constexpr uint32_t WG_STAGE_BYTES = 6 * WG_HALF;       // 2 A column blocks + up to 4 B column blocks
constexpr int WG_THREADS = 6 * 32;                     // producer, MMA issuer, 4 epilogue warps
constexpr size_t WG_SMEM = (size_t)WG_STAGES * WG_STAGE_BYTES + 16 * 8 + 1024;

struct WgArgs {
  const unsigned char* a_img;
  const unsigned char* b_img;
  float* c;
  long long ldc, m_tiles;
  int KBa, KBb;            // column blocks of the two images
  int P, Q;                // real output rows / columns
  int pad_p;               // output row index runs over 24-per-feature padded parameters (row 23 dropped)
  int n_qt, splits;
  int qnb[8], qb0[8];      // per Q tile: column blocks, first block
};

__device__ __forceinline__ bool wg_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// MN-major SWIZZLE_128B descriptor: leading byte offset = distance between 64-element MN groups,
// stride byte offset = distance between 8-row K groups (1024 B)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t mn_group_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(mn_group_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor with both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_ws_kernel(const __grid_constant__ WgArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + WG_STAGES;
  uint64_t* tfull = bars + 2 * WG_STAGES;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 32) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  // work item of this CTA: (P tile, Q tile, batch slice)
  int w = blockIdx.x;
  const int split = w % a.splits;
  w /= a.splits;
  const int qt = w % a.n_qt, pt = w / a.n_qt;
  const int nbq = a.qnb[qt], qb0 = a.qb0[qt];
  const int pa0 = 2 * pt, pa1 = min(2 * pt + 1, a.KBa - 1);       // second block clamped: its rows are discarded
  const long long per = (a.m_tiles + a.splits - 1) / a.splits;
  const long long mt0 = split * per, mt1 = min(a.m_tiles, mt0 + per);
  const long long steps = mt1 > mt0 ? 2 * (mt1 - mt0) : 0;        // half blocks (64 batch rows each)

  if (warp == 0) {
    // ---- producer
    uint32_t s = 0, ph = 0;
    for (long long st = 0; st < steps; ++st) {
      const long long mt = mt0 + (st >> 1);
      const uint32_t half = (uint32_t)(st & 1) * WG_HALF;
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) {
        unsigned char* dst = sm + s * WG_STAGE_BYTES;
        mbar_expect_tx(&full[s], (2 + nbq) * WG_HALF);
        const unsigned char* ab = a.a_img + ((size_t)mt * a.KBa) * (2 * WG_HALF) + half;
        bulk_g2s(dst, ab + (size_t)pa0 * (2 * WG_HALF), WG_HALF, &full[s]);
        bulk_g2s(dst + WG_HALF, ab + (size_t)pa1 * (2 * WG_HALF), WG_HALF, &full[s]);
        const unsigned char* bb = a.b_img + ((size_t)mt * a.KBb + qb0) * (2 * WG_HALF) + half;
        for (int j = 0; j < nbq; ++j) bulk_g2s(dst + (2 + j) * WG_HALF, bb + (size_t)j * (2 * WG_HALF), WG_HALF, &full[s]);
      }
      __syncwarp();
      if (++s == WG_STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    const uint32_t idesc = make_idesc_bf16_mn(128, nbq * 64);
    uint32_t s = 0, ph = 0;
    for (long long st = 0; st < steps; ++st) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (wg_elect_one()) {
        const uint32_t aa = smem_u32(sm + s * WG_STAGE_BYTES), ba = aa + 2 * WG_HALF;
#pragma unroll
        for (int k = 0; k < 4; ++k)                        // 16 batch rows = 2048 bytes per MMA
          umma_bf16(tmem, make_desc_mn_sw128(aa + k * 2048, WG_HALF), make_desc_mn_sw128(ba + k * 2048, WG_HALF),
                    idesc, (st | k) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (st == steps - 1) umma_commit(tfull);
      }
      __syncwarp();
      if (++s == WG_STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (steps > 0) {
    // ---- epilogue: partial tile -> C (+=)
    const int q = warp & 3;
    const int p_log = pt * 128 + q * 32 + lane;                  // output row of this thread
    int p_out = p_log;
    if (a.pad_p) p_out = (p_log % 24 == 23) ? -1 : (p_log / 24) * 23 + (p_log % 24);
    const bool row_ok = p_out >= 0 && p_out < a.P && (q * 32 + lane < 64 || 2 * pt + 1 < a.KBa);
    mbar_wait(tfull, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < nbq * 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (row_ok) {
        float* cr = a.c + (long long)p_out * a.ldc + qb0 * 64 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (qb0 * 64 + c0 + j < a.Q) atomicAdd(cr + j, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace nfk

using namespace nfk;

extern "C" int nfk_wgrad_ws(const void* a_img, const void* b_img, float* C, int64_t ldc, int64_t M, int KBa, int KBb,
                            int P, int Q, int pad_p, void* stream) {
  NFK_REQUIRE(M >= 0 && KBa > 0 && KBb > 0 && P > 0 && Q > 0 && ldc >= Q, "wgrad_ws: bad shape");
  NFK_REQUIRE(Q <= KBb * 64 && (pad_p ? P <= (KBa * 64 / 24) * 23 + 23 : P <= KBa * 64), "wgrad_ws: P/Q exceed the images");
  if (M == 0) return NFK_OK;
  NFK_REQUIRE(a_img && b_img && C, "wgrad_ws: null device pointer");
  NFK_REQUIRE(((reinterpret_cast<uintptr_t>(a_img) | reinterpret_cast<uintptr_t>(b_img)) & 15) == 0,
              "wgrad_ws: images must be 16-byte aligned");
  WgArgs a{};
  a.a_img = reinterpret_cast<const unsigned char*>(a_img);
  a.b_img = reinterpret_cast<const unsigned char*>(b_img);
  a.c = C;
  a.ldc = ldc;
  a.m_tiles = (M + 127) / 128;
  a.KBa = KBa;
  a.KBb = KBb;
  a.P = P;
  a.Q = Q;
  a.pad_p = pad_p;
  const int qblocks = (Q + 63) / 64;
  a.n_qt = (qblocks + 3) / 4;
  NFK_REQUIRE(a.n_qt <= 8, "wgrad_ws: at most 2048 output columns");
  {
    const int base = qblocks / a.n_qt, rem = qblocks % a.n_qt;
    int b0 = 0;
    for (int t = 0; t < a.n_qt; ++t) {
      a.qnb[t] = base + (t < rem ? 1 : 0);
      a.qb0[t] = b0;
      b0 += a.qnb[t];
    }
  }
  const int p_log = pad_p ? ((P + 22) / 23) * 24 : P;
  const int n_pt = (p_log + 127) / 128;
  const long long tiles = (long long)n_pt * a.n_qt;
  long long splits = sm_count() / tiles;
  if (splits < 1) splits = 1;
  if (splits > a.m_tiles) splits = a.m_tiles;
  a.splits = (int)splits;
  cudaError_t e = cudaFuncSetAttribute(wgrad_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM);
  if (e != cudaSuccess) {
    set_error("wgrad_ws: cannot set %zu B dynamic shared memory: %s", WG_SMEM, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  wgrad_ws_kernel<<<(unsigned)(tiles * splits), WG_THREADS, WG_SMEM, (cudaStream_t)stream>>>(a);
  count_launch();
  return check_launch("wgrad_ws");
}
