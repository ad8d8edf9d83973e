// Row-wise transforms around the spline kernel: affine half-coupling (RealNVP), planar stack,
// radial, Gaussian log-prob reduction, conditioner input gather, leapfrog updates.
// All are one pass over HBM with 128-bit accesses where the row length allows; per-row
// reductions are sub-warp shuffles.
#include "nfk_common.cuh"

namespace nfk {

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int pick_group(long long per_row_items) {
  int g = 1;
  while (g < 32 && g < per_row_items) g <<= 1;
  return g;
}

// ---- affine half-coupling (nf/flows.py:56, :59, :61-62 / :69, :72, :74-75) ---------------
template <int G>
__global__ void __launch_bounds__(256)
affine_half_kernel(const float* __restrict__ x, long long ld_x, int v_off,
                   const float* __restrict__ s, const float* __restrict__ t,
                   float* __restrict__ out, long long ld_out, int y_off,
                   float* __restrict__ logdet, long long N, int h, int inverse, int accumulate) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  const long long row = gid / G;
  const int g = (int)(gid % G);
  const bool live = row < N;
  float acc = 0.f;
  if (live) {
    const float* xr = x + row * ld_x + v_off;
    const float* sr = s + row * h;
    const float* tr = t + row * h;
    float* yr = out + row * ld_out + y_off;
    for (int j = g; j < h; j += G) {
      const float sv = sr[j], tv = tr[j], v = xr[j];
      // forward: t + v*exp(s); inverse: (v - t)*exp(-s), each product rounded as in the reference
      yr[j] = inverse ? __fmul_rn(__fsub_rn(v, tv), expf(-sv)) : __fadd_rn(tv, __fmul_rn(v, expf(sv)));
      acc += sv;
    }
  }
  acc = group_sum<G>(acc);
  if (live && g == 0) {
    const float v = inverse ? -acc : acc;
    logdet[row] = accumulate ? logdet[row] + v : v;
  }
}

// backward of y = t + v e^s (fwd) / y = (v - t) e^{-s} (inv), logdet = +-sum s
template <int G>
__global__ void __launch_bounds__(256)
affine_half_bwd_kernel(const float* __restrict__ x, long long ld_x, int v_off,
                       const float* __restrict__ s, const float* __restrict__ t,
                       const float* __restrict__ gy, long long ld_gy, int gy_off,
                       const float* __restrict__ gld, float* __restrict__ gv, long long ld_gv,
                       int gv_off, float* __restrict__ gs, float* __restrict__ gt, long long N,
                       int h, int inverse) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  const long long row = gid / G;
  const int g = (int)(gid % G);
  if (row >= N) return;
  const float* xr = x + row * ld_x + v_off;
  const float* sr = s + row * h;
  const float* tr = t + row * h;
  const float* gyr = gy + row * ld_gy + gy_off;
  float* gvr = gv + row * ld_gv + gv_off;
  const float gl = gld ? gld[row] : 0.f;
  for (int j = g; j < h; j += G) {
    const float sv = sr[j], tv = tr[j], v = xr[j], go = gyr[j];
    if (!inverse) {
      const float e = expf(sv);
      gvr[j] = go * e;
      gt[row * h + j] = go;
      gs[row * h + j] = go * v * e + gl;
    } else {
      const float e = expf(-sv);
      gvr[j] = go * e;
      gt[row * h + j] = -go * e;
      gs[row * h + j] = -go * (v - tv) * e - gl;
    }
  }
}

// ---- planar stack (nf/flows_1.py:42-60, quirk Q8) ------------------------------------------
// uhat_l and w_l.uhat_l depend only on the parameters: a prologue kernel computes them once.
__global__ void planar_prepare_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                      float* __restrict__ uhat, float* __restrict__ wuhat, int d) {
  // one warp per layer
  const int l = blockIdx.x, lane = threadIdx.x;
  const float* wl = w + (size_t)l * d;
  const float* ul = u + (size_t)l * d;
  float wu = 0.f, ww = 0.f;
  for (int j = lane; j < d; j += 32) {
    wu += wl[j] * ul[j];
    ww += wl[j] * wl[j];
  }
  wu = warp_sum(wu);
  ww = warp_sum(ww);
  // flows_1.py:52-53: scal = log(1 + exp(w.u)) - w.u - 1; uhat = u + scal * w / ||w||^2
  const float scal = logf(1.f + expf(wu)) - wu - 1.f;
  const float nrm = sqrtf(ww);
  const float n2 = nrm * nrm;
  float dot = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float uh = ul[j] + scal * wl[j] / n2;
    uhat[(size_t)l * d + j] = uh;
    dot += wl[j] * uh;
  }
  dot = warp_sum(dot);
  if (lane == 0) wuhat[l] = dot;
}

// G lanes share one row, NV values per lane at columns g + G*i: x stays in registers across all
// L layers; w / uhat of the current layer are read from shared memory.
template <int G, int NV>
__global__ void __launch_bounds__(256)
planar_stack_kernel(const float* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ uhat, const float* __restrict__ wuhat,
                    const float* __restrict__ b, float* __restrict__ out,
                    float* __restrict__ logdet, long long N, int d, int L, int accumulate) {
  extern __shared__ float sm[];
  float* sw = sm;                         // [L][d]
  float* su = sm + (size_t)L * d;         // [L][d]
  float* sb = su + (size_t)L * d;         // [L]
  float* sd = sb + L;                     // [L]  w.uhat
  for (int i = threadIdx.x; i < L * d; i += blockDim.x) {
    sw[i] = w[i];
    su[i] = uhat[i];
  }
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    sb[i] = b[i];
    sd[i] = wuhat[i];
  }
  __syncthreads();
  const int g = threadIdx.x % G;
  const long long rows_per_block = blockDim.x / G;
  for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < N;
       row0 += (long long)gridDim.x * rows_per_block) {
    const long long row = row0 + threadIdx.x / G;
    const bool live = row < N;                 // whole groups stay in the loop for the shuffles
    float xv[NV];
    const float* xr = x + (live ? row : 0) * d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = g + G * i;
      xv[i] = col < d ? xr[col] : 0.f;
    }
    float ld = 0.f;
    for (int l = 0; l < L; ++l) {
      const float* wl = sw + (size_t)l * d;
      const float* ul = su + (size_t)l * d;
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        if (col < d) dot = fmaf(xv[i], wl[col], dot);
      }
      dot = group_sum<G>(dot);
      const float th = tanhf(dot + sb[l]);                             // flows_1.py:56-57
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        if (col < d) xv[i] = fmaf(ul[col], th, xv[i]);                 // flows_1.py:57
      }
      // phi.uhat = (1 - th^2) * (w.uhat); log(|1 + .| + 1e-4)            flows_1.py:58-59
      ld += logf(fabsf(1.f + (1.f - th * th) * sd[l]) + 1e-4f);
    }
    if (live) {
      float* orow = out + row * d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        if (col < d) orow[col] = xv[i];
      }
      if (g == 0) logdet[row] = accumulate ? logdet[row] + ld : ld;
    }
  }
}

// backward of the fused stack.  Forward is recomputed keeping tanh(lin_l) per layer (L <= 64);
// going back, x_l = x_{l+1} - uhat_l th_l reconstructs each layer's input.  Parameter gradients
// are reduced over the rows of a warp with shuffles, then accumulated with atomics.
constexpr int PL_MAXL = 64;
template <int G, int NV>
__global__ void __launch_bounds__(256)
planar_stack_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                        const float* __restrict__ uhat, const float* __restrict__ wuhat,
                        const float* __restrict__ b, const float* __restrict__ gout,
                        const float* __restrict__ gld, float* __restrict__ gx,
                        float* __restrict__ gw, float* __restrict__ guhat, float* __restrict__ gb,
                        float* __restrict__ gwuhat, long long N, int d, int L) {
  extern __shared__ float sm[];
  float* sw = sm;
  float* su = sm + (size_t)L * d;
  float* sb = su + (size_t)L * d;
  float* sd = sb + L;
  for (int i = threadIdx.x; i < L * d; i += blockDim.x) {
    sw[i] = w[i];
    su[i] = uhat[i];
  }
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    sb[i] = b[i];
    sd[i] = wuhat[i];
  }
  __syncthreads();
  const int g = threadIdx.x % G;
  const long long rows_per_block = blockDim.x / G;
  for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < N;
       row0 += (long long)gridDim.x * rows_per_block) {
    const long long row = row0 + threadIdx.x / G;
    const bool live = row < N;
    float xv[NV], gz[NV];
    float th[PL_MAXL];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = g + G * i;
      const bool ok = live && col < d;
      xv[i] = ok ? x[row * d + col] : 0.f;
      gz[i] = ok ? gout[row * d + col] : 0.f;
    }
    const float gl = (live && gld) ? gld[row] : 0.f;
    for (int l = 0; l < L; ++l) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        if (col < d) dot = fmaf(xv[i], sw[l * d + col], dot);
      }
      dot = group_sum<G>(dot);
      const float t = tanhf(dot + sb[l]);
      th[l] = t;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        if (col < d) xv[i] = fmaf(su[l * d + col], t, xv[i]);
      }
    }
    for (int l = L - 1; l >= 0; --l) {
      const float t = th[l];
      const float psi = 1.f - t * t;
      const float qv = 1.f + psi * sd[l];
      const float dld_dq = (qv >= 0.f ? 1.f : -1.f) / (fabsf(qv) + 1e-4f);
      float gth = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        if (col < d) {
          xv[i] = fmaf(-su[l * d + col], t, xv[i]);            // input of layer l
          gth = fmaf(gz[i], su[l * d + col], gth);
        }
      }
      gth = group_sum<G>(gth);
      gth += gl * dld_dq * (-2.f * t) * sd[l];
      const float glin = live ? gth * psi : 0.f;
      // parameter gradients: reduce over the row-groups of this warp, then one atomic per column
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        float a_w = glin * xv[i], a_u = live ? gz[i] * t : 0.f;
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
          a_w += __shfl_xor_sync(0xffffffffu, a_w, o);
          a_u += __shfl_xor_sync(0xffffffffu, a_u, o);
        }
        if (col < d && (threadIdx.x & 31) < G) {
          atomicAdd(gw + (size_t)l * d + col, a_w);
          atomicAdd(guhat + (size_t)l * d + col, a_u);
        }
      }
      float a_b = (g == 0) ? glin : 0.f, a_q = (g == 0 && live) ? gl * dld_dq * psi : 0.f;
      a_b = warp_sum(a_b);
      a_q = warp_sum(a_q);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(gb + l, a_b);
        atomicAdd(gwuhat + l, a_q);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        if (col < d) gz[i] = fmaf(glin, sw[l * d + col], gz[i]);   // grad w.r.t. layer input
      }
    }
    if (live) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = g + G * i;
        if (col < d) gx[row * d + col] = gz[i];
      }
    }
  }
}

// ---- radial (nf/flows_1.py:85-97, quirk Q9) -------------------------------------------------
__global__ void __launch_bounds__(256)
radial_sumsq_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                    float* __restrict__ sumsq, long long total, int d) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float df = x[i] - x0[i % d];
    acc = fmaf(df, df, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(sumsq, v);
  }
}

template <int G>
__global__ void __launch_bounds__(256)
radial_kernel(const float* __restrict__ x, const float* __restrict__ x0,
              const float* __restrict__ log_alpha, const float* __restrict__ beta_raw,
              const float* __restrict__ sumsq, float* __restrict__ out,
              float* __restrict__ logdet, long long N, int d, int per_sample, int accumulate) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  const long long row = gid / G;
  const int g = (int)(gid % G);
  const bool live = row < N;
  const float alpha = expf(log_alpha[0]);
  const float beta = -alpha + logf(1.f + expf(beta_raw[0]));           // flows_1.py:92
  float r;
  if (per_sample) {
    float acc = 0.f;
    if (live)
      for (int j = g; j < d; j += G) {
        const float df = x[row * d + j] - x0[j];
        acc = fmaf(df, df, acc);
      }
    r = sqrtf(group_sum<G>(acc));
  } else {
    r = sqrtf(sumsq[0]);                                               // flows_1.py:90
  }
  const float h = 1.f / (alpha + r);
  const float bh = beta * h;
  if (live) {
    for (int j = g; j < d; j += G) {
      const float xv = x[row * d + j];
      out[row * d + j] = xv + bh * (xv - x0[j]);                       // flows_1.py:93
    }
  }
  // flows_1.py:94-95
  const float ar = alpha + r;
  const float ld = (float)(d - 1) * logf(1.f + bh) + logf(1.f + bh - beta * r / (ar * ar));
  if (per_sample) {
    if (live && g == 0) logdet[row] = accumulate ? logdet[row] + ld : ld;
  } else if (gid == 0) {
    logdet[0] = accumulate ? logdet[0] + ld : ld;
  }
}

// backward of the radial layer.  dot[0] must hold sum(gz * (x - x0)) over the whole batch in the
// batch-global mode (radial_dot_kernel).  Parameter gradients are accumulated with atomics into
// zero-initialised gx0 [d], gla [1], gbeta [1].
__global__ void __launch_bounds__(256)
radial_dot_kernel(const float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ gz,
                  float* __restrict__ dot, long long total, int d) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    acc = fmaf(gz[i], x[i] - x0[i % d], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(dot, v);
  }
}

// Persistent over rows (grid-stride): a thread keeps its columns' share of grad_x0 in registers
// (d <= 8 G) or in shared-memory atomics (wider rows), the block folds them in shared memory and
// issues ONE global atomic per column -- not one per element.
template <int G>
__global__ void __launch_bounds__(256)
radial_bwd_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                  const float* __restrict__ log_alpha, const float* __restrict__ beta_raw,
                  const float* __restrict__ sumsq, const float* __restrict__ dot,
                  const float* __restrict__ gz, const float* __restrict__ gld, float* __restrict__ gx,
                  float* __restrict__ gx0, float* __restrict__ gla, float* __restrict__ gbeta,
                  long long N, int d, int per_sample) {
  extern __shared__ float s_gx0[];                           // [d]
  for (int i = threadIdx.x; i < d; i += blockDim.x) s_gx0[i] = 0.f;
  __syncthreads();
  constexpr int RPB = 256 / G;                               // rows per block and pass
  const int g = threadIdx.x % G, lr = threadIdx.x / G;
  const bool in_regs = d <= 8 * G;
  float acc0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float a_acc = 0.f, b_acc = 0.f;
  const float alpha = expf(log_alpha[0]);
  const float braw = beta_raw[0];
  const float beta = -alpha + logf(1.f + expf(braw));
  const float n1 = (float)(d - 1);
  for (long long row0 = (long long)blockIdx.x * RPB; row0 < N; row0 += (long long)gridDim.x * RPB) {
    const long long row = row0 + lr;
    const bool live = row < N;
    float r, S1, gl;
    if (per_sample) {
      float a2 = 0.f, a1 = 0.f;
      if (live)
        for (int j = g; j < d; j += G) {
          const float df = x[row * d + j] - x0[j];
          a2 = fmaf(df, df, a2);
          a1 = fmaf(gz[row * d + j], df, a1);
        }
      r = sqrtf(group_sum<G>(a2));
      S1 = group_sum<G>(a1);
      gl = (live && gld) ? gld[row] : 0.f;
    } else {
      r = sqrtf(sumsq[0]);
      S1 = dot[0];
      gl = gld ? gld[0] : 0.f;
    }
    const float A = alpha + r, h = 1.f / A;
    const float p = 1.f + beta * h, qv = 1.f + beta * h - beta * r * h * h;
    // ld = (n-1) log p + log q  (flows_1.py:94-95)
    float g_h = beta * S1 + gl * (n1 * beta / p + (beta - 2.f * beta * r * h) / qv);
    float g_b = h * S1 + gl * (n1 * h / p + (h - r * h * h) / qv);
    float g_r = gl * (-beta * h * h) / qv;
    const float g_A = -g_h * h * h;
    g_r += g_A;
    const float g_alpha = g_A - g_b;                           // beta = -alpha + softplus(braw)
    const float g_braw = g_b / (1.f + expf(-braw));
    const float scale = 1.f + beta * h;
    const float rinv = r > 0.f ? 1.f / r : 0.f;
    if (live) {
      if (in_regs) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int j = g + c * G;
          if (j < d) {
            const float df = x[row * d + j] - x0[j];
            const float gzv = gz[row * d + j];
            gx[row * d + j] = gzv * scale + g_r * df * rinv;
            // z = x + beta h (x - x0): d/dx0 = -(beta h gz) - g_r df / r
            acc0[c] += -(gzv * beta * h) - g_r * df * rinv;
          }
        }
      } else {
        for (int j = g; j < d; j += G) {
          const float df = x[row * d + j] - x0[j];
          const float gzv = gz[row * d + j];
          gx[row * d + j] = gzv * scale + g_r * df * rinv;
          atomicAdd(&s_gx0[j], -(gzv * beta * h) - g_r * df * rinv);
        }
      }
    }
    if (per_sample) {
      if (live && g == 0) {
        a_acc += g_alpha * alpha;
        b_acc += g_braw;
      }
    } else if (row0 == 0 && threadIdx.x == 0) {              // batch-global: one value for the whole batch
      a_acc = g_alpha * alpha;
      b_acc = g_braw;
    }
  }
  if (in_regs) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = g + c * G;
      if (j < d) atomicAdd(&s_gx0[j], acc0[c]);
    }
  }
  a_acc = warp_sum(a_acc);
  b_acc = warp_sum(b_acc);
  if ((threadIdx.x & 31) == 0 && (a_acc != 0.f || b_acc != 0.f)) {
    atomicAdd(gla, a_acc);
    atomicAdd(gbeta, b_acc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) atomicAdd(gx0 + i, s_gx0[i]);
}

// ---- log-prob reduction (nf/models.py:19-20, :34, :39) --------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
gauss_logprob_kernel(const float* __restrict__ z, const float* __restrict__ add, float add_sign,
                     float* __restrict__ out, long long N, int d, float inv_var, float cst) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  const long long row = gid / G;
  const int g = (int)(gid % G);
  const bool live = row < N;
  float acc = 0.f;
  if (live) {
    const float* zr = z + row * d;
    if ((d & 3) == 0) {
      const float4* z4 = reinterpret_cast<const float4*>(zr);
      for (int j = g; j < (d >> 2); j += G) {
        const float4 v = ldg_stream4(z4 + j);
        acc = fmaf(v.x, v.x, acc);
        acc = fmaf(v.y, v.y, acc);
        acc = fmaf(v.z, v.z, acc);
        acc = fmaf(v.w, v.w, acc);
      }
    } else {
      for (int j = g; j < d; j += G) acc = fmaf(zr[j], zr[j], acc);
    }
  }
  acc = group_sum<G>(acc);
  if (live && g == 0) {
    float v = -0.5f * acc * inv_var + cst;
    if (add) v += add_sign * add[row];
    out[row] = v;
  }
}

// ---- conditioner input gather (nf/flows.py:230): x[:, :, cols].flatten(1) --------------------
template <bool BF16>
__global__ void __launch_bounds__(256)
gather_cols_kernel(const float* __restrict__ x, void* __restrict__ out, long long N, int size,
                   int dim, int n_cols, int c0, int c1, int c2, int c3, long long ld_out) {
  const int per_row = size * n_cols;
  const long long total = N * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / per_row;
    const int rem = (int)(i - row * per_row);
    const int s = rem / n_cols, ci = rem - s * n_cols;
    const int col = ci == 0 ? c0 : ci == 1 ? c1 : ci == 2 ? c2 : c3;
    const float v = x[row * (long long)(size * dim) + s * dim + col];
    if (BF16)
      reinterpret_cast<__nv_bfloat16*>(out)[row * ld_out + rem] = __float2bfloat16_rn(v);
    else
      reinterpret_cast<float*>(out)[row * ld_out + rem] = v;
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// ---- leapfrog (velocity Verlet; pattern of applications/src/systems.py:331-336) --------------
__global__ void __launch_bounds__(256)
kick_drift_kernel(float* __restrict__ q, float* __restrict__ p, const float* __restrict__ f,
                  long long n, float dt, float inv_mass) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float pv = fmaf(0.5f * dt, f[i], p[i]);
    p[i] = pv;
    q[i] = fmaf(dt * inv_mass, pv, q[i]);
  }
}
__global__ void __launch_bounds__(256)
kick_kernel(float* __restrict__ p, const float* __restrict__ f, long long n, float dt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = fmaf(0.5f * dt, f[i], p[i]);
}

static unsigned ew_grid(long long n, int per_block) {
  long long g = (n + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace nfk

using namespace nfk;

#define NFK_GROUP_SWITCH(G, CALL)        \
  switch (G) {                           \
    case 1: { constexpr int GG = 1; CALL; } break;   \
    case 2: { constexpr int GG = 2; CALL; } break;   \
    case 4: { constexpr int GG = 4; CALL; } break;   \
    case 8: { constexpr int GG = 8; CALL; } break;   \
    case 16: { constexpr int GG = 16; CALL; } break; \
    default: { constexpr int GG = 32; CALL; } break; \
  }

extern "C" {

int nfk_affine_halfcoupling(const float* x, int64_t ld_x, int v_off, const float* s,
                            const float* t, float* out, int64_t ld_out, int y_off,
                            float* logdet, int64_t N, int h, int inverse, int accumulate,
                            void* stream) {
  NFK_REQUIRE(N >= 0 && h > 0, "affine_halfcoupling: bad shape");
  NFK_REQUIRE(v_off >= 0 && y_off >= 0 && ld_x >= v_off + h && ld_out >= y_off + h,
              "affine_halfcoupling: column window outside the row");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && s && t && out && logdet, "affine_halfcoupling: null device pointer");
  const int G = pick_group(h);
  const long long threads = N * G;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  NFK_GROUP_SWITCH(G, (affine_half_kernel<GG><<<grid, 256, 0, st>>>(
                          x, ld_x, v_off, s, t, out, ld_out, y_off, logdet, N, h, inverse,
                          accumulate)));
  count_launch();
  return check_launch("affine_halfcoupling");
}

int nfk_affine_halfcoupling_bwd(const float* x, int64_t ld_x, int v_off, const float* s,
                                const float* t, const float* grad_y, int64_t ld_gy, int gy_off,
                                const float* grad_logdet, float* grad_v, int64_t ld_gv,
                                int gv_off, float* grad_s, float* grad_t, int64_t N, int h,
                                int inverse, void* stream) {
  NFK_REQUIRE(N >= 0 && h > 0, "affine_halfcoupling_bwd: bad shape");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && s && t && grad_y && grad_v && grad_s && grad_t,
              "affine_halfcoupling_bwd: null device pointer");
  const int G = pick_group(h);
  const long long threads = N * G;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  NFK_GROUP_SWITCH(G, (affine_half_bwd_kernel<GG><<<grid, 256, 0, st>>>(
                          x, ld_x, v_off, s, t, grad_y, ld_gy, gy_off, grad_logdet, grad_v, ld_gv,
                          gv_off, grad_s, grad_t, N, h, inverse)));
  count_launch();
  return check_launch("affine_halfcoupling_bwd");
}

int nfk_planar_prepare(const float* w, const float* u, float* uhat, float* wuhat, int d, int L,
                       void* stream) {
  NFK_REQUIRE(d > 0 && L > 0, "planar_prepare: bad shape");
  NFK_REQUIRE(w && u && uhat && wuhat, "planar_prepare: null device pointer");
  planar_prepare_kernel<<<L, 32, 0, (cudaStream_t)stream>>>(w, u, uhat, wuhat, d);
  count_launch();
  return check_launch("planar_prepare");
}

int nfk_planar_stack(const float* x, const float* w, const float* uhat, const float* wuhat,
                     const float* b, float* out, float* logdet, int64_t N, int d, int L,
                     int accumulate, void* stream) {
  NFK_REQUIRE(N >= 0 && d > 0 && L > 0, "planar_stack: bad shape");
  NFK_REQUIRE(d <= 1024, "planar_stack: d <= 1024 supported (got %d)", d);
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && w && uhat && wuhat && b && out && logdet, "planar_stack: null device pointer");
  const size_t smem = ((size_t)2 * L * d + 2 * L) * sizeof(float);
  NFK_REQUIRE(smem <= 200 * 1024, "planar_stack: L*d = %d too large for shared memory; split the stack",
              L * d);
  cudaStream_t st = (cudaStream_t)stream;
  // G lanes x NV values cover the row
  int G, NV;
  if (d <= 8) { G = 1; NV = 8; }
  else if (d <= 32) { G = 4; NV = 8; }
  else if (d <= 128) { G = 8; NV = 16; }
  else if (d <= 512) { G = 32; NV = 16; }
  else { G = 32; NV = 32; }
  const long long rows_per_block = 256 / G;
  long long grid = (N + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)sm_count() * 4;
  if (grid > cap) grid = cap;
#define NFK_PL(GG, VV)                                                                         \
  do {                                                                                         \
    cudaFuncSetAttribute(planar_stack_kernel<GG, VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)smem);                                                           \
    planar_stack_kernel<GG, VV><<<(unsigned)grid, 256, smem, st>>>(x, w, uhat, wuhat, b, out,  \
                                                                   logdet, N, d, L, accumulate); \
  } while (0)
  if (G == 1) NFK_PL(1, 8);
  else if (G == 4) NFK_PL(4, 8);
  else if (G == 8) NFK_PL(8, 16);
  else if (NV == 16) NFK_PL(32, 16);
  else NFK_PL(32, 32);
#undef NFK_PL
  count_launch();
  return check_launch("planar_stack");
}

int nfk_planar_stack_bwd(const float* x, const float* w, const float* uhat, const float* wuhat,
                         const float* b, const float* grad_out, const float* grad_logdet,
                         float* grad_x, float* grad_w, float* grad_uhat, float* grad_b,
                         float* grad_wuhat, int64_t N, int d, int L, void* stream) {
  NFK_REQUIRE(N >= 0 && d > 0 && L > 0, "planar_stack_bwd: bad shape");
  NFK_REQUIRE(d <= 1024 && L <= PL_MAXL, "planar_stack_bwd: d <= 1024 and L <= %d supported", PL_MAXL);
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && w && uhat && wuhat && b && grad_out && grad_x && grad_w && grad_uhat && grad_b && grad_wuhat,
              "planar_stack_bwd: null device pointer");
  const size_t smem = ((size_t)2 * L * d + 2 * L) * sizeof(float);
  NFK_REQUIRE(smem <= 200 * 1024, "planar_stack_bwd: L*d too large for shared memory; split the stack");
  cudaStream_t st = (cudaStream_t)stream;
  int G, NV;
  if (d <= 8) { G = 1; NV = 8; }
  else if (d <= 32) { G = 4; NV = 8; }
  else if (d <= 128) { G = 8; NV = 16; }
  else if (d <= 512) { G = 32; NV = 16; }
  else { G = 32; NV = 32; }
  const long long rows_per_block = 256 / G;
  long long grid = (N + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)sm_count() * 2;
  if (grid > cap) grid = cap;
#define NFK_PLB(GG, VV)                                                                            \
  do {                                                                                             \
    cudaFuncSetAttribute(planar_stack_bwd_kernel<GG, VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)smem);                                                               \
    planar_stack_bwd_kernel<GG, VV><<<(unsigned)grid, 256, smem, st>>>(                            \
        x, w, uhat, wuhat, b, grad_out, grad_logdet, grad_x, grad_w, grad_uhat, grad_b, grad_wuhat, N, d, L); \
  } while (0)
  if (G == 1) NFK_PLB(1, 8);
  else if (G == 4) NFK_PLB(4, 8);
  else if (G == 8) NFK_PLB(8, 16);
  else if (NV == 16) NFK_PLB(32, 16);
  else NFK_PLB(32, 32);
#undef NFK_PLB
  count_launch();
  return check_launch("planar_stack_bwd");
}

int nfk_radial_sumsq(const float* x, const float* x0, float* sumsq, int64_t N, int d,
                     void* stream) {
  NFK_REQUIRE(N >= 0 && d > 0, "radial_sumsq: bad shape");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && x0 && sumsq, "radial_sumsq: null device pointer");
  const long long total = (long long)N * d;
  radial_sumsq_kernel<<<ew_grid(total, 256 * 8), 256, 0, (cudaStream_t)stream>>>(x, x0, sumsq,
                                                                               total, d);
  count_launch();
  return check_launch("radial_sumsq");
}

int nfk_radial(const float* x, const float* x0, const float* log_alpha, const float* beta,
               const float* sumsq, float* out, float* logdet, int64_t N, int d, int per_sample,
               int accumulate, void* stream) {
  NFK_REQUIRE(N >= 0 && d > 0, "radial: bad shape");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && x0 && log_alpha && beta && out && logdet, "radial: null device pointer");
  NFK_REQUIRE(per_sample || sumsq, "radial: batch-global mode needs sumsq");
  const int G = pick_group(d);
  const long long threads = (long long)N * G;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  NFK_GROUP_SWITCH(G, (radial_kernel<GG><<<grid, 256, 0, st>>>(x, x0, log_alpha, beta, sumsq, out,
                                                                logdet, N, d, per_sample,
                                                                accumulate)));
  count_launch();
  return check_launch("radial");
}

int nfk_radial_dot(const float* x, const float* x0, const float* grad_out, float* dot, int64_t N, int d,
                   void* stream) {
  NFK_REQUIRE(N >= 0 && d > 0, "radial_dot: bad shape");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && x0 && grad_out && dot, "radial_dot: null device pointer");
  const long long total = (long long)N * d;
  radial_dot_kernel<<<ew_grid(total, 256 * 8), 256, 0, (cudaStream_t)stream>>>(x, x0, grad_out, dot, total, d);
  count_launch();
  return check_launch("radial_dot");
}

int nfk_radial_bwd(const float* x, const float* x0, const float* log_alpha, const float* beta,
                   const float* sumsq, const float* dot, const float* grad_out, const float* grad_logdet,
                   float* grad_x, float* grad_x0, float* grad_log_alpha, float* grad_beta, int64_t N, int d,
                   int per_sample, void* stream) {
  NFK_REQUIRE(N >= 0 && d > 0, "radial_bwd: bad shape");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && x0 && log_alpha && beta && grad_out && grad_x && grad_x0 && grad_log_alpha && grad_beta,
              "radial_bwd: null device pointer");
  NFK_REQUIRE(per_sample || (sumsq && dot), "radial_bwd: batch-global mode needs sumsq and dot");
  const int G = pick_group(d);
  NFK_REQUIRE((size_t)d * sizeof(float) <= 48 * 1024, "radial_bwd: d = %d exceeds the shared-memory accumulator", d);
  const long long blocks_needed = ((long long)N * G + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  const unsigned grid = (unsigned)(blocks_needed < cap ? blocks_needed : cap);
  cudaStream_t st = (cudaStream_t)stream;
  NFK_GROUP_SWITCH(G, (radial_bwd_kernel<GG><<<grid, 256, d * sizeof(float), st>>>(x, x0, log_alpha, beta, sumsq, dot, grad_out,
                                                                    grad_logdet, grad_x, grad_x0, grad_log_alpha,
                                                                    grad_beta, N, d, per_sample)));
  count_launch();
  return check_launch("radial_bwd");
}

int nfk_gauss_logprob(const float* z, const float* add, float add_sign, float* out, int64_t N,
                      int d, float var, void* stream) {
  NFK_REQUIRE(N >= 0 && d > 0 && var > 0.f, "gauss_logprob: bad shape");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(z && out, "gauss_logprob: null device pointer");
  const int G = pick_group((d & 3) == 0 ? d / 4 : d);
  const long long threads = (long long)N * G;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  const float cst = (float)(-0.5 * (double)d * log(2.0 * 3.14159265358979323846 * (double)var));
  cudaStream_t st = (cudaStream_t)stream;
  NFK_GROUP_SWITCH(G, (gauss_logprob_kernel<GG><<<grid, 256, 0, st>>>(z, add, add_sign, out, N, d,
                                                                       1.f / var, cst)));
  count_launch();
  return check_launch("gauss_logprob");
}

int nfk_gather_cols(const float* x, void* out, int64_t N, int size, int dim, const int32_t* cols,
                    int n_cols, int out_bf16, int64_t ld_out, void* stream) {
  NFK_REQUIRE(N >= 0 && size > 0 && dim > 0, "gather_cols: bad shape");
  NFK_REQUIRE(cols && n_cols >= 1 && n_cols <= 4, "gather_cols: 1..4 columns supported");
  NFK_REQUIRE(ld_out >= (int64_t)size * n_cols, "gather_cols: ld_out too small");
  for (int i = 0; i < n_cols; ++i)
    NFK_REQUIRE(cols[i] >= 0 && cols[i] < dim, "gather_cols: column %d outside [0,%d)", cols[i], dim);
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && out, "gather_cols: null device pointer");
  const int c0 = cols[0], c1 = n_cols > 1 ? cols[1] : 0, c2 = n_cols > 2 ? cols[2] : 0,
            c3 = n_cols > 3 ? cols[3] : 0;
  const long long total = (long long)N * size * n_cols;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_bf16)
    gather_cols_kernel<true><<<ew_grid(total, 256 * 4), 256, 0, st>>>(x, out, N, size, dim, n_cols,
                                                                     c0, c1, c2, c3, ld_out);
  else
    gather_cols_kernel<false><<<ew_grid(total, 256 * 4), 256, 0, st>>>(x, out, N, size, dim, n_cols,
                                                                      c0, c1, c2, c3, ld_out);
  count_launch();
  return check_launch("gather_cols");
}

int nfk_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  if (n <= 0) return NFK_OK;
  NFK_REQUIRE(in && out, "cast_f32_bf16: null device pointer");
  cast_bf16_kernel<<<ew_grid(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out), n);
  count_launch();
  return check_launch("cast_f32_bf16");
}

int nfk_leapfrog_kick_drift(float* q, float* p, const float* force, int64_t n, float dt,
                            float inv_mass, void* stream) {
  if (n <= 0) return NFK_OK;
  NFK_REQUIRE(q && p && force, "leapfrog_kick_drift: null device pointer");
  kick_drift_kernel<<<ew_grid(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(q, p, force, n, dt,
                                                                         inv_mass);
  count_launch();
  return check_launch("leapfrog_kick_drift");
}

int nfk_leapfrog_kick(float* p, const float* force, int64_t n, float dt, void* stream) {
  if (n <= 0) return NFK_OK;
  NFK_REQUIRE(p && force, "leapfrog_kick: null device pointer");
  kick_kernel<<<ew_grid(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(p, force, n, dt);
  count_launch();
  return check_launch("leapfrog_kick");
}

}  // extern "C"
