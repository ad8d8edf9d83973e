// Priors / targets evaluated on either side of the flow: the Einstein-crystal
// prior, the Lennard-Jones pair potential with minimum image + cutoff/shift, and the Gaussian
// mixture log-density of applications/src/systems.py — one pass over the batch each, per-sample
// reduction in registers / warp shuffles.
#include "nfk_common.cuh"

namespace nfk {

// minimum image: v -= (|v| > L/2) * sign(v) * L   (systems.py:155-158, :363-365)
__device__ __forceinline__ float min_image(float v, float L, float halfL) {
  if (L > 0.f && fabsf(v) > halfL) v -= (v > 0.f ? L : -L);
  return v;
}

// ---- EinsteinCrystal.log_prob (systems.py:360-366): sum over atoms of log N(dev; 0, I/alpha),
// dev = x - centers with the minimum-image wrap.  One warp per sample.
__global__ void __launch_bounds__(256)
einstein_logprob_kernel(const float* __restrict__ x, const float* __restrict__ centers, float* __restrict__ out,
                        float* __restrict__ gx, long long N, int nd, float alpha, float L, float cst_total) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float halfL = 0.5f * L;
  float acc = 0.f;
  for (int j = lane; j < nd; j += 32) {
    const float dev = min_image(x[row * nd + j] - centers[j], L, halfL);
    acc = fmaf(dev, dev, acc);
    if (gx) gx[row * nd + j] = -alpha * dev;                  // d log_prob / dx
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = fmaf(-0.5f * alpha, acc, cst_total);
}

// ---- LJ.potential (systems.py:154-189): pair_dist with minimum image, r, 1/r with the cutoff
// zeroing (r > cutoff -> 1/r := 0), 4 eps ((s/r)^12 - (s/r)^6 [- shift^2 + shift]) * (1/r * r), half the
// double sum.  One warp per sample, positions staged in shared memory.  With gpos != null also
// writes dU/dpos (analytic, inside the cutoff).
constexpr int LJ_MAXP = 128;      // particles per sample
constexpr int LJ_MAXDIM = 3;
template <int DIM>
__global__ void __launch_bounds__(128)
lj_potential_kernel(const float* __restrict__ pos, float* __restrict__ out, float* __restrict__ gpos, long long N,
                    int n, float L, float eps, float sigma, float cutoff, int shift, float s6, float s12) {
  __shared__ float sp[4][LJ_MAXP * LJ_MAXDIM];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + w;
  if (row >= N) return;
  float* p = sp[w];
  const float* src = pos + row * (long long)(n * DIM);
  for (int i = lane; i < n * DIM; i += 32) p[i] = src[i];
  __syncwarp();
  const float halfL = 0.5f * L;
  const bool has_cut = cutoff > 0.f;
  float acc = 0.f;
  for (int i = lane; i < n; i += 32) {
    float g[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) g[c] = 0.f;
    for (int j = 0; j < n; ++j) {
      float dv[DIM];
      float r2 = 0.f;
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        dv[c] = min_image(p[i * DIM + c] - p[j * DIM + c], L, halfL);
        r2 = fmaf(dv[c], dv[c], r2);
      }
      const float r = sqrtf(r2);
      const float scaled = r + (r == 0.f ? 1.f : 0.f);         // distances + (distances == 0)
      float inv = 1.f / scaled;
      if (has_cut && r > cutoff) inv = inv - inv;              // distances_inverse - (r > cutoff) * distances_inverse
      const float q = sigma * inv;
      const float q2 = q * q;
      const float pow6 = q2 * q2 * q2;
      float pair = has_cut && shift ? eps * 4.f * (pow6 * pow6 - pow6 - s12 + s6)
                                    : eps * 4.f * (pow6 * pow6 - pow6);
      pair = pair * inv * r;                                   // kills i == j and r > cutoff
      acc += pair;
      if (gpos && r > 0.f && inv > 0.f) {
        // dU/dr of 4 eps ((s/r)^12 - (s/r)^6) = 4 eps (-12 s^12/r^13 + 6 s^6/r^7); each pair counts once in U
        const float dUdr = 4.f * eps * (-12.f * pow6 * pow6 + 6.f * pow6) * inv;
#pragma unroll
        for (int c = 0; c < DIM; ++c) g[c] = fmaf(dUdr * inv, dv[c], g[c]);
      }
    }
    if (gpos) {
#pragma unroll
      for (int c = 0; c < DIM; ++c) gpos[row * (long long)(n * DIM) + i * DIM + c] = g[c];
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = 0.5f * acc;                        // sum(pair_potential) / 2
}

// ---- GaussianMixture.log_prob (systems.py:287-292): per point prob = sum_c exp(log N_c(x)) / nc
// (plain exp-sum, as the reference), per sample sum of log prob over its npoints points.
// One warp per sample.
__global__ void __launch_bounds__(256)
gmm_logprob_kernel(const float* __restrict__ x, const float* __restrict__ centers, const float* __restrict__ vars,
                   float* __restrict__ out, long long N, int npoints, int dim, int nc) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float acc = 0.f;
  const float inv_nc = 1.f / (float)nc;
  for (int pnt = lane; pnt < npoints; pnt += 32) {
    const float* xp = x + (row * npoints + pnt) * (long long)dim;
    float prob = 0.f;
    for (int c = 0; c < nc; ++c) {
      float m = 0.f;
      for (int k = 0; k < dim; ++k) {
        const float dlt = xp[k] - centers[c * dim + k];
        m = fmaf(dlt, dlt, m);
      }
      const float var = vars[c];
      // log N(x; mu, var I) = -0.5 |x-mu|^2 / var - 0.5 dim log(2 pi) - 0.5 dim log(var)
      const float lp = -0.5f * m / var - 0.5f * dim * 1.8378770664093453f - 0.5f * dim * logf(var);
      prob += inv_nc * expf(lp);
    }
    acc += logf(prob);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

}  // namespace nfk

using namespace nfk;

extern "C" {

int nfk_einstein_logprob(const float* x, const float* centers, float* out, float* grad_x, int64_t N, int natoms,
                         int dim, float alpha, float boxlength, void* stream) {
  NFK_REQUIRE(N >= 0 && natoms > 0 && dim > 0 && alpha > 0.f, "einstein_logprob: bad shape / alpha");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && centers && out, "einstein_logprob: null device pointer");
  const int nd = natoms * dim;
  // per atom: -0.5 dim log(2 pi) - 0.5 dim log(1/alpha)
  const float cst = (float)natoms * (-0.5f * dim * 1.8378770664093453f + 0.5f * dim * logf(alpha));
  const long long threads = (long long)N * 32;
  einstein_logprob_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, centers, out, grad_x, N, nd, alpha, boxlength > 0.f ? boxlength : 0.f, cst);
  count_launch();
  return check_launch("einstein_logprob");
}

int nfk_lj_potential(const float* pos, float* out, float* grad_pos, int64_t N, int nparticles, int dim,
                     float boxlength, float epsilon, float sigma, float cutoff, int shift, void* stream) {
  NFK_REQUIRE(N >= 0 && nparticles > 0 && nparticles <= LJ_MAXP, "lj_potential: 1..%d particles supported", LJ_MAXP);
  NFK_REQUIRE(dim == 2 || dim == 3, "lj_potential: dim must be 2 or 3");
  NFK_REQUIRE(sigma > 0.f, "lj_potential: sigma must be positive");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(pos && out, "lj_potential: null device pointer");
  const unsigned grid = (unsigned)((N + 3) / 4);
  cudaStream_t st = (cudaStream_t)stream;
  const float L = boxlength > 0.f ? boxlength : 0.f, cut = cutoff > 0.f ? cutoff : 0.f;
  // pow_6_shift = (sigma/cutoff)**6 and pow_6_shift**2 are Python doubles in the reference (systems.py:168-170)
  double sh6 = 0.0;
  if (cut > 0.f && shift) {
    const double sc = (double)sigma / (double)cutoff;
    sh6 = sc * sc * sc * sc * sc * sc;
  }
  const float s6 = (float)sh6, s12 = (float)(sh6 * sh6);
  if (dim == 3)
    lj_potential_kernel<3><<<grid, 128, 0, st>>>(pos, out, grad_pos, N, nparticles, L, epsilon, sigma, cut, shift, s6,
                                                 s12);
  else
    lj_potential_kernel<2><<<grid, 128, 0, st>>>(pos, out, grad_pos, N, nparticles, L, epsilon, sigma, cut, shift, s6,
                                                 s12);
  count_launch();
  return check_launch("lj_potential");
}

int nfk_gmm_logprob(const float* x, const float* centers, const float* vars, float* out, int64_t N, int npoints,
                    int dim, int ncenters, void* stream) {
  NFK_REQUIRE(N >= 0 && npoints > 0 && dim > 0 && ncenters > 0, "gmm_logprob: bad shape");
  if (N == 0) return NFK_OK;
  NFK_REQUIRE(x && centers && vars && out, "gmm_logprob: null device pointer");
  const long long threads = (long long)N * 32;
  gmm_logprob_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, centers, vars, out, N,
                                                                                        npoints, dim, ncenters);
  count_launch();
  return check_launch("gmm_logprob");
}

}  // extern "C"
