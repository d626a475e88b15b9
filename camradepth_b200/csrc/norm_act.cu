// GroupNorm as a three-step protocol: per-(b,c) sums -> finalize into a per-(b,c)
// affine -> apply (optionally fused with GELU / Dropout2d scale) in one streaming pass; and the
// matching two-pass backward.  All kernels are HBM-bound: 16-byte vector accesses, channel-major
// thread mapping (coalesced NHWC rows), per-thread register accumulation, one atomic per
// (block, channel).
#include "common.cuh"
#include "chan_reduce.cuh"
#include "../../include/camradepth_b200.h"

namespace {

template <typename T>
__global__ void chan_stats_kernel(const T* __restrict__ x, float* sums, int B, long long N, int C, int ld,
                                  long long ppb) {
  chan_reduce2([&](int b, long long p, int c, float (&s0)[8], float (&s1)[8]) {
    float v[8];
    load8(x + ((long long)b * N + p) * ld + c, v);
#pragma unroll
    for (int j = 0; j < 8; j++) { s0[j] += v[j]; s1[j] = fmaf(v[j], v[j], s1[j]); }
  }, sums, B, N, C, ppb);
}

__global__ void gn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ ab,
                                   float* __restrict__ mean_rstd, float* __restrict__ xbar, int B, int C, int G,
                                   long long N, float eps) {
  const int b = blockIdx.x;
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    double s = 0.0, ss = 0.0;
    for (int j = 0; j < cpg; j++) {
      const float* p = sums + ((long long)b * C + g * cpg + j) * 2;
      s += (double)p[0]; ss += (double)p[1];
    }
    const double cnt = (double)cpg * (double)N;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = gamma[c] * rstd;
    const float bb = beta[c] - (float)mean * a;
    ab[((long long)b * C + c) * 2 + 0] = a;
    ab[((long long)b * C + c) * 2 + 1] = bb;
    if (xbar) xbar[(long long)b * C + c] = a * (sums[((long long)b * C + c) * 2] / (float)N) + bb;
    if (c == g * cpg) {
      mean_rstd[((long long)b * G + g) * 2 + 0] = (float)mean;
      mean_rstd[((long long)b * G + g) * 2 + 1] = rstd;
    }
  }
}

__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == CRD_ACT_GELU) return gelu_f(z);
  if (act == CRD_ACT_SIGMOID) return sigmoid_f(z);
  return z;
}
__device__ __forceinline__ float act_bwd(float z, int act) {
  if (act == CRD_ACT_GELU) return gelu_grad_f(z);
  if (act == CRD_ACT_SIGMOID) { float s = sigmoid_f(z); return s * (1.f - s); }
  return 1.f;
}

template <typename TI, typename TO>
__global__ void affine_act_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ ab,
                                  const float* __restrict__ post, int act, int B, long long N, int C, int ldx,
                                  int ldy) {
  const int cvec = C / 8;
  const long long total = (long long)B * N * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const long long pix = i / cvec;
    const int b = (int)(pix / N);
    const int c = cv * 8;
    float v[8];
    load8(x + pix * ldx + c, v);
    const float* abp = ab + ((long long)b * C + c) * 2;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float z = fmaf(abp[2 * j], v[j], abp[2 * j + 1]);
      z = act_fwd(z, act);
      if (post) z *= post[(long long)b * C + c + j];
      v[j] = z;
    }
    store8(y + pix * ldy + c, v);
  }
}

template <typename TD, typename TX>
__global__ void gnact_bwd_reduce_kernel(const TD* __restrict__ dy, const TX* __restrict__ x,
                                        const float* __restrict__ ab, const float* __restrict__ post,
                                        const float* __restrict__ addbc, int act, float* pq, int B, long long N,
                                        int C, int lddy, int ldx, long long ppb) {
  chan_reduce2([&](int b, long long p, int c, float (&s0)[8], float (&s1)[8]) {
    float g[8], v[8];
    const long long pix = (long long)b * N + p;
    load8(dy + pix * lddy + c, g);
    load8(x + pix * ldx + c, v);
    const float* abp = ab + ((long long)b * C + c) * 2;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float dz = g[j];
      if (addbc) dz += addbc[(long long)b * C + c + j];
      if (post) dz *= post[(long long)b * C + c + j];
      if (act != CRD_ACT_NONE) dz *= act_bwd(fmaf(abp[2 * j], v[j], abp[2 * j + 1]), act);
      s0[j] += dz;
      s1[j] = fmaf(dz, v[j], s1[j]);
    }
  }, pq, B, N, C, ppb);
}

__global__ void gn_bwd_finalize_kernel(const float* __restrict__ pq, const float* __restrict__ mean_rstd,
                                       const float* __restrict__ gamma, float* __restrict__ coef,
                                       float* dgamma, float* dbeta, int B, int C, int G, long long N) {
  const int b = blockIdx.x;
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mu = mean_rstd[((long long)b * G + g) * 2 + 0];
    const float r = mean_rstd[((long long)b * G + g) * 2 + 1];
    double t1 = 0.0, t2 = 0.0;
    for (int j = 0; j < cpg; j++) {
      const int cc = g * cpg + j;
      const float* p = pq + ((long long)b * C + cc) * 2;
      const double ga = (double)gamma[cc];
      t1 += ga * (double)p[0];
      t2 += ga * ((double)p[1] - (double)mu * (double)p[0]);
    }
    const double m = (double)cpg * (double)N;
    const double m1 = t1 / m;
    const double m2 = (double)r * t2 / m;
    float* cf = coef + ((long long)b * C + c) * 3;
    cf[0] = r * gamma[c];
    cf[1] = (float)(-(double)r * (double)r * m2);
    cf[2] = (float)(-(double)r * m1 + (double)r * (double)r * m2 * (double)mu);
    const float* p = pq + ((long long)b * C + c) * 2;
    if (dgamma) atomicAdd(dgamma + c, r * (p[1] - mu * p[0]));
    if (dbeta) atomicAdd(dbeta + c, p[0]);
  }
}

template <typename TD, typename TX, typename TO>
__global__ void gnact_bwd_apply_kernel(const TD* __restrict__ dy, const TX* __restrict__ x,
                                       const float* __restrict__ ab, const float* __restrict__ post,
                                       const float* __restrict__ addbc, int act, const float* __restrict__ coef,
                                       TO* __restrict__ dx, int accumulate, int B, long long N, int C, int lddy,
                                       int ldx, int lddx) {
  const int cvec = C / 8;
  const long long total = (long long)B * N * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const long long pix = i / cvec;
    const int b = (int)(pix / N);
    const int c = cv * 8;
    float g[8], v[8], o[8];
    load8(dy + pix * lddy + c, g);
    load8(x + pix * ldx + c, v);
    if (accumulate) load8(dx + pix * lddx + c, o);
    const float* abp = ab + ((long long)b * C + c) * 2;
    const float* cf = coef + ((long long)b * C + c) * 3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float dz = g[j];
      if (addbc) dz += addbc[(long long)b * C + c + j];
      if (post) dz *= post[(long long)b * C + c + j];
      if (act != CRD_ACT_NONE) dz *= act_bwd(fmaf(abp[2 * j], v[j], abp[2 * j + 1]), act);
      float r = fmaf(cf[3 * j], dz, fmaf(cf[3 * j + 1], v[j], cf[3 * j + 2]));
      o[j] = accumulate ? o[j] + r : r;
    }
    store8(dx + pix * lddx + c, o);
  }
}

}  // namespace

extern "C" int crd_chan_stats(const void* x, int dtype, float* sums, int B, long long N, int C, int ld,
                              crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256 && ld % 8 == 0);
  if (B == 0 || N == 0) return 0;
  ReduceLaunch r = plan_reduce(B, N, C);
  CRD_DISPATCH_1(dtype, T, chan_stats_kernel<T><<<r.grid, r.block, r.smem, (cudaStream_t)stream>>>(
                                (const T*)x, sums, B, N, C, ld, r.ppb));
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_gn_finalize(const float* sums, const float* gamma, const float* beta, float* ab,
                               float* mean_rstd, float* xbar, int B, int C, int G, long long N, float eps,
                               crd_stream_t stream) {
  CRD_REQUIRE(G > 0 && C % G == 0);
  if (B == 0) return 0;
  gn_finalize_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(sums, gamma, beta, ab, mean_rstd, xbar, B, C, G, N, eps);
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_affine_act(const void* x, int in_dtype, void* y, int out_dtype, const float* ab,
                              const float* post, int act, int B, long long N, int C, int ldx, int ldy,
                              crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0);
  const long long total = (long long)B * N * (C / 8);
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = ew_blocks(total);
  CRD_DISPATCH_1(in_dtype, TI, CRD_DISPATCH_1(out_dtype, TO, affine_act_kernel<TI, TO><<<nb, 256, 0, s>>>(
                                   (const TI*)x, (TO*)y, ab, post, act, B, N, C, ldx, ldy)));
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_gnact_bwd_reduce(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* ab,
                                    const float* post, const float* addbc, int act, float* pq, int B,
                                    long long N, int C, int lddy, int ldx, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256 && lddy % 8 == 0 && ldx % 8 == 0);
  if (B == 0 || N == 0) return 0;
  ReduceLaunch r = plan_reduce(B, N, C);
  cudaStream_t s = (cudaStream_t)stream;
  CRD_DISPATCH_1(dy_dtype, TD, CRD_DISPATCH_1(x_dtype, TX, gnact_bwd_reduce_kernel<TD, TX><<<r.grid, r.block, r.smem, s>>>(
                                   (const TD*)dy, (const TX*)x, ab, post, addbc, act, pq, B, N, C, lddy, ldx, r.ppb)));
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_gn_bwd_finalize(const float* pq, const float* mean_rstd, const float* gamma, float* coef,
                                   float* dgamma, float* dbeta, int B, int C, int G, long long N,
                                   crd_stream_t stream) {
  CRD_REQUIRE(G > 0 && C % G == 0);
  if (B == 0) return 0;
  gn_bwd_finalize_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(pq, mean_rstd, gamma, coef, dgamma, dbeta, B, C, G, N);
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_gnact_bwd_apply(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* ab,
                                   const float* post, const float* addbc, int act, const float* coef, void* dx,
                                   int dx_dtype, int accumulate, int B, long long N, int C, int lddy, int ldx,
                                   int lddx, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && lddy % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0);
  const long long total = (long long)B * N * (C / 8);
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = ew_blocks(total);
  CRD_DISPATCH_1(dy_dtype, TD, CRD_DISPATCH_1(x_dtype, TX, CRD_DISPATCH_1(dx_dtype, TO,
      gnact_bwd_apply_kernel<TD, TX, TO><<<nb, 256, 0, s>>>((const TD*)dy, (const TX*)x, ab, post, addbc, act, coef,
                                                           (TO*)dx, accumulate, B, N, C, lddy, ldx, lddx))));
  CRD_LAUNCH_CHECK();
  return 0;
}
