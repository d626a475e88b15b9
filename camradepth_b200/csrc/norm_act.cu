// GroupNorm as a three-step protocol: per-(b,c) sums -> finalize into a per-(b,c)
// affine -> apply (optionally fused with GELU / Dropout2d scale) in one streaming pass; and the
// matching two-pass backward.  All kernels are HBM-bound: 16-byte vector accesses, channel-major
// thread mapping (coalesced NHWC rows), per-thread register accumulation, one atomic per
// (block, channel).
#include "common.cuh"
#include "chan_reduce.cuh"
#include "gn_stream_tma.cuh"
#include "../../include/camradepth_b200.h"

namespace {

template <typename T>
__global__ void chan_stats_kernel(const T* __restrict__ x, float* sums, int B, long long N, int C, int ld,
                                  long long ppb) {
  CRD_PDL_ENTRY();
  const int c = threadIdx.x * 8, ry = threadIdx.y, rows = blockDim.y, b = blockIdx.y;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  const T* xb = x + (long long)b * N * ld + c;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { s0[j] = 0.f; s1[j] = 0.f; }
  constexpr int U = 8;                         // eight independent 16-byte loads in flight per thread
  for (long long p = p0 + ry; p < p1; p += (long long)U * rows) {
    typename Raw8<T>::type raw[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * rows;
      raw[u] = ldg16(xb + (q < p1 ? q : p) * ld);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (p + (long long)u * rows >= p1) break;
      float v[8];
      unpack8(raw[u], v);
#pragma unroll
      for (int j = 0; j < 8; j++) { s0[j] += v[j]; s1[j] = fmaf(v[j], v[j], s1[j]); }
    }
  }
  chan_reduce_finish(s0, s1, sums, b, C);
}

__global__ void gn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ ab,
                                   float* __restrict__ mean_rstd, float* __restrict__ xbar, int B, int C, int G,
                                   long long N, float eps) {
  CRD_PDL_ENTRY();
  extern __shared__ float gs[];             // [G][2]: mean, rstd
  const int b = blockIdx.x;
  const int cpg = C / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // blockIdx.y = slice of whole groups (gridDim.y slices): wide layers (512 / 1024 channels) no longer walk 4-8 groups
  // per warp one after the other
  const int gps = (G + gridDim.y - 1) / gridDim.y;
  const int g0 = blockIdx.y * gps, g1 = min(G, g0 + gps);
  for (int g = g0 + warp; g < g1; g += nw) {      // one warp per group
    double s = 0.0, ss = 0.0;
    for (int j = lane; j < cpg; j += 32) {
      const float* p = sums + ((long long)b * C + g * cpg + j) * 2;
      s += (double)p[0]; ss += (double)p[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) {
      const double cnt = (double)cpg * (double)N;
      const double mean = s / cnt;
      double var = ss / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      gs[2 * g] = (float)mean; gs[2 * g + 1] = rstd;
      mean_rstd[((long long)b * G + g) * 2 + 0] = (float)mean;
      mean_rstd[((long long)b * G + g) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  for (int c = g0 * cpg + threadIdx.x; c < g1 * cpg; c += blockDim.x) {
    const int g = c / cpg;
    const float a = gamma[c] * gs[2 * g + 1];
    const float bb = beta[c] - gs[2 * g] * a;
    ab[((long long)b * C + c) * 2 + 0] = a;
    ab[((long long)b * C + c) * 2 + 1] = bb;
    if (xbar) xbar[(long long)b * C + c] = a * (sums[((long long)b * C + c) * 2] / (float)N) + bb;
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
                                  int ldy, long long ppb) {
  CRD_PDL_ENTRY();
  // blockDim = (C/8, rows); each thread keeps the per-(b,c) affine of its 8 channels in registers and
  // streams over the block's pixel range (coalesced 16-byte accesses, no per-pixel parameter loads)
  const int c = threadIdx.x * 8, ry = threadIdx.y, rows = blockDim.y;
  const int b = blockIdx.y;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  float a[8], sh[8], ps[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = ab[((long long)b * C + c + j) * 2];
    sh[j] = ab[((long long)b * C + c + j) * 2 + 1];
    ps[j] = post ? post[(long long)b * C + c + j] : 1.f;
  }
  const TI* xb = x + (long long)b * N * ldx + c;
  TO* yb = y + (long long)b * N * ldy + c;
  // four independent 16-byte loads in flight per thread before any is consumed
  for (long long p = p0 + ry; p < p1; p += 4LL * rows) {
    typename Raw8<TI>::type raw[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long q = p + (long long)u * rows;
      raw[u] = ldg16(xb + (q < p1 ? q : p) * ldx);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long q = p + (long long)u * rows;
      if (q >= p1) break;
      float v[8];
      unpack8(raw[u], v);
      if (act == CRD_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = gelu_f(fmaf(a[j], v[j], sh[j])) * ps[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = act_fwd(fmaf(a[j], v[j], sh[j]), act) * ps[j];
      }
      store8(yb + q * ldy, v);
    }
  }
}

template <typename TD, typename TX>
__global__ void gnact_bwd_reduce_kernel(const TD* __restrict__ dy, const TX* __restrict__ x,
                                        const float* __restrict__ ab, const float* __restrict__ post,
                                        const float* __restrict__ addbc, int act, float* pq, TD* dz_out, int B,
                                        long long N, int C, int lddy, int ldx, long long ppb) {
  CRD_PDL_ENTRY();
  const int c = threadIdx.x * 8, ry = threadIdx.y, rows = blockDim.y, b = blockIdx.y;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  float a[8], sh[8], ps[8], ad[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = ab[((long long)b * C + c + j) * 2];
    sh[j] = ab[((long long)b * C + c + j) * 2 + 1];
    ps[j] = post ? post[(long long)b * C + c + j] : 1.f;
    ad[j] = addbc ? addbc[(long long)b * C + c + j] : 0.f;
  }
  const TD* dyb = dy + (long long)b * N * lddy + c;
  const TX* xb = x + (long long)b * N * ldx + c;
  TD* dzb = dz_out ? dz_out + (long long)b * N * lddy + c : nullptr;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { s0[j] = 0.f; s1[j] = 0.f; }
  constexpr int U = 4;                         // four (dy, x) pairs in flight per thread
  for (long long p = p0 + ry; p < p1; p += (long long)U * rows) {
    typename Raw8<TD>::type rg[U];
    typename Raw8<TX>::type rx[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * rows;
      const long long qq = q < p1 ? q : p;
      rg[u] = ldg16(dyb + qq * lddy);
      rx[u] = ldg16(xb + qq * ldx);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * rows;
      if (q >= p1) break;
      float g[8], v[8];
      unpack8(rg[u], g);
      unpack8(rx[u], v);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float dz = (g[j] + ad[j]) * ps[j];
        if (act == CRD_ACT_GELU) dz *= gelu_grad_f(fmaf(a[j], v[j], sh[j]));
        else if (act != CRD_ACT_NONE) dz *= act_bwd(fmaf(a[j], v[j], sh[j]), act);
        g[j] = dz;
        s0[j] += dz;
        s1[j] = fmaf(dz, v[j], s1[j]);
      }
      // optionally materialise dz (may alias dy): the apply pass then skips the activation derivative
      if (dzb) store8(dzb + q * lddy, g);
    }
  }
  chan_reduce_finish(s0, s1, pq, b, C);
}

__global__ void gn_bwd_finalize_kernel(const float* __restrict__ pq, const float* __restrict__ mean_rstd,
                                       const float* __restrict__ gamma, float* __restrict__ coef,
                                       float* dgamma, float* dbeta, int B, int C, int G, long long N) {
  CRD_PDL_ENTRY();
  extern __shared__ float gs[];             // [G][2]: m1, m2
  const int b = blockIdx.x;
  const int cpg = C / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int gps = (G + gridDim.y - 1) / gridDim.y;          // blockIdx.y = slice of whole groups
  const int g0 = blockIdx.y * gps, g1 = min(G, g0 + gps);
  for (int g = g0 + warp; g < g1; g += nw) {
    const float mu = mean_rstd[((long long)b * G + g) * 2 + 0];
    const float r = mean_rstd[((long long)b * G + g) * 2 + 1];
    double t1 = 0.0, t2 = 0.0;
    for (int j = lane; j < cpg; j += 32) {
      const int cc = g * cpg + j;
      const float* p = pq + ((long long)b * C + cc) * 2;
      const double ga = (double)gamma[cc];
      t1 += ga * (double)p[0];
      t2 += ga * ((double)p[1] - (double)mu * (double)p[0]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (lane == 0) {
      const double m = (double)cpg * (double)N;
      gs[2 * g] = (float)(t1 / m);
      gs[2 * g + 1] = (float)((double)r * t2 / m);
    }
  }
  __syncthreads();
  for (int c = g0 * cpg + threadIdx.x; c < g1 * cpg; c += blockDim.x) {
    const int g = c / cpg;
    const float mu = mean_rstd[((long long)b * G + g) * 2 + 0];
    const float r = mean_rstd[((long long)b * G + g) * 2 + 1];
    const float m1 = gs[2 * g], m2 = gs[2 * g + 1];
    float* cf = coef + ((long long)b * C + c) * 3;
    cf[0] = r * gamma[c];
    cf[1] = -r * r * m2;
    cf[2] = -r * m1 + r * r * m2 * mu;
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
                                       int ldx, int lddx, long long ppb) {
  CRD_PDL_ENTRY();
  const int c = threadIdx.x * 8, ry = threadIdx.y, rows = blockDim.y;
  const int b = blockIdx.y;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  float a[8], sh[8], ps[8], ad[8], cA[8], cB[8], cC[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const long long bc = (long long)b * C + c + j;
    a[j] = ab[bc * 2];
    sh[j] = ab[bc * 2 + 1];
    ps[j] = post ? post[bc] : 1.f;
    ad[j] = addbc ? addbc[bc] : 0.f;
    cA[j] = coef[bc * 3];
    cB[j] = coef[bc * 3 + 1];
    cC[j] = coef[bc * 3 + 2];
  }
  const TD* dyb = dy + (long long)b * N * lddy + c;
  const TX* xb = x + (long long)b * N * ldx + c;
  TO* dxb = dx + (long long)b * N * lddx + c;
  constexpr int U = 4;
  for (long long p = p0 + ry; p < p1; p += (long long)U * rows) {
    typename Raw8<TD>::type rg[U];
    typename Raw8<TX>::type rx[U];
    typename Raw8<TO>::type ro[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * rows;
      const long long qq = q < p1 ? q : p;
      rg[u] = ldg16(dyb + qq * lddy);
      rx[u] = ldg16(xb + qq * ldx);
      if (accumulate) ro[u] = ldg16(dxb + qq * lddx);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * rows;
      if (q >= p1) break;
      float g[8], v[8], o[8];
      unpack8(rg[u], g);
      unpack8(rx[u], v);
      if (accumulate) unpack8(ro[u], o);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float dz = (g[j] + ad[j]) * ps[j];
        if (act == CRD_ACT_GELU) dz *= gelu_grad_f(fmaf(a[j], v[j], sh[j]));
        else if (act != CRD_ACT_NONE) dz *= act_bwd(fmaf(a[j], v[j], sh[j]), act);
        const float r = fmaf(cA[j], dz, fmaf(cB[j], v[j], cC[j]));
        o[j] = accumulate ? o[j] + r : r;
      }
      store8(dxb + q * lddx, o);
    }
  }
}

}  // namespace

extern "C" int crd_chan_stats(const void* x, int dtype, float* sums, int B, long long N, int C, int ld,
                              crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256 && ld % 8 == 0);
  if (B == 0 || N == 0) return 0;
  if (dtype == CRD_BF16 && st_eligible(B, N, C, x, ld, nullptr, 0)) {
    StLaunch L = st_plan(B, N, C);
    L.p.red = sums;
    CUtensorMap m0;
    if (int e = st_map(&m0, x, B, N, C, ld, L.p)) return e;
    if (int e = st_launch<ST_STATS>(L, m0, m0, (cudaStream_t)stream)) return e;
    CRD_LAUNCH_CHECK();
    return 0;
  }
  ReduceLaunch r = plan_reduce(B, N, C);
  CRD_DISPATCH_1(dtype, T, crd_launch(chan_stats_kernel<T>, dim3(r.grid), dim3(r.block), r.smem, (cudaStream_t)stream, 
                                (const T*)x, sums, B, N, C, ld, r.ppb));
  CRD_LAUNCH_CHECK();
  return 0;
}

// slices of whole groups per sample for the finalize kernels: one per 256 channels and at most one per 8 groups
static int gn_fin_slices(int C, int G) {
  int s = C / 256;
  if (s > G / 8) s = G / 8;
  return s < 1 ? 1 : s;
}
extern "C" int crd_gn_finalize(const float* sums, const float* gamma, const float* beta, float* ab,
                               float* mean_rstd, float* xbar, int B, int C, int G, long long N, float eps,
                               crd_stream_t stream) {
  CRD_REQUIRE(G > 0 && C % G == 0);
  if (B == 0) return 0;
  crd_launch(gn_finalize_kernel, dim3(B, gn_fin_slices(C, G)), dim3(256), 2 * G * sizeof(float), (cudaStream_t)stream, sums, gamma, beta, ab, mean_rstd, xbar, B, C, G, N, eps);
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_affine_act(const void* x, int in_dtype, void* y, int out_dtype, const float* ab,
                              const float* post, int act, int B, long long N, int C, int ldx, int ldy,
                              crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0);
  CRD_REQUIRE(C / 8 <= 256);
  if ((long long)B * N == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (in_dtype == CRD_BF16 && out_dtype == CRD_BF16 && st_eligible(B, N, C, x, ldx, y, ldy)) {
    StLaunch L = st_plan(B, N, C);
    L.p.ab = ab; L.p.post = post; L.p.act = act; L.p.out = (bf16*)y; L.p.ldo = ldy;
    CUtensorMap m0;
    if (int e = st_map(&m0, x, B, N, C, ldx, L.p)) return e;
    if (int e = st_launch<ST_AFFINE>(L, m0, m0, s)) return e;
    CRD_LAUNCH_CHECK();
    return 0;
  }
  ReduceLaunch r = plan_stream(B, N, C);
  CRD_DISPATCH_1(in_dtype, TI, CRD_DISPATCH_1(out_dtype, TO, crd_launch(affine_act_kernel<TI, TO>, dim3(r.grid), dim3(r.block), 0, s, 
                                   (const TI*)x, (TO*)y, ab, post, act, B, N, C, ldx, ldy, r.ppb)));
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_gnact_bwd_reduce(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* ab,
                                    const float* post, const float* addbc, int act, float* pq, void* dz_out,
                                    int B, long long N, int C, int lddy, int ldx, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256 && lddy % 8 == 0 && ldx % 8 == 0);
  if (B == 0 || N == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dy_dtype == CRD_BF16 && x_dtype == CRD_BF16 && st_eligible(B, N, C, dy, lddy, x, ldx)) {
    StLaunch L = st_plan(B, N, C);
    L.p.ab = ab; L.p.post = post; L.p.addbc = addbc; L.p.act = act; L.p.red = pq;
    L.p.out = (bf16*)dz_out; L.p.ldo = lddy;
    CUtensorMap m0, m1;
    if (int e = st_map(&m0, dy, B, N, C, lddy, L.p)) return e;
    if (int e = st_map(&m1, x, B, N, C, ldx, L.p)) return e;
    if (int e = st_launch<ST_BWD_REDUCE>(L, m0, m1, s)) return e;
    CRD_LAUNCH_CHECK();
    return 0;
  }
  ReduceLaunch r = plan_reduce(B, N, C);
  CRD_DISPATCH_1(dy_dtype, TD, CRD_DISPATCH_1(x_dtype, TX, crd_launch(gnact_bwd_reduce_kernel<TD, TX>, dim3(r.grid), dim3(r.block), r.smem, s, 
                                   (const TD*)dy, (const TX*)x, ab, post, addbc, act, pq, (TD*)dz_out, B, N, C, lddy, ldx, r.ppb)));
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_gn_bwd_finalize(const float* pq, const float* mean_rstd, const float* gamma, float* coef,
                                   float* dgamma, float* dbeta, int B, int C, int G, long long N,
                                   crd_stream_t stream) {
  CRD_REQUIRE(G > 0 && C % G == 0);
  if (B == 0) return 0;
  crd_launch(gn_bwd_finalize_kernel, dim3(B, gn_fin_slices(C, G)), dim3(256), 2 * G * sizeof(float), (cudaStream_t)stream, pq, mean_rstd, gamma, coef, dgamma, dbeta, B, C, G, N);
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_gnact_bwd_apply(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* ab,
                                   const float* post, const float* addbc, int act, const float* coef, void* dx,
                                   int dx_dtype, int accumulate, int B, long long N, int C, int lddy, int ldx,
                                   int lddx, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && lddy % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0);
  CRD_REQUIRE(C / 8 <= 256);
  if ((long long)B * N == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dy_dtype == CRD_BF16 && x_dtype == CRD_BF16 && dx_dtype == CRD_BF16 && !accumulate &&
      st_eligible(B, N, C, dy, lddy, x, ldx) && ((uintptr_t)dx & 15) == 0) {
    StLaunch L = st_plan(B, N, C);
    L.p.ab = ab; L.p.post = post; L.p.addbc = addbc; L.p.act = act; L.p.coef = coef;
    L.p.out = (bf16*)dx; L.p.ldo = lddx;
    CUtensorMap m0, m1;
    if (int e = st_map(&m0, dy, B, N, C, lddy, L.p)) return e;
    if (int e = st_map(&m1, x, B, N, C, ldx, L.p)) return e;
    if (int e = st_launch<ST_BWD_APPLY>(L, m0, m1, s)) return e;
    CRD_LAUNCH_CHECK();
    return 0;
  }
  ReduceLaunch r = plan_stream(B, N, C);
  CRD_DISPATCH_1(dy_dtype, TD, CRD_DISPATCH_1(x_dtype, TX, CRD_DISPATCH_1(dx_dtype, TO,
      crd_launch(gnact_bwd_apply_kernel<TD, TX, TO>, dim3(r.grid), dim3(r.block), 0, s, (const TD*)dy, (const TX*)x, ab, post, addbc, act,
                                                                   coef, (TO*)dx, accumulate, B, N, C, lddy, ldx,
                                                                   lddx, r.ppb))));
  CRD_LAUNCH_CHECK();
  return 0;
}
