// GroupNorm (+GELU, +Dropout2d scale) forward and backward in ONE launch each, for the encoder-sized tensors.
//
// The three-step protocol of norm_act.cu (per-(b,c) sums -> finalize -> apply; backward: reduce -> finalize ->
// apply) costs three launches per GroupNorm and direction.  The encoder has 165 GroupNorms per step on tensors of
// 0.1 .. 80 MB whose kernels are a few microseconds each, so the step pays ~1000 launches of fixed cost for them
// (measured: gn_finalize 4.2 us, gn_bwd_finalize 5.0 us per launch for a few hundred bytes of work).  Here one CTA
// owns a (sample, channel tile of whole groups) slab [N pixels][CT channels] and sweeps it twice: sweep 1 reduces
// (the second read of the slab is an L2 hit: these tensors fit the 126 MB L2), the statistics are finalised inside
// the CTA (shared memory, fp64 for the mean / variance), sweep 2 applies.  Same arithmetic as the three-step path.
// Slabs with many pixels are cut over a THREAD-BLOCK CLUSTER of up to 8 CTAs (one pixel range each): the partial
// sums meet through distributed shared memory in rank order (no atomics: the result is bit-reproducible), every
// CTA finalises redundantly and applies its own range -- 8x the CTAs for the 5k-token maps of stage 1.
//   forward : y = act(a*x + b) * post,   a = gamma*rstd, b = beta - mean*a        (utils.py:223-228,
//             simplified_attention.py:36-38,141-145,184-187); optionally the sums come from the producing conv's
//             read-out (sums_in) and sweep 1 is skipped; ab / mean_rstd / xbar are written for the backward pass
//   backward: dz = (dy + addbc) * post * act'(a*x+b);  dx (+)= A*dz + Bq*x + Cq;  dgamma/dbeta += ...
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/camradepth_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int GF_THREADS = 256;

__device__ __forceinline__ float gf_act_fwd(float z, int act) {
  if (act == CRD_ACT_GELU) return gelu_f(z);
  if (act == CRD_ACT_SIGMOID) return sigmoid_f(z);
  return z;
}
__device__ __forceinline__ float gf_act_bwd(float z, int act) {
  if (act == CRD_ACT_GELU) return gelu_grad_f(z);
  if (act == CRD_ACT_SIGMOID) { const float s = sigmoid_f(z); return s * (1.f - s); }
  return 1.f;
}

struct GfGeom {
  int B, C, G, cpg, CT, cvec, rows;
  int cs;                   // cluster size (CTAs per slab)
  long long N, per;         // pixels per sample, pixels per CTA of a cluster
};

// totals of all CTAs of the cluster, summed in rank order (deterministic); all[] is this CTA's copy
__device__ __forceinline__ void gf_cluster_reduce(cg::cluster_group& cluster, float* tot, float* all, int n) {
  if (cluster.num_blocks() == 1) {
    for (int i = threadIdx.x; i < n; i += GF_THREADS) all[i] = tot[i];
    __syncthreads();
    return;
  }
  cluster.sync();                                   // every CTA's tot[] is complete and visible
  for (int i = threadIdx.x; i < n; i += GF_THREADS) {
    float a = 0.f;
    for (unsigned r = 0; r < cluster.num_blocks(); r++) a += cluster.map_shared_rank(tot, r)[i];
    all[i] = a;
  }
  cluster.sync();                                   // nobody leaves (or reuses tot) while a peer still reads it
}

// per-thread partial sums -> per-channel totals in shared memory: red[0][c], red[1][c]
__device__ __forceinline__ void gf_block_reduce(const float (&s0)[8], const float (&s1)[8], float* part /*[rows][CT][2]*/,
                                                float* tot /*[2][CT]*/, int tx, int ty, int CT, int rows) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
    part[(ty * CT + tx * 8 + j) * 2 + 0] = s0[j];
    part[(ty * CT + tx * 8 + j) * 2 + 1] = s1[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * CT; c += GF_THREADS) {
    const int ch = c >> 1, q = c & 1;
    float a = 0.f;
    for (int r = 0; r < rows; r++) a += part[(r * CT + ch) * 2 + q];
    tot[q * CT + ch] = a;
  }
  __syncthreads();
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(GF_THREADS)
gn_fused_fwd_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ sums_in, const float* __restrict__ post,
                    int act, float* __restrict__ ab_out, float* __restrict__ mr_out, float* __restrict__ xbar_out,
                    GfGeom g, int ldx, int ldy, float eps) {
  extern __shared__ float sm[];
  CRD_PDL_ENTRY();
  cg::cluster_group cluster = cg::this_cluster();
  float* part = sm;                                   // [rows][CT][2]
  float* tot = part + g.rows * g.CT * 2;              // [2][CT] this CTA's pixel range
  float* all = tot + 2 * g.CT;                        // [2][CT] whole slab
  float* coef = all + 2 * g.CT;                       // [2][CT]: a, b
  const int tx = threadIdx.x % g.cvec, ty = threadIdx.x / g.cvec;
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y, c0 = (blockIdx.x / g.cs) * g.CT, c = c0 + tx * 8;
  const long long pbeg = (long long)rank * g.per, pend = min(g.N, pbeg + g.per);
  const TI* xb = x + (long long)b * g.N * ldx + c;
  if (sums_in == nullptr) {
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { s0[j] = 0.f; s1[j] = 0.f; }
    constexpr int U = 4;
    for (long long p = pbeg + ty; p < pend; p += (long long)U * g.rows) {
      typename Raw8<TI>::type raw[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long q = p + (long long)u * g.rows;
        raw[u] = ldg16(xb + (q < pend ? q : p) * ldx);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (p + (long long)u * g.rows >= pend) break;
        float v[8];
        unpack8(raw[u], v);
#pragma unroll
        for (int j = 0; j < 8; j++) { s0[j] += v[j]; s1[j] = fmaf(v[j], v[j], s1[j]); }
      }
    }
    gf_block_reduce(s0, s1, part, tot, tx, ty, g.CT, g.rows);
    gf_cluster_reduce(cluster, tot, all, 2 * g.CT);
  } else {
    for (int i = threadIdx.x; i < 2 * g.CT; i += GF_THREADS) {
      const int ch = i >> 1, q = i & 1;
      all[q * g.CT + ch] = sums_in[((long long)b * g.C + c0 + ch) * 2 + q];
    }
    __syncthreads();
  }
  tot = all;
  const bool writer = rank == 0;                      // one CTA of the cluster publishes ab / mean_rstd / xbar
  // finalize: one thread per group of this tile
  const int gpt = g.CT / g.cpg;
  if (threadIdx.x < gpt) {
    double s = 0.0, ss = 0.0;
    for (int j = 0; j < g.cpg; j++) { s += (double)tot[threadIdx.x * g.cpg + j]; ss += (double)tot[g.CT + threadIdx.x * g.cpg + j]; }
    const double cnt = (double)g.cpg * (double)g.N;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const int gi = c0 / g.cpg + threadIdx.x;
    if (mr_out && writer) { mr_out[((long long)b * g.G + gi) * 2] = (float)mean; mr_out[((long long)b * g.G + gi) * 2 + 1] = rstd; }
    for (int j = 0; j < g.cpg; j++) {
      const int ch = threadIdx.x * g.cpg + j;
      const float a = gamma[c0 + ch] * rstd;
      const float bb = beta[c0 + ch] - (float)mean * a;
      coef[ch] = a; coef[g.CT + ch] = bb;
      if (ab_out && writer) { ab_out[((long long)b * g.C + c0 + ch) * 2] = a; ab_out[((long long)b * g.C + c0 + ch) * 2 + 1] = bb; }
      if (xbar_out && writer) xbar_out[(long long)b * g.C + c0 + ch] = a * (tot[ch] / (float)g.N) + bb;
    }
  }
  __syncthreads();
  if (y == nullptr) return;
  float a[8], sh[8], ps[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = coef[tx * 8 + j];
    sh[j] = coef[g.CT + tx * 8 + j];
    ps[j] = post ? post[(long long)b * g.C + c + j] : 1.f;
  }
  TO* yb = y + (long long)b * g.N * ldy + c;
  constexpr int U = 4;
  for (long long p = pbeg + ty; p < pend; p += (long long)U * g.rows) {
    typename Raw8<TI>::type raw[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * g.rows;
      raw[u] = ldg16(xb + (q < pend ? q : p) * ldx);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * g.rows;
      if (q >= pend) break;
      float v[8];
      unpack8(raw[u], v);
      if (act == CRD_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = gelu_f(fmaf(a[j], v[j], sh[j])) * ps[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = gf_act_fwd(fmaf(a[j], v[j], sh[j]), act) * ps[j];
      }
      store8(yb + q * ldy, v);
    }
  }
}

template <typename TD, typename TX, typename TO>
__global__ void __launch_bounds__(GF_THREADS, 2)
gn_fused_bwd_kernel(TD* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ ab,
                    const float* __restrict__ mean_rstd, const float* __restrict__ gamma, const float* __restrict__ post,
                    const float* __restrict__ addbc, int act, TO* __restrict__ dx, int accumulate, float* dgamma,
                    float* dbeta, GfGeom g, int lddy, int ldx, int lddx) {
  extern __shared__ float sm[];
  CRD_PDL_ENTRY();
  cg::cluster_group cluster = cg::this_cluster();
  float* part = sm;
  float* tot = part + g.rows * g.CT * 2;              // [2][CT]: sum dz, sum dz*x over this CTA's pixel range
  float* all = tot + 2 * g.CT;                        // [2][CT]: over the whole slab
  float* coef = all + 2 * g.CT;                       // [3][CT]: A, Bq, Cq
  const int tx = threadIdx.x % g.cvec, ty = threadIdx.x / g.cvec;
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y, c0 = (blockIdx.x / g.cs) * g.CT, c = c0 + tx * 8;
  const long long pbeg = (long long)rank * g.per, pend = min(g.N, pbeg + g.per);
  float a[8], sh[8], k1[8], k0[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const long long bc = (long long)b * g.C + c + j;
    a[j] = ab[bc * 2];
    sh[j] = ab[bc * 2 + 1];
    const float ps = post ? post[bc] : 1.f, ad = addbc ? addbc[bc] : 0.f;
    k1[j] = ps; k0[j] = ad * ps;                      // dz = (k1*dy + k0) * act'(z)
  }
  TD* dyb = dy + (long long)b * g.N * lddy + c;
  const TX* xb = x + (long long)b * g.N * ldx + c;
  const bool inplace = act != CRD_ACT_NONE;           // dz replaces dy, so act' is evaluated once (dy is consumed here)
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { s0[j] = 0.f; s1[j] = 0.f; }
  constexpr int U = 2;                     // (two CTAs of 128 registers per SM; four loads per array in flight spilled)
  for (long long p = pbeg + ty; p < pend; p += (long long)U * g.rows) {
    typename Raw8<TD>::type rg[U];
    typename Raw8<TX>::type rx[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * g.rows;
      const long long qq = q < pend ? q : p;
      rg[u] = ldg16(dyb + qq * lddy);
      rx[u] = ldg16(xb + qq * ldx);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * g.rows;
      if (q >= pend) break;
      float gd[8], v[8];
      unpack8(rg[u], gd);
      unpack8(rx[u], v);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float dz = fmaf(k1[j], gd[j], k0[j]);
        if (act == CRD_ACT_GELU) dz *= gelu_grad_f(fmaf(a[j], v[j], sh[j]));
        else if (act != CRD_ACT_NONE) dz *= gf_act_bwd(fmaf(a[j], v[j], sh[j]), act);
        gd[j] = dz;
        s0[j] += dz;
        s1[j] = fmaf(dz, v[j], s1[j]);
      }
      if (inplace) store8(dyb + q * lddy, gd);
    }
  }
  gf_block_reduce(s0, s1, part, tot, tx, ty, g.CT, g.rows);
  gf_cluster_reduce(cluster, tot, all, 2 * g.CT);
  tot = all;
  const bool writer = rank == 0;
  const int gpt = g.CT / g.cpg;
  if (threadIdx.x < gpt) {
    const int gi = c0 / g.cpg + threadIdx.x;
    const float mu = mean_rstd[((long long)b * g.G + gi) * 2], r = mean_rstd[((long long)b * g.G + gi) * 2 + 1];
    double t1 = 0.0, t2 = 0.0;
    for (int j = 0; j < g.cpg; j++) {
      const int ch = threadIdx.x * g.cpg + j;
      const double ga = (double)gamma[c0 + ch];
      t1 += ga * (double)tot[ch];
      t2 += ga * ((double)tot[g.CT + ch] - (double)mu * (double)tot[ch]);
    }
    const double m = (double)g.cpg * (double)g.N;
    const float m1 = (float)(t1 / m), m2 = (float)((double)r * t2 / m);
    for (int j = 0; j < g.cpg; j++) {
      const int ch = threadIdx.x * g.cpg + j;
      coef[ch] = r * gamma[c0 + ch];
      coef[g.CT + ch] = -r * r * m2;
      coef[2 * g.CT + ch] = -r * m1 + r * r * m2 * mu;
      if (dgamma && writer) atomicAdd(dgamma + c0 + ch, r * (tot[g.CT + ch] - mu * tot[ch]));
      if (dbeta && writer) atomicAdd(dbeta + c0 + ch, tot[ch]);
    }
  }
  __syncthreads();
  float cA[8], cB[8], cC[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    cA[j] = coef[tx * 8 + j];
    cB[j] = coef[g.CT + tx * 8 + j];
    cC[j] = coef[2 * g.CT + tx * 8 + j];
  }
  if (inplace) __threadfence_block();                 // this thread re-reads only what it wrote itself
  TO* dxb = dx + (long long)b * g.N * lddx + c;
  for (long long p = pbeg + ty; p < pend; p += (long long)U * g.rows) {
    typename Raw8<TD>::type rg[U];
    typename Raw8<TX>::type rx[U];
    typename Raw8<TO>::type ro[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * g.rows;
      const long long qq = q < pend ? q : p;
      rg[u] = ldg16(dyb + qq * lddy);
      rx[u] = ldg16(xb + qq * ldx);
      if (accumulate) ro[u] = ldg16(dxb + qq * lddx);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long q = p + (long long)u * g.rows;
      if (q >= pend) break;
      float gd[8], v[8], o[8];
      unpack8(rg[u], gd);
      unpack8(rx[u], v);
      if (accumulate) unpack8(ro[u], o);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float dz = inplace ? gd[j] : fmaf(k1[j], gd[j], k0[j]);
        const float rr = fmaf(cA[j], dz, fmaf(cB[j], v[j], cC[j]));
        o[j] = accumulate ? o[j] + rr : rr;
      }
      store8(dxb + q * lddx, o);
    }
  }
}

// channel tile: whole groups, at least 32 channels (64 .. 128-byte rows) when the channel count allows
inline bool gf_geom(int B, long long N, int C, int G, GfGeom& g) {
  if (G <= 0 || C % G) return false;
  const int cpg = C / G;
  if (cpg % 8 || cpg > 256) return false;
  int CT = cpg;
  while (CT < 32 && C % (CT * 2) == 0) CT *= 2;
  if (CT / 8 > GF_THREADS) return false;
  g.B = B; g.C = C; g.G = G; g.cpg = cpg; g.CT = CT; g.cvec = CT / 8; g.rows = GF_THREADS / g.cvec; g.N = N;
  if (GF_THREADS % g.cvec) return false;
  // cluster size: a function of the slab geometry ONLY (not of the batch size), so that a sample's statistics are
  // reduced in the same order whatever batch it sits in; every CTA keeps >= 16 pixels per thread row
  int cs = 1;
  while (cs < 8 && N / (cs * 2) >= (long long)g.rows * 16) cs *= 2;
  g.cs = cs;
  g.per = (N + cs - 1) / cs;
  return true;
}
inline size_t gf_smem(const GfGeom& g, int ncoef) { return (size_t)(g.rows * g.CT * 2 + 4 * g.CT + ncoef * g.CT) * sizeof(float); }

template <typename K, typename... Args>
inline cudaError_t gf_launch(K kernel, const GfGeom& g, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((g.C / g.CT) * g.cs), (unsigned)g.B, 1);
  cfg.blockDim = dim3(GF_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)g.cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = crd_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace

extern "C" int crd_gn_fused_supported(int B, long long N, int C, int G) {
  GfGeom g;
  return gf_geom(B, N, C, G, g) ? 1 : 0;
}

extern "C" int crd_gn_fused_fwd(const void* x, int x_dtype, void* y, int y_dtype, const float* gamma, const float* beta,
                                const float* sums_in, const float* post, int act, float* ab_out, float* mean_rstd_out,
                                float* xbar_out, int B, long long N, int C, int G, int ldx, int ldy, float eps,
                                crd_stream_t stream) {
  GfGeom g;
  CRD_REQUIRE(gf_geom(B, N, C, G, g));
  CRD_REQUIRE(ldx % 8 == 0 && (y == nullptr || ldy % 8 == 0) && gamma && beta);
  CRD_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0);
  if ((long long)B * N == 0) return 0;
  const size_t smem = gf_smem(g, 2);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t err = cudaSuccess;
  CRD_DISPATCH_1(x_dtype, TI, CRD_DISPATCH_1(y_dtype, TO, err = gf_launch(
                                  gn_fused_fwd_kernel<TI, TO>, g, smem, s, (const TI*)x, (TO*)y, gamma, beta, sums_in, post,
                                  act, ab_out, mean_rstd_out, xbar_out, g, ldx, ldy, eps)));
  if (err != cudaSuccess) return (int)err;
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_gn_fused_bwd(void* dy, int dy_dtype, const void* x, int x_dtype, const float* ab,
                                const float* mean_rstd, const float* gamma, const float* post, const float* addbc,
                                int act, void* dx, int dx_dtype, int accumulate, float* dgamma, float* dbeta, int B,
                                long long N, int C, int G, int lddy, int ldx, int lddx, crd_stream_t stream) {
  GfGeom g;
  CRD_REQUIRE(gf_geom(B, N, C, G, g));
  CRD_REQUIRE(lddy % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0 && ab && mean_rstd && gamma && dx);
  CRD_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0);
  if ((long long)B * N == 0) return 0;
  const size_t smem = gf_smem(g, 3);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t err = cudaSuccess;
  CRD_DISPATCH_1(dy_dtype, TD, CRD_DISPATCH_1(x_dtype, TX, CRD_DISPATCH_1(dx_dtype, TO, err = gf_launch(
      gn_fused_bwd_kernel<TD, TX, TO>, g, smem, s, (TD*)dy, (const TX*)x, ab, mean_rstd, gamma, post, addbc, act, (TO*)dx,
      accumulate, dgamma, dbeta, g, lddy, ldx, lddx))));
  if (err != cudaSuccess) return (int)err;
  CRD_LAUNCH_CHECK();
  return 0;
}
