// Encoder block pieces that are not plain contractions:
// depthwise 3x3 fused with the preceding GroupNorm apply, the max-pool "attention" score
// (QK^T -> max over keys -> sum over heads, never materialising N x M), the rank-1 attention
// output fused with the residual/DropPath add, and their backward passes.
#include "common.cuh"
#include "chan_reduce.cuh"
#include "dwconv_tma.cuh"
#include "../../include/camradepth_b200.h"

namespace {

// ------------------------------------------------------------------ depthwise 3x3
// blockDim = (C/8, rows), grid = (pixel chunks, B): each thread keeps the 9x8 filter taps, the GroupNorm affine
// and the bias of its 8 channels in registers and streams over pixels (16-byte loads, neighbours hit L1/L2).
template <typename T>
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const T* __restrict__ x, const float* __restrict__ ab,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         T* __restrict__ y, int B, int H, int W, int C,
                                                         long long ppb) {
  CRD_PDL_ENTRY();
  const int c = threadIdx.x * 8, ry = threadIdx.y, rows = blockDim.y;
  const int b = blockIdx.y;
  const long long N = (long long)H * W;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  float wt[9][8], a[8], sh[8], bs[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = ab[((long long)b * C + c + j) * 2];
    sh[j] = ab[((long long)b * C + c + j) * 2 + 1];
    bs[j] = bias ? bias[c + j] : 0.f;
#pragma unroll
    for (int t = 0; t < 9; t++) wt[t][j] = w[(c + j) * 9 + t];
  }
  const T* xb = x + (long long)b * N * C + c;
  T* yb = y + (long long)b * N * C + c;
  for (int p = (int)p0 + ry; p < (int)p1; p += rows) {
    const int hh = p / W, ww = p - hh * W;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = bs[j];
    // branch-free taps: clamped addresses are always loadable, out-of-image taps are masked (zero padding
    // applies to the NORMALISED tensor), so the nine 16-byte loads issue back to back
    typename Raw8<T>::type raw[9];
    bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int h2 = hh + t / 3 - 1, w2 = ww + t % 3 - 1;
      ok[t] = (h2 >= 0) & (h2 < H) & (w2 >= 0) & (w2 < W);
      const int hc = min(max(h2, 0), H - 1), wc = min(max(w2, 0), W - 1);
      raw[t] = ldg16(xb + (size_t)(hc * W + wc) * C);
    }
#pragma unroll
    for (int t = 0; t < 9; t++) {
      float v[8];
      unpack8(raw[t], v);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = fmaf(wt[t][j], ok[t] ? fmaf(a[j], v[j], sh[j]) : 0.f, acc[j]);
    }
    store8(yb + (size_t)p * C, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dwconv_bwd_input_kernel(const T* __restrict__ dy, const float* __restrict__ w,
                                                               T* __restrict__ dxn, int B, int H, int W, int C,
                                                               long long ppb) {
  CRD_PDL_ENTRY();
  const int c = threadIdx.x * 8, ry = threadIdx.y, rows = blockDim.y;
  const int b = blockIdx.y;
  const long long N = (long long)H * W;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  float wt[9][8];
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int t = 0; t < 9; t++) wt[t][j] = w[(c + j) * 9 + t];
  const T* gb = dy + (long long)b * N * C + c;
  T* ob = dxn + (long long)b * N * C + c;
  for (int p = (int)p0 + ry; p < (int)p1; p += rows) {
    const int hh = p / W, ww = p - hh * W;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    typename Raw8<T>::type raw[9];
    bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int h2 = hh - (t / 3 - 1), w2 = ww - (t % 3 - 1);   // output pixel that read this input with tap t
      ok[t] = (h2 >= 0) & (h2 < H) & (w2 >= 0) & (w2 < W);
      const int hc = min(max(h2, 0), H - 1), wc = min(max(w2, 0), W - 1);
      raw[t] = ldg16(gb + (size_t)(hc * W + wc) * C);
    }
#pragma unroll
    for (int t = 0; t < 9; t++) {
      float g[8];
      unpack8(raw[t], g);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = fmaf(wt[t][j], ok[t] ? g[j] : 0.f, acc[j]);
    }
    store8(ob + (size_t)p * C, acc);
  }
}

// dw[c][tap] += sum dy * xn(shifted); db[c] += sum dy.  blockDim = (cvec, rows), grid = (chunks, B)
template <typename T>
__global__ void dwconv_bwd_weight_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                         const float* __restrict__ ab, float* dw, float* db, int B, int H, int W,
                                         int C, long long ppb) {
  CRD_PDL_ENTRY();
  extern __shared__ float red[];
  const int cv = threadIdx.x, ry = threadIdx.y, rows = blockDim.y, cvec = blockDim.x;
  const int b = blockIdx.y;
  const int c = cv * 8;
  const long long N = (long long)H * W;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  float a[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = ab[((long long)b * C + c + j) * 2];
    sh[j] = ab[((long long)b * C + c + j) * 2 + 1];
  }
  float acc[10][8];
#pragma unroll
  for (int q = 0; q < 10; q++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[q][j] = 0.f;
  for (long long p = p0 + ry; p < p1; p += rows) {
    const int hh = (int)(p / W), ww = (int)(p % W);
    float g[8];
    load8(dy + ((long long)b * N + p) * C + c, g);
#pragma unroll
    for (int j = 0; j < 8; j++) acc[9][j] += g[j];
    typename Raw8<T>::type raw[9];
    bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int h2 = hh + t / 3 - 1, w2 = ww + t % 3 - 1;
      ok[t] = (h2 >= 0) & (h2 < H) & (w2 >= 0) & (w2 < W);
      const int hc = min(max(h2, 0), H - 1), wc = min(max(w2, 0), W - 1);
      raw[t] = ldg16(x + (((long long)b * H + hc) * W + wc) * C + c);
    }
#pragma unroll
    for (int t = 0; t < 9; t++) {
      float v[8];
      unpack8(raw[t], v);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[t][j] = fmaf(g[j], ok[t] ? fmaf(a[j], v[j], sh[j]) : 0.f, acc[t][j]);
    }
  }
#pragma unroll
  for (int q = 0; q < 10; q++) {
#pragma unroll
    for (int j = 0; j < 8; j++) red[(ry * cvec + cv) * 8 + j] = acc[q][j];
    __syncthreads();
    const int t = ry * cvec + cv, nt = rows * cvec;
    for (int cc = t; cc < C; cc += nt) {
      float s = 0.f;
      for (int r = 0; r < rows; r++) s += red[r * cvec * 8 + cc];
      if (q < 9) atomicAdd(dw + cc * 9 + q, s);
      else if (db) atomicAdd(db + cc, s);
    }
    __syncthreads();
  }
}

// Fused depthwise backward: one pass over dy produces BOTH the input gradient and the weight/bias gradients.
// With q the input pixel, dxn[q] = sum_t w[t] dy[q - t] and dw[t] += dy[q - t] * xn[q] use the same nine dy
// neighbours, so the separate weight-gradient kernel's nine extra loads per pixel disappear.
template <typename T>
__global__ void __launch_bounds__(256, 1) dwconv_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                  const float* __restrict__ ab,
                                                                  const float* __restrict__ w, T* __restrict__ dxn,
                                                                  float* dw, float* db, int B, int H, int W, int C,
                                                                  long long ppb) {
  CRD_PDL_ENTRY();
  extern __shared__ float red[];
  const int cv = threadIdx.x, ry = threadIdx.y, rows = blockDim.y, cvec = blockDim.x;
  const int c = cv * 8, b = blockIdx.y;
  const long long N = (long long)H * W;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  float wt[9][8], a[8], sh[8], acc[10][8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = ab[((long long)b * C + c + j) * 2];
    sh[j] = ab[((long long)b * C + c + j) * 2 + 1];
#pragma unroll
    for (int t = 0; t < 9; t++) wt[t][j] = w[(c + j) * 9 + t];
  }
#pragma unroll
  for (int q = 0; q < 10; q++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[q][j] = 0.f;
  const T* gb = dy + (long long)b * N * C + c;
  const T* xb = x + (long long)b * N * C + c;
  T* ob = dxn + (long long)b * N * C + c;
  for (int p = (int)p0 + ry; p < (int)p1; p += rows) {
    const int hh = p / W, ww = p - hh * W;
    typename Raw8<T>::type raw[9];
    bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int h2 = hh - (t / 3 - 1), w2 = ww - (t % 3 - 1);   // output pixel that read input p with tap t
      ok[t] = (h2 >= 0) & (h2 < H) & (w2 >= 0) & (w2 < W);
      const int hc = min(max(h2, 0), H - 1), wc = min(max(w2, 0), W - 1);
      raw[t] = ldg16(gb + (size_t)(hc * W + wc) * C);
    }
    float xv[8], o[8];
    load8(xb + (size_t)p * C, xv);
#pragma unroll
    for (int j = 0; j < 8; j++) { xv[j] = fmaf(a[j], xv[j], sh[j]); o[j] = 0.f; }
#pragma unroll
    for (int t = 0; t < 9; t++) {
      float g[8];
      unpack8(raw[t], g);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float gj = ok[t] ? g[j] : 0.f;
        o[j] = fmaf(wt[t][j], gj, o[j]);
        acc[t][j] = fmaf(gj, xv[j], acc[t][j]);
        if (t == 4) acc[9][j] += gj;                           // centre tap: dy at this pixel -> bias gradient
      }
    }
    store8(ob + (size_t)p * C, o);
  }
#pragma unroll
  for (int q = 0; q < 10; q++) {
#pragma unroll
    for (int j = 0; j < 8; j++) red[(ry * cvec + cv) * 8 + j] = acc[q][j];
    __syncthreads();
    const int t = ry * cvec + cv, nt = rows * cvec;
    for (int cc = t; cc < C; cc += nt) {
      float s = 0.f;
      for (int r = 0; r < rows; r++) s += red[r * cvec * 8 + cc];
      if (q < 9) atomicAdd(dw + cc * 9 + q, s);
      else if (db) atomicAdd(db + cc, s);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ attention score
// block: 64 tokens of one sample; loop heads; per head stage q[64][hd] and key chunks k[64][hd] in smem.
constexpr int ATN = 64, AMK = 64;
template <typename T>
__global__ void __launch_bounds__(256) attn_qkmax_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                             float* __restrict__ s, unsigned short* __restrict__ idx,
                                                             int B, int N, int M, int C, int heads, float scale) {
  CRD_PDL_ENTRY();
  extern __shared__ float sm[];
  const int hd = C / heads;
  const int hdp = hd + 1;
  float* qs = sm;                 // [ATN][hdp]
  float* ks = sm + ATN * hdp;     // [AMK][hdp]
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * ATN;
  const int tid = threadIdx.x;
  const int nl = tid >> 2, kl = tid & 3;
  float total = 0.f;
  for (int h = 0; h < heads; h++) {
    __syncthreads();
    for (int i = tid; i < ATN * hd; i += 256) {
      const int r = i / hd, dd = i - r * hd;
      const int n = n0 + r;
      qs[r * hdp + dd] = (n < N) ? to_f(q[((long long)b * N + n) * C + h * hd + dd]) : 0.f;
    }
    float best = -INFINITY;
    int besti = 0;
    for (int mc = 0; mc < M; mc += AMK) {
      __syncthreads();
      for (int i = tid; i < AMK * hd; i += 256) {
        const int r = i / hd, dd = i - r * hd;
        const int m = mc + r;
        ks[r * hdp + dd] = (m < M) ? to_f(k[((long long)b * M + m) * C + h * hd + dd]) : 0.f;
      }
      __syncthreads();
      const int mend = min(AMK, M - mc);
      for (int r = kl; r < mend; r += 4) {
        float dot = 0.f;
        const float* qp = qs + nl * hdp;
        const float* kp = ks + r * hdp;
#pragma unroll 8
        for (int dd = 0; dd < hd; dd++) dot = fmaf(qp[dd], kp[dd], dot);
        if (dot > best) { best = dot; besti = mc + r; }
      }
    }
    // combine the 4 key-lanes (first occurrence wins ties)
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    total += best;
    const int n = n0 + nl;
    if (kl == 0 && n < N) idx[((long long)b * heads + h) * N + n] = (unsigned short)besti;
  }
  const int n = n0 + nl;
  if (kl == 0 && n < N) s[(long long)b * N + n] = total * scale;
}

template <typename T>
__global__ void attn_qkmax_bwd_kernel(const float* __restrict__ ds, const T* __restrict__ q,
                                      const T* __restrict__ k, const unsigned short* __restrict__ idx,
                                      T* __restrict__ dq, float* dk, int B, int N, int M, int C, int heads,
                                      float scale) {
  CRD_PDL_ENTRY();
  const int cvec = C / 8;
  const int hd = C / heads;
  const long long total = (long long)B * N * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const long long tok = i / cvec;
    const int b = (int)(tok / N);
    const int n = (int)(tok - (long long)b * N);
    const int c = cv * 8;
    const int h = c / hd;           // hd is a multiple of 8
    const int m = idx[((long long)b * heads + h) * N + n];
    const float g = ds[tok] * scale;
    float kv[8], qv[8], o[8];
    load8(k + ((long long)b * M + m) * C + c, kv);
    load8(q + tok * C + c, qv);
#pragma unroll
    for (int j = 0; j < 8; j++) o[j] = g * kv[j];
    store8(dq + tok * C + c, o);
    float* dkp = dk + ((long long)b * M + m) * C + c;              // 32-byte aligned: two 16-byte vector reductions
#pragma unroll
    for (int j = 0; j < 8; j += 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dkp + j), "f"(g * qv[j]), "f"(g * qv[j + 1]),
                   "f"(g * qv[j + 2]), "f"(g * qv[j + 3])
                   : "memory");
  }
}

// ------------------------------------------------------------------ proj applied to the token mean
// one warp per (sample, output channel): B*C warps in flight instead of B blocks walking C outputs serially
__global__ void __launch_bounds__(256) attn_pv_fwd_kernel(const float* __restrict__ xbar, const float* __restrict__ Wp,
                                                          float* __restrict__ pv, int B, int C) {
  CRD_PDL_ENTRY();
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  if (o >= C) return;
  const float* wr = Wp + (long long)o * C;
  const float* xr = xbar + (long long)b * C;
  float acc = 0.f;
#pragma unroll 8
  for (int c = lane; c < C; c += 32) acc = fmaf(wr[c], xr[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) pv[(long long)b * C + o] = acc;
}
__global__ void attn_pv_bwd_w_kernel(const float* __restrict__ dpv, const float* __restrict__ xbar, float* dWp,
                                     int B, int C) {
  CRD_PDL_ENTRY();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)C * C) return;
  const int o = (int)(i / C), c = (int)(i % C);
  // blockIdx.y = slice of 8 samples: four times the thread count and a quarter of the dependent loop (the one-slice
  // version took ~10 us for a 64 x 64 matrix); the slices meet in the atomic
  const int b0 = blockIdx.y * 8, b1 = min(B, b0 + 8);
  float acc = 0.f;
#pragma unroll 8
  for (int b = b0; b < b1; b++) acc = fmaf(dpv[(long long)b * C + o], xbar[(long long)b * C + c], acc);
  atomicAdd(dWp + i, acc);
}
// block (32 channels, 8 slices of the output-channel sum) per (channel tile, sample)
__global__ void __launch_bounds__(256) attn_pv_bwd_x_kernel(const float* __restrict__ dpv, const float* __restrict__ Wp,
                                                            float* __restrict__ dxbar, float scale, int B, int C) {
  CRD_PDL_ENTRY();
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, os = threadIdx.x >> 5;
  const int b = blockIdx.y, c = blockIdx.x * 32 + cl;
  float acc = 0.f;
  if (c < C) {
    const float* g = dpv + (long long)b * C;
#pragma unroll 8
    for (int o = os; o < C; o += 8) acc = fmaf(Wp[(long long)o * C + c], g[o], acc);
  }
  red[os][cl] = acc;
  __syncthreads();
  if (os == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) t += red[k][cl];
    dxbar[(long long)b * C + c] = t * scale;
  }
}

__global__ void attn_out_residual_kernel(const float* __restrict__ x, const float* __restrict__ pv,
                                         const float* __restrict__ s, const float* __restrict__ bp,
                                         const float* __restrict__ dp, float* __restrict__ xout, int B, int N,
                                         int C) {
  CRD_PDL_ENTRY();
  const int c4 = C / 4;
  const long long total = (long long)B * N * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const long long tok = i / c4;
    const int b = (int)(tok / N);
    const float sc = dp ? dp[b] : 1.f;
    const float sv = s[tok];
    const float4 xv = *reinterpret_cast<const float4*>(x + tok * C + c);
    const float4 pvv = *reinterpret_cast<const float4*>(pv + (long long)b * C + c);
    const float4 bv = *reinterpret_cast<const float4*>(bp + c);
    float4 o;
    o.x = xv.x + sc * fmaf(pvv.x, sv, bv.x);
    o.y = xv.y + sc * fmaf(pvv.y, sv, bv.y);
    o.z = xv.z + sc * fmaf(pvv.z, sv, bv.z);
    o.w = xv.w + sc * fmaf(pvv.w, sv, bv.w);
    *reinterpret_cast<float4*>(xout + tok * C + c) = o;
  }
}

// ds[b][n] = dp[b] * sum_c dx[b][n][c] * pv[b][c]   (one warp per token)
__global__ void attn_out_bwd_ds_kernel(const float* __restrict__ dx, const float* __restrict__ pv,
                                       const float* __restrict__ dp, float* __restrict__ ds, int B, int N, int C) {
  CRD_PDL_ENTRY();
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)B * N) return;
  const int b = (int)(warp / N);
  float acc = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 g = *reinterpret_cast<const float4*>(dx + warp * C + c);
    const float4 p = *reinterpret_cast<const float4*>(pv + (long long)b * C + c);
    acc += g.x * p.x + g.y * p.y + g.z * p.z + g.w * p.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) ds[warp] = acc * (dp ? dp[b] : 1.f);
}
__global__ void attn_out_bwd_chan_kernel(const float* __restrict__ dx, const float* __restrict__ s, float* tmp,
                                         int B, long long N, int C, long long ppb) {
  CRD_PDL_ENTRY();
  chan_reduce2([&](int b, long long p, int c, float (&s0)[8], float (&s1)[8]) {
    float g[8];
    const long long tok = (long long)b * N + p;
    load8(dx + tok * C + c, g);
    const float sv = s[tok];
#pragma unroll
    for (int j = 0; j < 8; j++) { s0[j] = fmaf(g[j], sv, s0[j]); s1[j] += g[j]; }
  }, tmp, B, N, C, ppb);
}
// ds and the two per-channel sums from ONE pass over dx (C <= 256): a warp walks tokens, lane l holds channels
// 4l..4l+3 (+128); the token dot product is a warp sum, the channel sums stay in the lane's registers over the warp's
// tokens and meet in shared memory / one atomic per (block, channel).  grid = (token splits, B).
__global__ void __launch_bounds__(256) attn_out_bwd_fused_kernel(const float* __restrict__ dx,
                                                                 const float* __restrict__ pv,
                                                                 const float* __restrict__ s,
                                                                 const float* __restrict__ dp, float* __restrict__ ds,
                                                                 float* tmp, int N, int C, int ppb) {
  CRD_PDL_ENTRY();
  __shared__ float red[8][2][256];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = blockIdx.x * ppb, p1 = min(N, p0 + ppb);
  const float sc = dp ? dp[b] : 1.f;
  float4 pvv[2], a0[2], a1[2];
  bool on[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int c = lane * 4 + 128 * k;
    on[k] = c < C;
    pvv[k] = on[k] ? *reinterpret_cast<const float4*>(pv + (long long)b * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    a0[k] = a1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int p = p0 + warp; p < p1; p += 8) {
    const long long tok = (long long)b * N + p;
    const float sv = s[tok];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      if (!on[k]) continue;
      const float4 g = *reinterpret_cast<const float4*>(dx + tok * C + lane * 4 + 128 * k);
      dot += g.x * pvv[k].x + g.y * pvv[k].y + g.z * pvv[k].z + g.w * pvv[k].w;
      a0[k].x = fmaf(g.x, sv, a0[k].x); a0[k].y = fmaf(g.y, sv, a0[k].y);
      a0[k].z = fmaf(g.z, sv, a0[k].z); a0[k].w = fmaf(g.w, sv, a0[k].w);
      a1[k].x += g.x; a1[k].y += g.y; a1[k].z += g.z; a1[k].w += g.w;
    }
    dot = warp_sum(dot);
    if (lane == 0) ds[tok] = dot * sc;
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int c = lane * 4 + 128 * k;
    red[warp][0][c] = a0[k].x; red[warp][0][c + 1] = a0[k].y; red[warp][0][c + 2] = a0[k].z; red[warp][0][c + 3] = a0[k].w;
    red[warp][1][c] = a1[k].x; red[warp][1][c + 1] = a1[k].y; red[warp][1][c + 2] = a1[k].z; red[warp][1][c + 3] = a1[k].w;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * C; e += 256) {
    const int c = e >> 1, q = e & 1;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) t += red[w][q][c];
    atomicAdd(tmp + ((long long)b * C + c) * 2 + q, t);
  }
}
__global__ void attn_out_bwd_fin_kernel(const float* __restrict__ tmp, const float* __restrict__ dp,
                                        float* __restrict__ dpv, float* dbp, int B, int C) {
  CRD_PDL_ENTRY();
  // one thread per (sample, channel): the serial loop over the batch of the first version (C / 128 blocks, 32 dependent
  // iterations) took 11.5 us for 32 x 160 values
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const float sc = dp ? dp[b] : 1.f;
  const float2 t = *reinterpret_cast<const float2*>(tmp + (long long)i * 2);
  dpv[i] = sc * t.x;
  atomicAdd(dbp + c, sc * t.y);
}

template <typename T>
__global__ void residual_add_kernel(const float* __restrict__ x, const T* __restrict__ y,
                                    const float* __restrict__ dp, float* __restrict__ xout, int B, long long N,
                                    int C) {
  CRD_PDL_ENTRY();
  const int cvec = C / 8;
  const long long total = (long long)B * N * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i / cvec;
    const int b = (int)(tok / N);
    const float sc = dp ? dp[b] : 1.f;
    float xv[8], yv[8];
    load8(x + i * 8, xv);
    load8(y + i * 8, yv);
#pragma unroll
    for (int j = 0; j < 8; j++) xv[j] = fmaf(sc, yv[j], xv[j]);
    store8(xout + i * 8, xv);
  }
}
template <typename T>
__global__ void scale_cast_kernel(const float* __restrict__ dx, const float* __restrict__ dp, T* __restrict__ dy,
                                  int B, long long N, int C) {
  CRD_PDL_ENTRY();
  const int cvec = C / 8;
  const long long total = (long long)B * N * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i / cvec;
    const int b = (int)(tok / N);
    const float sc = dp ? dp[b] : 1.f;
    float v[8];
    load8(dx + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] *= sc;
    store8(dy + i * 8, v);
  }
}
template <typename T>
__global__ void add_f32_kernel(float* __restrict__ dst, const T* __restrict__ src, long long n8) {
  CRD_PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    load8(dst + i * 8, a);
    load8(src + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] += b[j];
    store8(dst + i * 8, a);
  }
}

}  // namespace

extern "C" int crd_dwconv3x3_fwd(const void* x, int dtype, const float* ab, const float* w, const float* bias,
                                 void* y, int B, int H, int W, int C, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256);
  if ((long long)B * H * W == 0) return 0;
  if (dw_tma_eligible(dtype, B, H, W, C, x, x, y)) {
    DwParams p = {};
    p.B = B; p.H = H; p.W = W; p.C = C; p.ab = ab; p.w = w; p.bias = bias; p.out = (bf16*)y;
    if (int e = dw_tma_launch<false>(x, nullptr, p, (cudaStream_t)stream)) return e;
    CRD_LAUNCH_CHECK();
    return 0;
  }
  ReduceLaunch r = plan_stream(B, (long long)H * W, C);
  CRD_DISPATCH_1(dtype, T, crd_launch(dwconv_fwd_kernel<T>, dim3(r.grid), dim3(r.block), 0, (cudaStream_t)stream, 
                               (const T*)x, ab, w, bias, (T*)y, B, H, W, C, r.ppb));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_dwconv3x3_bwd_input(const void* dy, int dtype, const float* w, void* dxn, int B, int H, int W,
                                       int C, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256);
  if ((long long)B * H * W == 0) return 0;
  ReduceLaunch r = plan_stream(B, (long long)H * W, C);
  CRD_DISPATCH_1(dtype, T, crd_launch(dwconv_bwd_input_kernel<T>, dim3(r.grid), dim3(r.block), 0, (cudaStream_t)stream, 
                               (const T*)dy, w, (T*)dxn, B, H, W, C, r.ppb));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_dwconv3x3_bwd_weight(const void* dy, int dtype, const void* x, const float* ab, float* dw,
                                        float* db, int B, int H, int W, int C, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256);
  if (B == 0 || H * W == 0) return 0;
  ReduceLaunch r = plan_reduce(B, (long long)H * W, C);
  const size_t smem = (size_t)r.block.x * r.block.y * 8 * sizeof(float);
  CRD_DISPATCH_1(dtype, T, crd_launch(dwconv_bwd_weight_kernel<T>, dim3(r.grid), dim3(r.block), smem, (cudaStream_t)stream, 
                               (const T*)dy, (const T*)x, ab, dw, db, B, H, W, C, r.ppb));
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_dwconv3x3_bwd(const void* dy, int dtype, const void* x, const float* ab, const float* w, void* dxn,
                                 float* dw, float* db, int B, int H, int W, int C, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256);
  if (B == 0 || H * W == 0) return 0;
  if (dw_tma_eligible(dtype, B, H, W, C, dy, x, dxn)) {
    DwParams p = {};
    p.B = B; p.H = H; p.W = W; p.C = C; p.ab = ab; p.w = w; p.out = (bf16*)dxn; p.dw = dw; p.db = db;
    if (int e = dw_tma_launch<true>(dy, x, p, (cudaStream_t)stream)) return e;
    CRD_LAUNCH_CHECK();
    return 0;
  }
  ReduceLaunch r = plan_reduce(B, (long long)H * W, C);
  const size_t smem = (size_t)r.block.x * r.block.y * 8 * sizeof(float);
  CRD_DISPATCH_1(dtype, T, crd_launch(dwconv_bwd_fused_kernel<T>, dim3(r.grid), dim3(r.block), smem, (cudaStream_t)stream, 
                               (const T*)dy, (const T*)x, ab, w, (T*)dxn, dw, db, B, H, W, C, r.ppb));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_attn_qkmax_fwd(const void* q, const void* k, int dtype, float* s, unsigned short* idx, int B,
                                  int N, int M, int C, int heads, float scale, crd_stream_t stream) {
  CRD_REQUIRE(heads > 0 && C % heads == 0 && M >= 1 && M <= 65535);   // idx is uint16
  if (B == 0 || N == 0) return 0;
  static int qk_tc = -1;
  if (qk_tc < 0) { const char* e = getenv("CAMRADEPTH_TC_QKMAX"); qk_tc = (e && e[0] == '0') ? 0 : 1; }
  if (qk_tc && dtype == CRD_BF16) {
    const int rc = crd_attn_qkmax_fwd_tc(q, k, s, idx, B, N, M, C, heads, scale, stream);
    if (rc <= 0) return rc;                       // 1 = shape not covered: CUDA-core kernel below
  }
  const int hd = C / heads;
  const size_t smem = (size_t)(ATN + AMK) * (hd + 1) * sizeof(float);
  dim3 grid(crd_div_up(N, ATN), B);
  CRD_DISPATCH_1(dtype, T, crd_launch(attn_qkmax_fwd_kernel<T>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, 
                               (const T*)q, (const T*)k, s, idx, B, N, M, C, heads, scale));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_attn_qkmax_bwd(const float* ds, const void* q, const void* k, int dtype,
                                  const unsigned short* idx, void* dq, float* dk, int B, int N, int M, int C,
                                  int heads, float scale, crd_stream_t stream) {
  CRD_REQUIRE(heads > 0 && C % heads == 0 && (C / heads) % 8 == 0);
  const long long total = (long long)B * N * (C / 8);
  if (total == 0) return 0;
  CRD_DISPATCH_1(dtype, T, crd_launch(attn_qkmax_bwd_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                               ds, (const T*)q, (const T*)k, idx, (T*)dq, dk, B, N, M, C, heads, scale));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_attn_pv_fwd(const float* xbar, const float* Wp, float* pv, int B, int C, crd_stream_t stream) {
  if (B == 0) return 0;
  crd_launch(attn_pv_fwd_kernel, dim3(dim3(crd_div_up(C, 8), B)), dim3(256), 0, (cudaStream_t)stream, xbar, Wp, pv, B, C);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_attn_pv_bwd(const float* dpv, const float* xbar, const float* Wp, float* dWp, float* dxbar,
                               float dxbar_scale, int B, int C, crd_stream_t stream) {
  if (B == 0) return 0;
  crd_launch(attn_pv_bwd_w_kernel, dim3(crd_div_up((long long)C * C, 256), crd_div_up(B, 8)), dim3(256), 0, (cudaStream_t)stream, dpv, xbar, dWp, B, C);
  CRD_LAUNCH_CHECK();
  crd_launch(attn_pv_bwd_x_kernel, dim3(dim3(crd_div_up(C, 32), B)), dim3(256), 0, (cudaStream_t)stream, dpv, Wp, dxbar, dxbar_scale, B, C);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_attn_out_residual(const float* x, const float* pv, const float* s, const float* bp,
                                     const float* dp, float* xout, int B, int N, int C, crd_stream_t stream) {
  CRD_REQUIRE(C % 4 == 0);
  const long long total = (long long)B * N * (C / 4);
  if (total == 0) return 0;
  crd_launch(attn_out_residual_kernel, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, x, pv, s, bp, dp, xout, B, N, C);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_attn_out_bwd(const float* dx, const float* pv, const float* s, const float* dp, float* ds,
                                float* dpv, float* dbp, float* tmp, int B, int N, int C, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && C / 8 <= 256);
  if (B == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(tmp, 0, (size_t)B * C * 2 * sizeof(float), st);
  static int fused = -1;
  if (fused < 0) { const char* e = getenv("CAMRADEPTH_ATTN_BWD_FUSED"); fused = (e && e[0] == '0') ? 0 : 1; }
  if (fused && C <= 256 && C % 4 == 0) {
    // token splits: enough blocks for two per SM, at least 32 tokens (4 per warp) each
    int splits = crd_div_up(2LL * sm_count(), B);
    const int most = crd_div_up(N, 32);
    if (splits > most) splits = most;
    if (splits < 1) splits = 1;
    const int ppb = crd_div_up(N, splits);
    crd_launch(attn_out_bwd_fused_kernel, dim3(crd_div_up(N, ppb), B), dim3(256), 0, st, dx, pv, s, dp, ds, tmp, N, C, ppb);
    CRD_LAUNCH_CHECK();
  } else {
    crd_launch(attn_out_bwd_ds_kernel, dim3(crd_div_up((long long)B * N * 32, 256)), dim3(256), 0, st, dx, pv, dp, ds, B, N, C);
    CRD_LAUNCH_CHECK();
    ReduceLaunch r = plan_reduce(B, N, C);
    crd_launch(attn_out_bwd_chan_kernel, dim3(r.grid), dim3(r.block), r.smem, st, dx, s, tmp, B, N, C, r.ppb);
    CRD_LAUNCH_CHECK();
  }
  crd_launch(attn_out_bwd_fin_kernel, dim3(crd_div_up((long long)B * C, 128)), dim3(128), 0, st, tmp, dp, dpv, dbp, B, C);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_residual_add(const float* x, const void* y, int y_dtype, const float* dp, float* xout, int B,
                                long long N, int C, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0);
  const long long total = (long long)B * N * (C / 8);
  if (total == 0) return 0;
  CRD_DISPATCH_1(y_dtype, T, crd_launch(residual_add_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                                 x, (const T*)y, dp, xout, B, N, C));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_scale_cast(const float* dx, const float* dp, void* dy, int dy_dtype, int B, long long N, int C,
                              crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0);
  const long long total = (long long)B * N * (C / 8);
  if (total == 0) return 0;
  CRD_DISPATCH_1(dy_dtype, T, crd_launch(scale_cast_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                                  dx, dp, (T*)dy, B, N, C));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_add_f32(float* dst, const void* src, int src_dtype, long long n, crd_stream_t stream) {
  CRD_REQUIRE(n % 8 == 0);
  if (n == 0) return 0;
  CRD_DISPATCH_1(src_dtype, T, crd_launch(add_f32_kernel<T>, dim3(ew_blocks(n / 8)), dim3(256), 0, (cudaStream_t)stream, 
                                   dst, (const T*)src, n / 8));
  CRD_LAUNCH_CHECK();
  return 0;
}
