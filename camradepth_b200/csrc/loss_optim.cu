// Fused losses and the multi-tensor diffGradNorm step (row a15).
// All HBM-bound single-pass kernels: no boolean-index compaction, no host syncs, no per-tensor launches.
#include "common.cuh"
#include "../../include/camradepth_b200.h"

namespace {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;   // valid in thread 0
}

__global__ void masked_l1_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                     float* acc, long long n) {
  CRD_PDL_ENTRY();
  __shared__ float sh[32];
  float s = 0.f, cnt = 0.f, sq = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float t = target[i];
    if (t > 0.f) {
      const float d = pred[i] - t;
      const float a = fabsf(d);
      s += (a < 1.f) ? 0.5f * d * d : a - 0.5f;
      sq = fmaf(d, d, sq);
      cnt += 1.f;
    }
  }
  s = block_sum(s, sh);
  cnt = block_sum(cnt, sh);
  sq = block_sum(sq, sh);
  if (threadIdx.x == 0) { atomicAdd(acc + 0, s); atomicAdd(acc + 1, cnt); atomicAdd(acc + 2, sq); }
}

__global__ void masked_l1_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                     const float* __restrict__ acc, const float* __restrict__ gout,
                                     float* __restrict__ dpred, long long n) {
  CRD_PDL_ENTRY();
  const float scale = gout[0] / acc[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float t = target[i];
    float g = 0.f;
    if (t > 0.f) {
      const float d = pred[i] - t;
      g = scale * fminf(fmaxf(d, -1.f), 1.f);
    }
    dpred[i] = g;
  }
}

__global__ void ce_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target, float* acc,
                              int B, int C, long long HW, int ignore_index) {
  CRD_PDL_ENTRY();
  __shared__ float sh[32];
  float s = 0.f, cnt = 0.f;
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = target[i];
    // labels outside [0, C) other than ignore_index make F.cross_entropy raise in the reference; here they are
    // treated like ignored pixels instead of indexing out of bounds
    if (t != ignore_index && t >= 0 && t < C) {
      const long long b = i / HW, r = i - b * HW;
      const float* lp = logits + b * C * HW + r;
      float m = -INFINITY;
      for (int c = 0; c < C; c++) m = fmaxf(m, lp[c * HW]);
      float se = 0.f;
      for (int c = 0; c < C; c++) se += __expf(lp[c * HW] - m);
      s += m + __logf(se) - lp[t * HW];
      cnt += 1.f;
    }
  }
  s = block_sum(s, sh);
  cnt = block_sum(cnt, sh);
  if (threadIdx.x == 0) { atomicAdd(acc + 0, s); atomicAdd(acc + 1, cnt); }
}

__device__ __forceinline__ float focal_dce(float ce, float gamma) {
  const float pt = __expf(-ce);
  const float om = 1.f - pt;
  // d/dce [(1-pt)^g * ce] = (1-pt)^g + g (1-pt)^(g-1) pt ce
  return powf(om, gamma) + gamma * powf(om, gamma - 1.f) * pt * ce;
}

__global__ void ce_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                              const float* __restrict__ acc, const float* __restrict__ gout, float gamma,
                              float* __restrict__ dlogits, int B, int C, long long HW, int ignore_index) {
  CRD_PDL_ENTRY();
  const float ce = acc[0] / acc[1];
  const float coef = gout[0] * (gamma > 0.f ? focal_dce(ce, gamma) : 1.f) / acc[1];
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = target[i];
    const long long b = i / HW, r = i - b * HW;
    const float* lp = logits + b * C * HW + r;
    float* dp = dlogits + b * C * HW + r;
    if (t == ignore_index || t < 0 || t >= C) {
      for (int c = 0; c < C; c++) dp[c * HW] = 0.f;
      continue;
    }
    float m = -INFINITY;
    for (int c = 0; c < C; c++) m = fmaxf(m, lp[c * HW]);
    float se = 0.f;
    for (int c = 0; c < C; c++) se += __expf(lp[c * HW] - m);
    const float inv = 1.f / se;
    for (int c = 0; c < C; c++) {
      const float p = __expf(lp[c * HW] - m) * inv;
      dp[c * HW] = coef * (p - (c == t ? 1.f : 0.f));
    }
  }
}

__global__ void loss_finalize_kernel(const float* __restrict__ acc, float* __restrict__ out, int kind,
                                     float gamma) {
  CRD_PDL_ENTRY();
  if (kind == 0) {
    out[0] = acc[0] / acc[1];
    out[1] = sqrtf(acc[2] / acc[1]);
  } else {
    const float ce = acc[0] / acc[1];
    const float om = 1.f - __expf(-ce);
    out[0] = powf(om, gamma) * ce;
    out[1] = ce;
  }
}

// ------------------------------------------------------------------ diffGradNorm
__global__ void __launch_bounds__(256) mt_sumsq_kernel(const crd_opt_tensor* __restrict__ table,
                                                       const crd_opt_chunk* __restrict__ chunks, float* sumsq) {
  CRD_PDL_ENTRY();
  __shared__ float sh[32];
  const crd_opt_chunk ck = chunks[blockIdx.x];
  const crd_opt_tensor t = table[ck.tensor];
  long long end = ck.start + CRD_OPT_CHUNK;
  if (end > t.numel) end = t.numel;
  float s = 0.f;
  for (long long i = ck.start + threadIdx.x; i < end; i += blockDim.x) { const float g = t.g[i]; s = fmaf(g, g, s); }
  s = block_sum(s, sh);
  // one partial per chunk, no atomics: the per-tensor total is formed in chunk order by the update kernel, so the
  // step is a deterministic function of the gradients and data-parallel replicas stay bit-identical
  if (threadIdx.x == 0) sumsq[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) diffgradnorm_kernel(const crd_opt_tensor* __restrict__ table,
                                                           const crd_opt_chunk* __restrict__ chunks,
                                                           const float* __restrict__ sumsq,
                                                           const float* __restrict__ egn_in, float* egn_out,
                                                           const float* __restrict__ step_size_p, float beta1,
                                                           float beta2, float eps) {
  CRD_PDL_ENTRY();
  const float step_size = step_size_p[0];
  const crd_opt_chunk ck = chunks[blockIdx.x];
  const crd_opt_tensor t = table[ck.tensor];
  long long end = ck.start + CRD_OPT_CHUNK;
  if (end > t.numel) end = t.numel;
  // Gradient-norm correction (diffGradNorm.py:82-88); the branch is evaluated on the device.
  // ||g||^2 of the tensor = its chunk partials (consecutive in the chunk list) summed in order
  __shared__ float tot_sh;
  if (threadIdx.x == 0) {
    const int idx = (int)(ck.start / CRD_OPT_CHUNK);
    const int cnt = (int)((t.numel + CRD_OPT_CHUNK - 1) / CRD_OPT_CHUNK);
    const float* pp = sumsq + ((long long)blockIdx.x - idx);
    float a = 0.f;
    for (int k = 0; k < cnt; k++) a += pp[k];
    tot_sh = a;
  }
  __syncthreads();
  const float gn = sqrtf(tot_sh);
  const float egn = 0.95f * egn_in[ck.tensor] + 0.05f * gn;
  const float corr = (egn > gn) ? egn / (gn + 1e-8f) : 1.f;
  if (ck.start == 0 && threadIdx.x == 0) egn_out[ck.tensor] = egn;
  auto upd = [&](float g, float& m, float& v, float& prev, float& pw) {
    m = beta1 * m + (1.f - beta1) * (g * corr);
    v = beta2 * v + (1.f - beta2) * g * g;
    const float denom = sqrtf(v) + eps;
    const float dfc = 1.f / (1.f + __expf(-fabsf(prev - g)));
    prev = g;
    pw = pw - step_size * (m * dfc) / denom;
  };
  long long i0 = ck.start;
  // nine 4-byte streams per element: 16-byte accesses where the five arrays of the tensor allow it
  const uintptr_t al = reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                       reinterpret_cast<uintptr_t>(t.v) | reinterpret_cast<uintptr_t>(t.prev) |
                       reinterpret_cast<uintptr_t>(t.p);
  if ((al & 15) == 0) {
    const long long end4 = ck.start + ((end - ck.start) & ~3LL);
    for (long long i = ck.start + 4 * threadIdx.x; i < end4; i += 4 * blockDim.x) {
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 m = *reinterpret_cast<const float4*>(t.m + i), v = *reinterpret_cast<const float4*>(t.v + i);
      float4 pr = *reinterpret_cast<const float4*>(t.prev + i), pw = *reinterpret_cast<const float4*>(t.p + i);
      upd(g.x, m.x, v.x, pr.x, pw.x); upd(g.y, m.y, v.y, pr.y, pw.y);
      upd(g.z, m.z, v.z, pr.z, pw.z); upd(g.w, m.w, v.w, pr.w, pw.w);
      *reinterpret_cast<float4*>(t.m + i) = m; *reinterpret_cast<float4*>(t.v + i) = v;
      *reinterpret_cast<float4*>(t.prev + i) = pr; *reinterpret_cast<float4*>(t.p + i) = pw;
    }
    i0 = end4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    const float g = t.g[i];
    float m = t.m[i], v = t.v[i], pr = t.prev[i], pw = t.p[i];
    upd(g, m, v, pr, pw);
    t.m[i] = m; t.v[i] = v; t.prev[i] = pr; t.p[i] = pw;
  }
}

// ------------------------------------------------------------------ test-mode metrics (runner.py:442-492)
// pred is clipped to [0,1]; both are scaled by max_depth; valid = 0 < gt <= thr1 (runner.py:455-457 zeroes
// gt > max_distances[0] first), and additionally gt >= thr2 for the second set (:473-475).
// acc[0..3] = (sum sq, sum abs, sum rel, count) within max range; acc[4..7] = same for gt >= thr2.
__global__ void depth_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* acc,
                                     long long n, float max_depth, float thr1, float thr2) {
  CRD_PDL_ENTRY();
  __shared__ float sh[32];
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float g = gt[i] * max_depth;
    if (g > 0.f && !(g > thr1)) {
      const float p = fminf(fmaxf(pred[i], 0.f), 1.f) * max_depth;
      const float e = p - g, ae = fabsf(e), re = ae / g;
      a[0] = fmaf(e, e, a[0]); a[1] += ae; a[2] += re; a[3] += 1.f;
      if (g >= thr2) { a[4] = fmaf(e, e, a[4]); a[5] += ae; a[6] += re; a[7] += 1.f; }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float r = block_sum(a[j], sh);
    if (threadIdx.x == 0) atomicAdd(acc + j, r);
  }
}
// out[0..2] = RMSE, MAE, REL (max range); out[3..5] = same (second range)
__global__ void depth_metrics_finalize_kernel(const float* __restrict__ acc, float* __restrict__ out) {
  CRD_PDL_ENTRY();
  for (int s = 0; s < 2; s++) {
    const float c = acc[4 * s + 3];
    out[3 * s + 0] = sqrtf(acc[4 * s + 0] / c);
    out[3 * s + 1] = acc[4 * s + 1] / c;
    out[3 * s + 2] = acc[4 * s + 2] / c;
  }
}
// conf[t][p] += 1 for every pixel with target t != ignore_index and p = argmax_c logits
__global__ void confusion_kernel(const float* __restrict__ logits, const long long* __restrict__ target, float* conf,
                                 int B, int C, long long HW, int ignore_index) {
  CRD_PDL_ENTRY();
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = target[i];
    if (t == ignore_index || t < 0 || t >= C) continue;
    const long long b = i / HW, r = i - b * HW;
    const float* lp = logits + b * C * HW + r;
    float best = lp[0];
    int bi = 0;
    for (int c = 1; c < C; c++) { const float v = lp[c * HW]; if (v > best) { best = v; bi = c; } }
    atomicAdd(conf + t * C + bi, 1.f);
  }
}

// ------------------------------------------------------------------ input pipeline (dataloader.py:213-245)
// inverse-normalised lidar GT: g = d > 0 ? (max_depth - min(d, max_depth)) / max_depth : 0
__global__ void gt_normalize_kernel(const float* __restrict__ d, float* __restrict__ g, long long n, float max_depth) {
  CRD_PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = fminf(fmaxf(d[i], 0.f), max_depth);
    g[i] = v > 0.f ? (max_depth - v) / max_depth : 0.f;
  }
}
// zero-ignoring 3x3 stride-2 pad-1 min-pool (zeros count as "no measurement"; 255 never survives)
__global__ void minpool_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int Ho, int Wo) {
  CRD_PDL_ENTRY();
  const long long total = (long long)B * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % Wo), oh = (int)((i / Wo) % Ho);
    const long long b = i / ((long long)Wo * Ho);
    float m = 255.f;
    for (int dh = -1; dh <= 1; dh++) {
      const int h = 2 * oh + dh;
      if (h < 0 || h >= H) continue;
      for (int dw = -1; dw <= 1; dw++) {
        const int w = 2 * ow + dw;
        if (w < 0 || w >= W) continue;
        const float v = x[(b * H + h) * W + w];
        m = fminf(m, v == 0.f ? 255.f : v);
      }
    }
    y[i] = m == 255.f ? 0.f : m;
  }
}
// uint8 HWC image -> fp32 channels [c0, c0+3) of an NCHW batch tensor, (v/255 - mean[c]) / std[c]
__global__ void image_normalize_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, int B, int H,
                                       int W, int Ctot, float m0, float m1, float m2, float s0, float s1, float s2) {
  CRD_PDL_ENTRY();
  const long long hw = (long long)H * W, total = (long long)B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, r = i - b * hw;
    const unsigned char* p = img + i * 3;
    float* o = out + b * Ctot * hw + r;
    o[0] = ((float)p[0] * (1.f / 255.f) - m0) / s0;
    o[hw] = ((float)p[1] * (1.f / 255.f) - m1) / s1;
    o[2 * hw] = ((float)p[2] * (1.f / 255.f) - m2) / s2;
  }
}

// nearest-neighbour resize of a label map with the index rule of skimage.transform.resize(order=0,
// anti_aliasing=False) = scipy.ndimage.zoom(order=0, grid_mode=True): src = floor((dst + 0.5) * in / out), clamped
// (dataloader.py:262-268: the 416x800 and 208x400 segmentation ground truths)
template <typename TS>
__global__ void seg_resize_kernel(const TS* __restrict__ src, long long* __restrict__ dst, int B, int Hi, int Wi, int Ho,
                                  int Wo) {
  CRD_PDL_ENTRY();
  const long long total = (long long)B * Ho * Wo;
  const double sh = (double)Hi / (double)Ho, sw = (double)Wi / (double)Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % Wo), oh = (int)((i / Wo) % Ho);
    const long long b = i / ((long long)Wo * Ho);
    int ih = (int)floor(((double)oh + 0.5) * sh), iw = (int)floor(((double)ow + 0.5) * sw);
    ih = ih < Hi ? ih : Hi - 1; iw = iw < Wi ? iw : Wi - 1;
    dst[i] = (long long)src[(b * Hi + ih) * Wi + iw];
  }
}

// Network input straight in the engine's layout: NHWC bf16 with `ld` channels per pixel = [normalised RGB | extra
// float planes (radar depth / u / v / velocity, already scaled by the caller like dataloader.py:301-323) | zeros]
__global__ void pack_input_nhwc_kernel(const unsigned char* __restrict__ img, const float* __restrict__ extra,
                                       bf16* __restrict__ dst, int B, int H, int W, int Ce, int ld, float m0, float m1,
                                       float m2, float s0, float s1, float s2) {
  CRD_PDL_ENTRY();
  const long long hw = (long long)H * W, total = (long long)B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, r = i - b * hw;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = 0.f;
    const unsigned char* p = img + i * 3;
    v[0] = ((float)p[0] * (1.f / 255.f) - m0) / s0;
    v[1] = ((float)p[1] * (1.f / 255.f) - m1) / s1;
    v[2] = ((float)p[2] * (1.f / 255.f) - m2) / s2;
    for (int c = 0; c < Ce && c < 5; c++) v[3 + c] = extra[(b * Ce + c) * hw + r];
    store8(dst + i * ld, v);
    for (int c = 8; c < ld; c += 8) {
      float z[8];
#pragma unroll
      for (int j = 0; j < 8; j++) z[j] = 0.f;
      store8(dst + i * ld + c, z);
    }
  }
}

// ---- stochastic masks (timm DropPath per call, nn.Dropout2d per (sample, channel) plane) from ONE launch ----------
// Philox4x32-10 keyed by (seed, step counter); the counter lives in device memory and is advanced by the kernel
// itself, so a CUDA-graph replay draws fresh masks.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// out[i] = u_i < keep[row(i)] ? 1 / keep : 0 ; rows [0, n_dp) have B entries (keep_dp[row]), the following n_d2 rows
// have B*C2 entries (keep = keep_d2).  Single block: thread 0 advances the counter after everyone has read it.
__global__ void make_masks_kernel(float* __restrict__ out, const float* __restrict__ keep_dp, int n_dp, int B, int n_d2,
                                  int C2, float keep_d2, unsigned long long* __restrict__ state) {
  CRD_PDL_ENTRY();
  const unsigned long long seed = state[0], step = state[1];
  const long long n_a = (long long)n_dp * B, total = n_a + (long long)n_d2 * B * C2;
  for (long long q = threadIdx.x; q * 4 < total; q += blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const long long i = q * 4 + j;
      if (i >= total) break;
      const float u = (float)(c[j] >> 8) * (1.0f / 16777216.0f);          // [0, 1)
      const float keep = i < n_a ? keep_dp[i / B] : keep_d2;
      out[i] = u < keep ? 1.0f / keep : 0.f;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) state[1] = step + 1;
}

inline int red_blocks(long long n) {
  long long b = (n + 1023) / 1024;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace

extern "C" int crd_masked_l1_fwd(const float* pred, const float* target, float* acc, long long n,
                                 crd_stream_t stream) {
  if (n == 0) return 0;
  crd_launch(masked_l1_fwd_kernel, dim3(red_blocks(n)), dim3(256), 0, (cudaStream_t)stream, pred, target, acc, n);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_masked_l1_bwd(const float* pred, const float* target, const float* acc, const float* gout,
                                 float* dpred, long long n, crd_stream_t stream) {
  if (n == 0) return 0;
  crd_launch(masked_l1_bwd_kernel, dim3(red_blocks(n)), dim3(256), 0, (cudaStream_t)stream, pred, target, acc, gout, dpred, n);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_ce_fwd(const float* logits, const long long* target, float* acc, int B, int C, long long HW,
                          int ignore_index, crd_stream_t stream) {
  if ((long long)B * HW == 0) return 0;
  crd_launch(ce_fwd_kernel, dim3(red_blocks((long long)B * HW)), dim3(256), 0, (cudaStream_t)stream, logits, target, acc, B, C, HW,
                                                                                 ignore_index);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_ce_bwd(const float* logits, const long long* target, const float* acc, const float* gout,
                          float gamma, float* dlogits, int B, int C, long long HW, int ignore_index,
                          crd_stream_t stream) {
  if ((long long)B * HW == 0) return 0;
  crd_launch(ce_bwd_kernel, dim3(red_blocks((long long)B * HW)), dim3(256), 0, (cudaStream_t)stream, logits, target, acc, gout, gamma,
                                                                                 dlogits, B, C, HW, ignore_index);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_loss_finalize(const float* acc, float* out, int kind, float gamma, crd_stream_t stream) {
  crd_launch(loss_finalize_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, acc, out, kind, gamma);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_mt_sumsq(const crd_opt_tensor* table, const crd_opt_chunk* chunks, int nchunks, float* sumsq,
                            crd_stream_t stream) {
  if (nchunks == 0) return 0;
  crd_launch(mt_sumsq_kernel, dim3(nchunks), dim3(256), 0, (cudaStream_t)stream, table, chunks, sumsq);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_diffgradnorm_update(const crd_opt_tensor* table, const crd_opt_chunk* chunks, int nchunks,
                                       const float* sumsq, const float* egn_in, float* egn_out,
                                       const float* step_size, float beta1, float beta2, float eps,
                                       crd_stream_t stream) {
  if (nchunks == 0) return 0;
  crd_launch(diffgradnorm_kernel, dim3(nchunks), dim3(256), 0, (cudaStream_t)stream, table, chunks, sumsq, egn_in, egn_out, step_size,
                                                                 beta1, beta2, eps);
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_depth_metrics(const float* pred, const float* gt, float* acc, float* out, long long n,
                                 float max_depth, float thr1, float thr2, crd_stream_t stream) {
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(acc, 0, 8 * sizeof(float), s);
  crd_launch(depth_metrics_kernel, dim3(red_blocks(n)), dim3(256), 0, s, pred, gt, acc, n, max_depth, thr1, thr2);
  CRD_LAUNCH_CHECK();
  crd_launch(depth_metrics_finalize_kernel, dim3(1), dim3(1), 0, s, acc, out);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_confusion(const float* logits, const long long* target, float* conf, int B, int C, long long HW,
                             int ignore_index, crd_stream_t stream) {
  if ((long long)B * HW == 0) return 0;
  crd_launch(confusion_kernel, dim3(red_blocks((long long)B * HW)), dim3(256), 0, (cudaStream_t)stream, logits, target, conf, B, C, HW,
                                                                                    ignore_index);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_gt_normalize(const float* d, float* g, long long n, float max_depth, crd_stream_t stream) {
  if (n == 0) return 0;
  crd_launch(gt_normalize_kernel, dim3(red_blocks(n)), dim3(256), 0, (cudaStream_t)stream, d, g, n, max_depth);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_minpool3x3s2(const float* x, float* y, int B, int H, int W, crd_stream_t stream) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)B * Ho * Wo;
  if (total == 0) return 0;
  crd_launch(minpool_kernel, dim3(red_blocks(total)), dim3(256), 0, (cudaStream_t)stream, x, y, B, H, W, Ho, Wo);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_seg_resize_nearest(const void* src, int src_is_u8, long long* dst, int B, int Hi, int Wi, int Ho,
                                      int Wo, crd_stream_t stream) {
  const long long total = (long long)B * Ho * Wo;
  if (total == 0) return 0;
  CRD_REQUIRE(Hi > 0 && Wi > 0);
  if (src_is_u8) crd_launch(seg_resize_kernel<unsigned char>, dim3(red_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
        (const unsigned char*)src, dst, B, Hi, Wi, Ho, Wo);
  else crd_launch(seg_resize_kernel<long long>, dim3(red_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
        (const long long*)src, dst, B, Hi, Wi, Ho, Wo);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_pack_input_nhwc(const unsigned char* img, const float* extra, void* dst, int B, int H, int W, int Ce,
                                   int ld, const float* mean3_host, const float* std3_host, crd_stream_t stream) {
  const long long total = (long long)B * H * W;
  if (total == 0) return 0;
  CRD_REQUIRE(img && dst && ld >= 8 && ld % 8 == 0 && Ce >= 0 && Ce <= 5 && (Ce == 0 || extra));
  crd_launch(pack_input_nhwc_kernel, dim3(red_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
      img, extra, (bf16*)dst, B, H, W, Ce, ld, mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0], std3_host[1],
      std3_host[2]);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_make_masks(float* out, const float* keep_dp, int n_dp, int B, int n_d2, int C2, float keep_d2,
                              unsigned long long* state, crd_stream_t stream) {
  CRD_REQUIRE(out && state && (n_dp == 0 || keep_dp) && keep_d2 > 0.f);
  if ((long long)n_dp * B + (long long)n_d2 * B * C2 == 0) return 0;
  crd_launch(make_masks_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, out, keep_dp, n_dp, B, n_d2, C2, keep_d2, state);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_image_normalize(const unsigned char* img, float* out, int B, int H, int W, int Ctot,
                                   const float* mean3_host, const float* std3_host, crd_stream_t stream) {
  const long long total = (long long)B * H * W;
  if (total == 0) return 0;
  crd_launch(image_normalize_kernel, dim3(red_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
      img, out, B, H, W, Ctot, mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0], std3_host[1], std3_host[2]);
  CRD_LAUNCH_CHECK();
  return 0;
}
