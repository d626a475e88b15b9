// Decoder pieces that are not plain contractions: bicubic x2 upsample
// written straight into a channel slice of the concat buffer, the Cout=1 depth-head stencil, argmax
// segmentation maps, the NCHW<->NHWC boundary converts and weight (un)packing.
#include "common.cuh"
#include "chan_reduce.cuh"
#include "bicubic_tma.cuh"
#include "../../include/camradepth_b200.h"

namespace {

// PyTorch upsample_bicubic2d, align_corners=False, scale 2, A=-0.75 (utils.py:241):
// even output 2y: taps rows y-2..y+1 with t=.75 ; odd output 2y+1: rows y-1..y+2 with t=.25
__device__ __forceinline__ void bicubic_taps(int o, int& base, float (&cf)[4]) {
  const float c25[4] = {-0.10546875f, 0.87890625f, 0.26171875f, -0.03515625f};
  if (o & 1) {
    base = (o >> 1) - 1;
#pragma unroll
    for (int i = 0; i < 4; i++) cf[i] = c25[i];
  } else {
    base = (o >> 1) - 2;
#pragma unroll
    for (int i = 0; i < 4; i++) cf[i] = c25[3 - i];
  }
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <typename T>
__global__ void bicubic_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C,
                                   int ldx, int ldy) {
  CRD_PDL_ENTRY();
  const int cvec = C / 8;
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)B * Ho * Wo * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int ox = (int)(pix % Wo);
    const int oy = (int)((pix / Wo) % Ho);
    const int b = (int)(pix / ((long long)Wo * Ho));
    int by, bx;
    float cy[4], cx[4];
    bicubic_taps(oy, by, cy);
    bicubic_taps(ox, bx, cx);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
    for (int iy = 0; iy < 4; iy++) {
      const int sy = clampi(by + iy, 0, H - 1);
#pragma unroll
      for (int ix = 0; ix < 4; ix++) {
        const int sx = clampi(bx + ix, 0, W - 1);
        float v[8];
        load8(x + (((long long)b * H + sy) * W + sx) * ldx + cv * 8, v);
        const float wgt = cy[iy] * cx[ix];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = fmaf(wgt, v[j], acc[j]);
      }
    }
    store8(y + pix * ldy + cv * 8, acc);
  }
}

// weight with which output coordinate o reads input coordinate s (border clamping folded in)
__device__ __forceinline__ float bicubic_weight(int o, int s, int n) {
  int base;
  float cf[4];
  bicubic_taps(o, base, cf);
  float w = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) w += (clampi(base + i, 0, n - 1) == s) ? cf[i] : 0.f;
  return w;
}

template <typename T>
__global__ void bicubic_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int accumulate, int B, int H,
                                   int W, int C, int lddy, int lddx) {
  CRD_PDL_ENTRY();
  const int cvec = C / 8;
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)B * H * W * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int sx = (int)(pix % W);
    const int sy = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float wy[8], wx[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const int oy = 2 * sy - 3 + t, ox = 2 * sx - 3 + t;
      wy[t] = (oy >= 0 && oy < Ho) ? bicubic_weight(oy, sy, H) : 0.f;
      wx[t] = (ox >= 0 && ox < Wo) ? bicubic_weight(ox, sx, W) : 0.f;
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
    for (int ty = 0; ty < 8; ty++) {
      const int oy = min(max(2 * sy - 3 + ty, 0), Ho - 1);        // weight is 0 wherever the clamp acts
      typename Raw8<T>::type raw[8];
#pragma unroll
      for (int tx = 0; tx < 8; tx++) {
        const int ox = min(max(2 * sx - 3 + tx, 0), Wo - 1);
        raw[tx] = ldg16(dy + (((long long)b * Ho + oy) * Wo + ox) * lddy + cv * 8);
      }
#pragma unroll
      for (int tx = 0; tx < 8; tx++) {
        float g[8];
        unpack8(raw[tx], g);
        const float wgt = wy[ty] * wx[tx];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = fmaf(wgt, g[j], acc[j]);
      }
    }
    T* o = dx + pix * lddx + cv * 8;
    if (accumulate) {
      float old[8];
      load8(o, old);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] += old[j];
    }
    store8(o, acc);
  }
}

// ------------------------------------------------------------------ 3x3 conv with a single output channel
template <typename T>
__global__ void conv_c1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W,
                                   int Cin, int ldx) {
  CRD_PDL_ENTRY();
  extern __shared__ float ws[];   // [9][Cin]
  for (int i = threadIdx.x; i < 9 * Cin; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const long long total = (long long)B * H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int ww = (int)(pix % W);
    const int hh = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    // eight independent accumulators (one per lane of the 16-byte vector) instead of one 288-long FMA chain;
    // clamped addresses + masks keep the nine taps branch-free
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int h2 = hh + t / 3 - 1, w2 = ww + t % 3 - 1;
      const bool ok = (h2 >= 0) & (h2 < H) & (w2 >= 0) & (w2 < W);
      const int hc = min(max(h2, 0), H - 1), wc = min(max(w2, 0), W - 1);
      const T* xp = x + (((long long)b * H + hc) * W + wc) * ldx;
      const float* wp = ws + t * Cin;
      for (int c = 0; c < Cin; c += 8) {
        float v[8];
        load8(xp + c, v);
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = fmaf(ok ? v[j] : 0.f, wp[c + j], acc[j]);
      }
    }
    float tot = bias ? bias[0] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) tot += acc[j];
    y[pix] = tot;
  }
}

// ---- the production shape (Depth_Activation.conv_2: 32 bf16 channels -> 1) --------------------------------------
// The generic kernels above / below spend ~1700 (forward) and ~470 (input gradient) instructions per thread on index
// arithmetic, per-tap weight loads and one load per (tap, 8 channels): ncu showed L1 at 91 % / issue slots at 78 % for
// 164 MB tensors that stream in 30 us.  These variants share the loads:
//   forward : four lanes own FOUR adjacent output pixels of a row, one lane per 8-channel group; per input row a lane
//             loads its 16 bytes of the six input columns once for 96 FMAs against weights held in registers (18
//             instead of 36 sixteen-byte loads per pixel, each warp load = 8 runs of 64 contiguous bytes)
//   backward: a thread keeps the 9x8 weights of ITS channel group in registers over the whole grid-stride loop and
//             multiplies by the sigmoid derivative on the way out (no dsg buffer, no separate sigmoid pass)
constexpr int C1_CIN = 32;

__global__ void __launch_bounds__(256, 2) conv_c1_fwd32_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ y,
                                                               int B, int H, int W, int ldx) {
  CRD_PDL_ENTRY();
  // four lanes per pixel quad, one per 8-channel group: a warp's 16-byte loads of one (row, column) cover 8 runs of
  // 64 contiguous bytes (with one lane per quad they were 32 separate lines per load: L1 at 84-91 %), and the lane's
  // 9x8 weights stay in registers for the whole grid-stride loop (no shared memory)
  const int lane = threadIdx.x & 31, cg = lane & 3;
  float2 wreg[9][4];                      // channel pairs: even / odd channels accumulate in the halves of one FFMA2
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int j = 0; j < 4; j++) wreg[t][j] = make_float2(w[t * C1_CIN + cg * 8 + 2 * j], w[t * C1_CIN + cg * 8 + 2 * j + 1]);
  const int Wg = (W + 3) >> 2;
  const long long total = (long long)B * H * Wg;
  const float b0 = bias ? bias[0] : 0.f;
  const long long stride_q = ((long long)gridDim.x * blockDim.x) >> 2;
  for (long long q0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2) - (lane >> 2); q0 < total;
       q0 += stride_q) {                                  // q0 is warp-uniform: the shuffles below see all lanes
    const long long q = q0 + (lane >> 2);
    const bool qok = q < total;
    const long long qq = qok ? q : 0;
    const int g = (int)(qq % Wg);
    const long long bh = qq / Wg;                // b * H + h
    const int h = (int)(bh % H);
    const int w0 = g * 4;
    float2 acc2[4];
#pragma unroll
    for (int pz = 0; pz < 4; pz++) acc2[pz] = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const int hh = h + r - 1;
      if (!qok || hh < 0 || hh >= H) continue;
      const bf16* rowp = x + (bh + (r - 1)) * (long long)W * ldx + cg * 8;
      uint4 raw[6];
#pragma unroll
      for (int d = 0; d < 6; d++) {
        const int col = w0 + d - 1;
        raw[d] = (col >= 0 && col < W) ? ldg16(rowp + (long long)col * ldx) : make_uint4(0, 0, 0, 0);
      }
      float2 v[6][4];
#pragma unroll
      for (int d = 0; d < 6; d++) {
        float t8[8];
        unpack8(raw[d], t8);
#pragma unroll
        for (int j = 0; j < 4; j++) v[d][j] = make_float2(t8[2 * j], t8[2 * j + 1]);
      }
#pragma unroll
      for (int pz = 0; pz < 4; pz++)
#pragma unroll
        for (int kw = 0; kw < 3; kw++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc2[pz] = __ffma2_rn(v[pz + kw][j], wreg[r * 3 + kw][j], acc2[pz]);
    }
    float acc[4];
#pragma unroll
    for (int pz = 0; pz < 4; pz++) {
      acc[pz] = acc2[pz].x + acc2[pz].y;
      acc[pz] += __shfl_xor_sync(0xffffffffu, acc[pz], 1);
      acc[pz] += __shfl_xor_sync(0xffffffffu, acc[pz], 2);
    }
    // lane cg stores pixel w0 + cg: 32 consecutive floats per warp
    const float mine = cg == 0 ? acc[0] : (cg == 1 ? acc[1] : (cg == 2 ? acc[2] : acc[3]));
    if (qok && w0 + cg < W) y[bh * W + w0 + cg] = mine + b0;
  }
}

// dx[pix][c] = (sum_taps dy[pix - tap] w[tap][c]) * (SIG ? s (1 - s) : 1),  s = x[pix][c] (the sigmoid OUTPUT)
template <bool SIG>
__global__ void __launch_bounds__(256) conv_c1_bwd_input32_kernel(const float* __restrict__ dy,
                                                                  const float* __restrict__ w,
                                                                  const bf16* __restrict__ x, bf16* __restrict__ dx,
                                                                  int B, int H, int W, int ldx, int lddx) {
  CRD_PDL_ENTRY();
  // blockDim and the grid stride are multiples of 4: a thread's channel group never changes
  const int cv = threadIdx.x & 3;
  float2 wreg[9][4];                      // channel pairs: the taps run on FFMA2
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int j = 0; j < 4; j++) wreg[t][j] = make_float2(w[t * C1_CIN + cv * 8 + 2 * j], w[t * C1_CIN + cv * 8 + 2 * j + 1]);
  const long long npix = (long long)B * H * W;
  for (long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2; pix < npix;
       pix += ((long long)gridDim.x * blockDim.x) >> 2) {
    const int ww = (int)(pix % W);
    const long long bh = pix / W;
    const int hh = (int)(bh % H);
    const float* dyp = dy + pix;
    float2 acc2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) acc2[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int dh = t / 3 - 1, dw = t % 3 - 1;
      const int h2 = hh - dh, w2 = ww - dw;
      const bool ok = (h2 >= 0) & (h2 < H) & (w2 >= 0) & (w2 < W);
      const float g = ok ? dyp[-(dh * W + dw)] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; j++) acc2[j] = __ffma2_rn(f2(g), wreg[t][j], acc2[j]);
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 4; j++) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
    if (SIG) {
      float sv[8];
      load8(x + pix * ldx + cv * 8, sv);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] *= sv[j] * (1.f - sv[j]);
    }
    store8(dx + pix * lddx + cv * 8, acc);
  }
}

template <typename T>
__global__ void conv_c1_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                         T* __restrict__ dx, int B, int H, int W, int Cin, int lddx) {
  CRD_PDL_ENTRY();
  const int cvec = Cin / 8;
  const long long total = (long long)B * H * W * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int ww = (int)(pix % W);
    const int hh = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    for (int dh = -1; dh <= 1; dh++) {
      const int h2 = hh - dh;
      if (h2 < 0 || h2 >= H) continue;
      for (int dw = -1; dw <= 1; dw++) {
        const int w2 = ww - dw;
        if (w2 < 0 || w2 >= W) continue;
        const float g = dy[((long long)b * H + h2) * W + w2];
        const float* wp = w + ((dh + 1) * 3 + (dw + 1)) * Cin + cv * 8;
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = fmaf(g, wp[j], acc[j]);
      }
    }
    store8(dx + pix * lddx + cv * 8, acc);
  }
}

template <typename T>
__global__ void conv_c1_bwd_weight_kernel(const float* __restrict__ dy, const T* __restrict__ x, float* dw,
                                          float* db, int B, int H, int W, int Cin, int ldx, long long ppb) {
  CRD_PDL_ENTRY();
  extern __shared__ float red[];
  const int cv = threadIdx.x, ry = threadIdx.y, rows = blockDim.y, cvec = blockDim.x;
  const int b = blockIdx.y;
  const long long N = (long long)H * W;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > N) p1 = N;
  float acc[10][8];
#pragma unroll
  for (int q = 0; q < 10; q++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[q][j] = 0.f;
  // dw[tap] = sum_p dy[p] x[p + tap] = sum_q x[q] dy[q - tap]: ONE 16-byte load of x per pixel, the nine dy scalars
  // (1 channel, fp32) come from L1; clamped addresses + masks keep the loop branch-free
  const float* dyb = dy + (long long)b * N;
  const T* xb = x + (long long)b * N * ldx + cv * 8;
  for (long long p = p0 + ry; p < p1; p += rows) {
    const int hh = (int)(p / W), ww = (int)(p % W);
    const typename Raw8<T>::type raw = ldg16(xb + p * ldx);
    float g[9];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int h2 = hh - (t / 3 - 1), w2 = ww - (t % 3 - 1);
      const bool ok = (h2 >= 0) & (h2 < H) & (w2 >= 0) & (w2 < W);
      const int hc = min(max(h2, 0), H - 1), wc = min(max(w2, 0), W - 1);
      const float gv = dyb[(long long)hc * W + wc];
      g[t] = ok ? gv : 0.f;
    }
    float v[8];
    unpack8(raw, v);
    if (cv == 0) acc[9][0] += g[4];
#pragma unroll
    for (int t = 0; t < 9; t++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[t][j] = fmaf(g[t], v[j], acc[t][j]);
  }
#pragma unroll
  for (int q = 0; q < 10; q++) {
#pragma unroll
    for (int j = 0; j < 8; j++) red[(ry * cvec + cv) * 8 + j] = acc[q][j];
    __syncthreads();
    const int t = ry * cvec + cv, nt = rows * cvec;
    for (int cc = t; cc < Cin; cc += nt) {
      float s = 0.f;
      for (int r = 0; r < rows; r++) s += red[r * cvec * 8 + cc];
      if (q < 9) atomicAdd(dw + q * Cin + cc, s);
      else if (db && cc == 0) atomicAdd(db, s * 1.0f);
    }
    __syncthreads();
  }
}

// Weight gradient of the 32-channel case: dw[tap][c] = sum_q x[q][c] dy[q - tap].  A thread (channel group cv, strip)
// walks 16 consecutive pixels of one image row with the 3x3 neighbourhood of dy in registers: three new dy values per
// pixel instead of nine clamped loads (the generic kernel spends ~260 instructions per pixel and channel group).
constexpr int C1_STRIP = 16;
__global__ void __launch_bounds__(256, 2) conv_c1_bwd_weight32_kernel(const float* __restrict__ dy,
                                                                      const bf16* __restrict__ x, float* dw, float* db,
                                                                      int B, int H, int W, int ldx) {
  CRD_PDL_ENTRY();
  __shared__ float red[256 * 8];
  const int cv = threadIdx.x, ry = threadIdx.y, rows = blockDim.y;       // blockDim = (4, 64)
  const int spr = (W + C1_STRIP - 1) / C1_STRIP;                           // strips per image row
  const long long nstrips = (long long)B * H * spr;
  float2 acc[10][4];                      // channel pairs (FFMA2)
#pragma unroll
  for (int q = 0; q < 10; q++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[q][j] = make_float2(0.f, 0.f);
  for (long long sidx = (long long)blockIdx.x * rows + ry; sidx < nstrips; sidx += (long long)gridDim.x * rows) {
    const int sw = (int)(sidx % spr);
    const long long bh = sidx / spr;                     // b * H + h
    const int hh = (int)(bh % H);
    const int w0 = sw * C1_STRIP, w1 = min(W, w0 + C1_STRIP);
    // G[a][c] = dy[b][hh - 1 + a][ww - 1 + c] (0 outside the image); tap t = (kh, kw) reads G[2 - kh][2 - kw]
    const float* drow[3];
    bool rok[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int h2 = hh - 1 + a;
      rok[a] = h2 >= 0 && h2 < H;
      drow[a] = dy + (bh + (a - 1)) * W;
    }
    float G[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      G[a][0] = 0.f;
      G[a][1] = (rok[a] && w0 - 1 >= 0) ? drow[a][w0 - 1] : 0.f;
      G[a][2] = rok[a] ? drow[a][w0] : 0.f;
    }
    const bf16* xp = x + (bh * W + w0) * (long long)ldx + cv * 8;
    for (int ww = w0; ww < w1; ww++, xp += ldx) {
      const uint4 raw = ldg16(xp);
#pragma unroll
      for (int a = 0; a < 3; a++) {
        G[a][0] = G[a][1]; G[a][1] = G[a][2];
        G[a][2] = (rok[a] && ww + 1 < W) ? drow[a][ww + 1] : 0.f;
      }
      float v[8];
      unpack8(raw, v);
      float2 v2[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v2[j] = make_float2(v[2 * j], v[2 * j + 1]);
#pragma unroll
      for (int t = 0; t < 9; t++) {
        const float2 g = f2(G[2 - t / 3][2 - t % 3]);
#pragma unroll
        for (int j = 0; j < 4; j++) acc[t][j] = __ffma2_rn(g, v2[j], acc[t][j]);
      }
      if (cv == 0) acc[9][0].x += G[1][1];
    }
  }
  const int cvec = 4, Cin = C1_CIN;
#pragma unroll
  for (int q = 0; q < 10; q++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      red[(ry * cvec + cv) * 8 + 2 * j] = acc[q][j].x;
      red[(ry * cvec + cv) * 8 + 2 * j + 1] = acc[q][j].y;
    }
    __syncthreads();
    const int t = ry * cvec + cv, nt = rows * cvec;
    for (int cc = t; cc < Cin; cc += nt) {
      float sum = 0.f;
      for (int r = 0; r < rows; r++) sum += red[r * cvec * 8 + cc];
      if (q < 9) atomicAdd(dw + q * Cin + cc, sum);
      else if (db && cc == 0) atomicAdd(db, sum);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void sigmoid_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx,
                                   long long n8) {
  CRD_PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float g[8], s[8];
    load8(dy + i * 8, g);
    load8(y + i * 8, s);
#pragma unroll
    for (int j = 0; j < 8; j++) g[j] = g[j] * s[j] * (1.f - s[j]);
    store8(dx + i * 8, g);
  }
}

template <typename T, typename TD>
__global__ void argmax_map_kernel(const T* __restrict__ logits, int ld, int ncls, TD* dst, int ld_dst,
                                  float* dst_f32, long long npix) {
  CRD_PDL_ENTRY();
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const T* lp = logits + pix * ld;
    float best = to_f(lp[0]);
    int bi = 0;
    for (int c = 1; c < ncls; c++) {
      const float v = to_f(lp[c]);
      if (v > best) { best = v; bi = c; }
    }
    const float m = (float)bi / (float)ncls;
    if (dst) dst[pix * ld_dst] = from_f<TD>(m);
    if (dst_f32) dst_f32[pix] = m;
  }
}

// ------------------------------------------------------------------ boundary layout converts
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int B, int C, int H,
                                    int W, int ld) {
  CRD_PDL_ENTRY();
  const long long hw = (long long)H * W;
  const long long total = (long long)B * hw;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long b = pix / hw, r = pix - b * hw;
    for (int c = 0; c < C; c++) dst[pix * ld + c] = from_f<T>(src[(b * C + c) * hw + r]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int B, int C, int H,
                                    int W, int ld) {
  CRD_PDL_ENTRY();
  const long long hw = (long long)H * W;
  const long long total = (long long)B * hw;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long b = pix / hw, r = pix - b * hw;
    for (int c = 0; c < C; c++) dst[(b * C + c) * hw + r] = to_f(src[pix * ld + c]);
  }
}

template <typename T>
__global__ void weight_pack_kernel(const float* __restrict__ w, T* __restrict__ dst, const int* __restrict__ map,
                                   int Cout, int Cin, int taps, int Cin_p, int Cout_p, int mode) {
  CRD_PDL_ENTRY();
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long long)taps * Cin));
    const int cm = map ? map[ci] : ci;
    const long long o = mode == 0 ? ((long long)co * taps + tap) * Cin_p + cm
                      : mode == 1 ? ((long long)cm * taps + tap) * Cout_p + co
                                  : ((long long)tap * Cin_p + cm) * Cout_p + co;
    dst[o] = from_f<T>(w[i]);
  }
}
// Every packed weight copy of the model in ONE launch (the per-tensor launcher costs ~3.5 us x 380 tensors per
// training step).  table: n+1 rows of 12 int64 {w, dst, map, first block, Cout, Cin, taps, Cin_p, Cout_p, mode,
// dst dtype, elements}; row n carries the total block count; then one int64 per block = its item.  One block packs one (output-channel tile x
// input-channel tile x all taps) brick through shared memory: the source brick is read as contiguous runs of
// ci_t * taps floats per output channel, the destination is written with the packed layout's fastest index across
// the threads (ci for mode 0, co for modes 1 / 2).  The first version walked the SOURCE order and scattered
// 2-byte stores Cin_p elements apart: 0.34 ms per step for 130 MB.
__host__ __device__ inline void wp_tile(int Cin, int taps, int& ci_t, int& co_t) {
  ci_t = Cin < 64 ? Cin : 64;
  const int c = 8192 / (ci_t * taps);
  co_t = c < 1 ? 1 : (c > 64 ? 64 : c);
}
__global__ void __launch_bounds__(256) weight_pack_batch_kernel(const long long* __restrict__ table, int n) {
  CRD_PDL_ENTRY();
  __shared__ float sm[8192 + 64];
  // block -> item map behind the rows (one load; a binary search over the rows was 9 dependent L2 round trips per block)
  const int item = (int)table[(long long)(n + 1) * 12 + blockIdx.x];
  const long long* t = table + (long long)item * 12;
  const float* w = reinterpret_cast<const float*>(t[0]);
  const int* map = reinterpret_cast<const int*>(t[2]);
  const int Cout = (int)t[4], Cin = (int)t[5], taps = (int)t[6], Cin_p = (int)t[7], Cout_p = (int)t[8];
  const int mode = (int)t[9], dtype = (int)t[10];
  int ci_t, co_t;
  wp_tile(Cin, taps, ci_t, co_t);
  const int tiles_ci = (Cin + ci_t - 1) / ci_t;
  const int tb = (int)((long long)blockIdx.x - t[3]);
  const int co0 = (tb / tiles_ci) * co_t, ci0 = (tb % tiles_ci) * ci_t;
  if (co0 >= Cout) return;
  const int nco = min(co_t, Cout - co0), nci = min(ci_t, Cin - ci0);
  const int run = nci * taps, run_p = run | 1;          // odd row pitch: the co-fastest read-out is conflict free
  for (int idx = threadIdx.x; idx < nco * run; idx += 256) {
    const int co = idx / run, r = idx - co * run;
    sm[co * run_p + r] = w[((long long)(co0 + co) * Cin + ci0) * taps + r];
  }
  __syncthreads();
  const int total = nco * run;
  for (int e = threadIdx.x; e < total; e += 256) {
    int co, ci, tap;
    if (mode == 0) { ci = e % nci; const int q = e / nci; tap = q % taps; co = q / taps; }
    else { co = e % nco; const int q = e / nco; tap = q % taps; ci = q / taps; }
    const float v = sm[co * run_p + ci * taps + tap];
    const int cm = map ? map[ci0 + ci] : ci0 + ci;
    const int cog = co0 + co;
    const long long o = mode == 0 ? ((long long)cog * taps + tap) * Cin_p + cm
                      : mode == 1 ? ((long long)cm * taps + tap) * Cout_p + cog
                                  : ((long long)tap * Cin_p + cm) * Cout_p + cog;
    if (dtype == CRD_BF16) reinterpret_cast<bf16*>(t[1])[o] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(t[1])[o] = v;
  }
}
__global__ void weight_unpack_grad_kernel(const float* __restrict__ dwp, float* __restrict__ grad,
                                          const int* __restrict__ map, int Cout, int Cin, int taps, int Cin_p,
                                          int accumulate) {
  CRD_PDL_ENTRY();
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long long)taps * Cin));
    const int cm = map ? map[ci] : ci;
    const float v = dwp[((long long)co * taps + tap) * Cin_p + cm];
    grad[i] = accumulate ? grad[i] + v : v;
  }
}

// ------------------------------------------------------------------ strided convs as GEMMs
// col[m][(kh*KW+kw)*Cin + c] = x[b, oh*stride-pad+kh, ow*stride-pad+kw, c]   (zero outside the image)
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, T* __restrict__ col, int B, int H, int W, int Cin, int ldx,
                              int Ho, int Wo, int KH, int KW, int stride, int pad) {
  CRD_PDL_ENTRY();
  const int cvec = Cin / 8;
  const int taps = KH * KW;
  const long long total = (long long)B * Ho * Wo * taps * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long t = i / cvec;
    const int tap = (int)(t % taps);
    const long long m = t / taps;
    const int ow = (int)(m % Wo);
    const int oh = (int)((m / Wo) % Ho);
    const int b = (int)(m / ((long long)Wo * Ho));
    const int kh = tap / KW, kw = tap - kh * KW;
    const int ih = oh * stride - pad + kh, iw = ow * stride - pad + kw;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = 0.f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) load8(x + (((long long)b * H + ih) * W + iw) * ldx + cv * 8, v);
    store8(col + (m * taps + tap) * Cin + cv * 8, v);
  }
}
// dx[b,ih,iw,c] (+)= sum over the (oh,ow,kh,kw) that read it of dcol[m][(kh*KW+kw)*Cin + c]  (gather form)
template <typename T>
__global__ void col2im_kernel(const T* __restrict__ dcol, T* __restrict__ dx, int accumulate, int B, int H, int W,
                              int Cin, int lddx, int Ho, int Wo, int KH, int KW, int stride, int pad) {
  CRD_PDL_ENTRY();
  const int cvec = Cin / 8;
  const int taps = KH * KW;
  const long long total = (long long)B * H * W * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int iw = (int)(pix % W);
    const int ih = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    for (int kh = 0; kh < KH; kh++) {
      const int th = ih + pad - kh;
      if (th < 0 || th % stride) continue;
      const int oh = th / stride;
      if (oh >= Ho) continue;
      for (int kw = 0; kw < KW; kw++) {
        const int tw = iw + pad - kw;
        if (tw < 0 || tw % stride) continue;
        const int ow = tw / stride;
        if (ow >= Wo) continue;
        const long long m = ((long long)b * Ho + oh) * Wo + ow;
        float g[8];
        load8(dcol + (m * taps + kh * KW + kw) * Cin + cv * 8, g);
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] += g[j];
      }
    }
    T* o = dx + pix * lddx + cv * 8;
    if (accumulate) {
      float old[8];
      load8(o, old);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] += old[j];
    }
    store8(o, acc);
  }
}

// db[n] += sum_m dy[m][n]; blockDim = (cvec, rows)
template <typename T>
__global__ void col_sum_kernel(const T* __restrict__ dy, float* db, long long M, int N, int ld, long long ppb) {
  CRD_PDL_ENTRY();
  extern __shared__ float red[];
  const int cv = threadIdx.x, ry = threadIdx.y, rows = blockDim.y, cvec = blockDim.x;
  long long p0 = (long long)blockIdx.x * ppb, p1 = p0 + ppb;
  if (p1 > M) p1 = M;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; j++) s[j] = 0.f;
  for (long long p = p0 + ry; p < p1; p += rows) {
    float v[8];
    load8(dy + p * ld + cv * 8, v);
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; j++) red[(ry * cvec + cv) * 8 + j] = s[j];
  __syncthreads();
  const int t = ry * cvec + cv, nt = rows * cvec;
  for (int c = t; c < N; c += nt) {
    float a = 0.f;
    for (int r = 0; r < rows; r++) a += red[r * cvec * 8 + c];
    atomicAdd(db + c, a);
  }
}

}  // namespace

extern "C" int crd_bicubic2x_fwd(const void* x, void* y, int dtype, int B, int H, int W, int C, int ldx, int ldy,
                                 crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0);
  const long long total = (long long)B * H * W * 4 * (C / 8);
  if (total == 0) return 0;
  if (bc_tma_eligible(dtype, B, H, W, C, x, ldx, y, ldy)) {
    BcParams p = {};
    p.B = B; p.H = H; p.W = W; p.C = C; p.ld_in = ldx; p.ld_out = ldy; p.out = (bf16*)y;
    if (int e = bc_tma_launch<false>(x, p, (cudaStream_t)stream)) return e;
    CRD_LAUNCH_CHECK();
    return 0;
  }
  CRD_DISPATCH_1(dtype, T, crd_launch(bicubic_fwd_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                               (const T*)x, (T*)y, B, H, W, C, ldx, ldy));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_bicubic2x_bwd(const void* dy, void* dx, int dtype, int accumulate, int B, int H, int W, int C,
                                 int lddy, int lddx, crd_stream_t stream) {
  CRD_REQUIRE(C % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0);
  const long long total = (long long)B * H * W * (C / 8);
  if (total == 0) return 0;
  if (bc_tma_eligible(dtype, B, H, W, C, dy, lddy, dx, lddx)) {
    BcParams p = {};
    p.B = B; p.H = H; p.W = W; p.C = C; p.ld_in = lddy; p.ld_out = lddx; p.accumulate = accumulate; p.out = (bf16*)dx;
    if (int e = bc_tma_launch<true>(dy, p, (cudaStream_t)stream)) return e;
    CRD_LAUNCH_CHECK();
    return 0;
  }
  CRD_DISPATCH_1(dtype, T, crd_launch(bicubic_bwd_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                               (const T*)dy, (T*)dx, accumulate, B, H, W, C, lddy, lddx));
  CRD_LAUNCH_CHECK();
  return 0;
}
static bool c1_fast() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("CAMRADEPTH_C1_FAST"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
extern "C" int crd_conv3x3_c1_fwd(const void* x, int dtype, const float* w, const float* bias, float* y, int B,
                                  int H, int W, int Cin, int ldx, crd_stream_t stream) {
  CRD_REQUIRE(Cin % 8 == 0 && ldx % 8 == 0 && Cin <= 1024);
  const long long total = (long long)B * H * W;
  if (total == 0) return 0;
  if (dtype == CRD_BF16 && Cin == C1_CIN && c1_fast() && ((uintptr_t)x & 15) == 0) {
    const long long quads = (long long)B * H * ((W + 3) / 4);
    crd_launch(conv_c1_fwd32_kernel, dim3(ew_blocks(quads * 4)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, w,
               bias, y, B, H, W, ldx);
    CRD_LAUNCH_CHECK();
    return 0;
  }
  CRD_DISPATCH_1(dtype, T, crd_launch(conv_c1_fwd_kernel<T>, dim3(ew_blocks(total)), dim3(256), 9 * Cin * sizeof(float), (cudaStream_t)stream, (const T*)x, w, bias, y, B, H, W, Cin, ldx));
  CRD_LAUNCH_CHECK();
  return 0;
}
static int conv3x3_c1_bwd_impl(const float* dy, const void* x, int dtype, const float* w, void* dx, float* dw,
                               float* db, int B, int H, int W, int Cin, int ldx, int lddx, int sig,
                               crd_stream_t stream) {
  CRD_REQUIRE(Cin % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0 && Cin / 8 <= 256);
  const long long total = (long long)B * H * W * (Cin / 8);
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dx && dtype == CRD_BF16 && Cin == C1_CIN && c1_fast() && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dx & 15) == 0) {
    if (sig) crd_launch(conv_c1_bwd_input32_kernel<true>, dim3(ew_blocks(total)), dim3(256), 0, s, dy, w, (const bf16*)x,
                        (bf16*)dx, B, H, W, ldx, lddx);
    else crd_launch(conv_c1_bwd_input32_kernel<false>, dim3(ew_blocks(total)), dim3(256), 0, s, dy, w, (const bf16*)x,
                    (bf16*)dx, B, H, W, ldx, lddx);
    CRD_LAUNCH_CHECK();
  } else if (dx) {
    CRD_DISPATCH_1(dtype, T, crd_launch(conv_c1_bwd_input_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, s, 
                                 dy, w, (T*)dx, B, H, W, Cin, lddx));
    CRD_LAUNCH_CHECK();
    if (sig) {            // generic route: the separate sigmoid-derivative pass, in place (needs a dense dx)
      CRD_REQUIRE(lddx == Cin && ldx == Cin);
      CRD_DISPATCH_1(dtype, T, crd_launch(sigmoid_bwd_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, s, (const T*)dx,
                                          (const T*)x, (T*)dx, total));
      CRD_LAUNCH_CHECK();
    }
  }
  if (dw && dtype == CRD_BF16 && Cin == C1_CIN && c1_fast() && ((uintptr_t)x & 15) == 0) {
    const long long nstrips = (long long)B * H * ((W + C1_STRIP - 1) / C1_STRIP);
    long long blocks = (nstrips + 63) / 64;
    const long long cap = 2LL * sm_count();                 // two resident blocks per SM, grid-stride beyond that
    if (blocks > cap) blocks = cap;
    crd_launch(conv_c1_bwd_weight32_kernel, dim3((unsigned)blocks), dim3(4, 64), 0, s, dy, (const bf16*)x, dw, db, B, H, W,
               ldx);
    CRD_LAUNCH_CHECK();
  } else if (dw) {
    ReduceLaunch r = plan_reduce(B, (long long)H * W, Cin);
    const size_t smem = (size_t)r.block.x * r.block.y * 8 * sizeof(float);
    CRD_DISPATCH_1(dtype, T, crd_launch(conv_c1_bwd_weight_kernel<T>, dim3(r.grid), dim3(r.block), smem, s, 
                                 dy, (const T*)x, dw, db, B, H, W, Cin, ldx, r.ppb));
    CRD_LAUNCH_CHECK();
  }
  return 0;
}
extern "C" int crd_conv3x3_c1_bwd(const float* dy, const void* x, int dtype, const float* w, void* dx, float* dw,
                                  float* db, int B, int H, int W, int Cin, int ldx, int lddx,
                                  crd_stream_t stream) {
  return conv3x3_c1_bwd_impl(dy, x, dtype, w, dx, dw, db, B, H, W, Cin, ldx, lddx, 0, stream);
}
// as crd_conv3x3_c1_bwd with x = the OUTPUT of the sigmoid that feeds the conv (Depth_Activation, utils.py:285-289):
// dx receives the gradient with respect to the sigmoid's INPUT, dx = dgrad * x * (1 - x)
extern "C" int crd_conv3x3_c1_bwd_sigmoid(const float* dy, const void* x, int dtype, const float* w, void* dx,
                                          float* dw, float* db, int B, int H, int W, int Cin, int ldx, int lddx,
                                          crd_stream_t stream) {
  CRD_REQUIRE(dx != nullptr);
  return conv3x3_c1_bwd_impl(dy, x, dtype, w, dx, dw, db, B, H, W, Cin, ldx, lddx, 1, stream);
}
extern "C" int crd_sigmoid_bwd(const void* dy, const void* y, void* dx, int dtype, long long n,
                               crd_stream_t stream) {
  CRD_REQUIRE(n % 8 == 0);
  if (n == 0) return 0;
  CRD_DISPATCH_1(dtype, T, crd_launch(sigmoid_bwd_kernel<T>, dim3(ew_blocks(n / 8)), dim3(256), 0, (cudaStream_t)stream, 
                               (const T*)dy, (const T*)y, (T*)dx, n / 8));
  CRD_LAUNCH_CHECK();
  return 0;
}
namespace {
// one thread per pixel, C (small, any count / alignment) consecutive channels
template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ src, int ld_src, T* __restrict__ dst, int ld_dst, int C,
                                     long long npix) {
  CRD_PDL_ENTRY();
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const T* s = src + pix * ld_src;
    T* d = dst + pix * ld_dst;
    for (int c = 0; c < C; c++) d[c] = s[c];
  }
}
}  // namespace
namespace {
// zero C consecutive channels of every pixel: one 16-byte store per pixel where the slice allows it (the tail
// channels of the [features | depth | seg maps | pad] buffers), element stores otherwise
template <typename T>
__global__ void zero_channels_kernel(T* __restrict__ dst, int ld, int C, long long npix, int vec) {
  CRD_PDL_ENTRY();
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    T* d = dst + pix * ld;
    if (vec) {
      for (int c = 0; c < C; c += 16 / (int)sizeof(T)) *reinterpret_cast<uint4*>(d + c) = make_uint4(0, 0, 0, 0);
    } else {
      for (int c = 0; c < C; c++) d[c] = from_f<T>(0.f);
    }
  }
}
}  // namespace
extern "C" int crd_zero_channels(void* dst, int ld, int dtype, int C, long long npix, crd_stream_t stream) {
  CRD_REQUIRE(dst && C >= 0 && ld >= C);
  if (npix == 0 || C == 0) return 0;
  const int es = dtype == CRD_BF16 ? 2 : 4;
  const int vec = (((uintptr_t)dst & 15) == 0 && (ld * es) % 16 == 0 && (C * es) % 16 == 0) ? 1 : 0;
  CRD_DISPATCH_1(dtype, T, crd_launch(zero_channels_kernel<T>, dim3(ew_blocks(npix)), dim3(256), 0, (cudaStream_t)stream,
                                      (T*)dst, ld, C, npix, vec));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_copy_channels(const void* src, int ld_src, void* dst, int ld_dst, int dtype, int C, long long npix,
                                 crd_stream_t stream) {
  CRD_REQUIRE(src && dst && C >= 0);
  if (npix == 0 || C == 0) return 0;
  CRD_DISPATCH_1(dtype, T, crd_launch(copy_channels_kernel<T>, dim3(ew_blocks(npix)), dim3(256), 0, (cudaStream_t)stream, 
                               (const T*)src, ld_src, (T*)dst, ld_dst, C, npix));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_argmax_map(const void* logits, int dtype, int ld, int ncls, void* dst, int dst_dtype,
                              int ld_dst, float* dst_f32, long long npix, crd_stream_t stream) {
  if (npix == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  CRD_DISPATCH_1(dtype, T, CRD_DISPATCH_1(dst_dtype, TD, crd_launch(argmax_map_kernel<T, TD>, dim3(ew_blocks(npix)), dim3(256), 0, s, 
                               (const T*)logits, ld, ncls, (TD*)dst, ld_dst, dst_f32, npix)));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int B, int C, int H, int W,
                                int ld_dst, crd_stream_t stream) {
  const long long total = (long long)B * H * W;
  if (total == 0) return 0;
  CRD_DISPATCH_1(dst_dtype, T, crd_launch(nchw_to_nhwc_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                                   src, (T*)dst, B, C, H, W, ld_dst));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int B, int C, int H, int W,
                                int ld_src, crd_stream_t stream) {
  const long long total = (long long)B * H * W;
  if (total == 0) return 0;
  CRD_DISPATCH_1(src_dtype, T, crd_launch(nhwc_to_nchw_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                                   (const T*)src, dst, B, C, H, W, ld_src));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_weight_pack(const float* w, void* dst, int dst_dtype, const int* map, int Cout, int Cin,
                               int taps, int Cin_p, int Cout_p, int mode, crd_stream_t stream) {
  const long long total = (long long)Cout * Cin * taps;
  if (total == 0) return 0;
  CRD_DISPATCH_1(dst_dtype, T, crd_launch(weight_pack_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                                   w, (T*)dst, map, Cout, Cin, taps, Cin_p, Cout_p, mode));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_weight_pack_blocks(int Cout, int Cin, int taps) {
  if (Cout <= 0 || Cin <= 0 || taps <= 0 || taps > 128) return -1;
  int ci_t, co_t;
  wp_tile(Cin, taps, ci_t, co_t);
  return ((Cout + co_t - 1) / co_t) * ((Cin + ci_t - 1) / ci_t);
}
extern "C" int crd_weight_pack_batch(const long long* table, int n_items, int n_blocks, crd_stream_t stream) {
  CRD_REQUIRE(table != nullptr || n_items == 0);
  if (n_items <= 0 || n_blocks <= 0) return 0;
  crd_launch(weight_pack_batch_kernel, dim3(n_blocks), dim3(256), 0, (cudaStream_t)stream, table, n_items);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_weight_unpack_grad(const float* dwp, float* grad, const int* map, int Cout, int Cin, int taps,
                                      int Cin_p, int accumulate, crd_stream_t stream) {
  const long long total = (long long)Cout * Cin * taps;
  if (total == 0) return 0;
  crd_launch(weight_unpack_grad_kernel, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, dwp, grad, map, Cout, Cin, taps,
                                                                               Cin_p, accumulate);
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_im2col(const void* x, void* col, int dtype, int B, int H, int W, int Cin, int ldx, int Ho, int Wo,
                          int KH, int KW, int stride, int pad, crd_stream_t stream) {
  CRD_REQUIRE(Cin % 8 == 0 && ldx % 8 == 0);
  const long long total = (long long)B * Ho * Wo * KH * KW * (Cin / 8);
  if (total == 0) return 0;
  CRD_DISPATCH_1(dtype, T, crd_launch(im2col_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                               (const T*)x, (T*)col, B, H, W, Cin, ldx, Ho, Wo, KH, KW, stride, pad));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_col2im(const void* dcol, void* dx, int dtype, int accumulate, int B, int H, int W, int Cin,
                          int lddx, int Ho, int Wo, int KH, int KW, int stride, int pad, crd_stream_t stream) {
  CRD_REQUIRE(Cin % 8 == 0 && lddx % 8 == 0);
  const long long total = (long long)B * H * W * (Cin / 8);
  if (total == 0) return 0;
  CRD_DISPATCH_1(dtype, T, crd_launch(col2im_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, 
                               (const T*)dcol, (T*)dx, accumulate, B, H, W, Cin, lddx, Ho, Wo, KH, KW, stride, pad));
  CRD_LAUNCH_CHECK();
  return 0;
}
extern "C" int crd_col_sum(const void* dy, int dtype, float* db, long long M, int N, int ld, crd_stream_t stream) {
  CRD_REQUIRE(ld % 8 == 0 && N <= ld);
  if (M == 0 || N == 0) return 0;
  const int Np = (N + 7) / 8 * 8;
  CRD_REQUIRE(Np <= ld && Np / 8 <= 256);
  ReduceLaunch r = plan_reduce(1, M, Np);
  const size_t smem = (size_t)r.block.x * r.block.y * 8 * sizeof(float);
  CRD_DISPATCH_1(dtype, T, crd_launch(col_sum_kernel<T>, dim3(dim3(r.grid.x)), dim3(r.block), smem, (cudaStream_t)stream, 
                               (const T*)dy, db, M, N, ld, r.ppb));
  CRD_LAUNCH_CHECK();
  return 0;
}
