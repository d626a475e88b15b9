// TMA-staged depthwise 3x3 (forward fused with the preceding GroupNorm apply; fused backward producing the
// input gradient and the weight / bias gradients).  bf16 NHWC [B][H][W][C] tensors.
//
// The register-staged kernels issue nine 16-byte global loads per output pixel (L1/L2 amplification 9x,
// 130-254 registers per thread) and reach 0.7 TB/s.  Here a persistent CTA owns one 64-channel tile and walks a
// list of (sample, column strip, row range) work items; a producer warp streams boxes {64 ch, TW+2 columns,
// 6 rows} (halo columns and rows come from the box coordinates, out-of-image taps are zero-filled by the TMA
// unit) through an mbarrier ring, and each consumer thread marches down ONE image column for 4 channels with a
// 3x3 register window: three 8-byte shared-memory loads per output instead of nine global loads.
//   forward : y = a * (sum_t w_t x_t) + (bias + sh * sum_{t inside the image} w_t)   (zero padding applies to
//             the NORMALISED tensor, so the constant term drops the taps that fall outside)
//   backward: dxn[q] = sum_t w_t dy[q - t];  dw_t += dy[q - t] * (a x[q] + sh);  db += dy[q]
// Weight-gradient partials stay in registers over all work items of the CTA: 640 atomics per CTA per launch.
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "tma_util.cuh"
#include "../../include/camradepth_b200.h"

namespace {

constexpr int DW_CH = 64;              // channels per CTA tile (128 B per pixel)
constexpr int DW_RH = 6;               // rows per TMA box (multiple of 3: the register window rotates statically)
constexpr int DW_COLS = 14;             // image columns per strip (upper bound)
constexpr int DW_CONSUMERS = 16 * DW_COLS;   // 16 channel groups x DW_COLS columns = 7 warps (+1 producer warp = 256 threads)
constexpr int DW_THREADS = DW_CONSUMERS + 32;
constexpr int DW_MAX_STAGES = 8;

struct DwParams {
  int B, H, W, C;
  int TW, strips, rsplit, rows_per_split, nwork;
  int stages, a_bytes, stage_bytes;
  const float* ab;       // [B][C][2]
  const float* w;        // [C][9]
  const float* bias;     // [C] or null (forward)
  bf16* out;             // y (forward) / dxn (backward)
  float* dw;             // [C][9] (backward)
  float* db;             // [C] or null (backward)
};

__device__ __forceinline__ void lds8_unpack(uint32_t addr, float (&v)[4]) {
  uint32_t lo, hi;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
  v[0] = __uint_as_float(lo << 16); v[1] = __uint_as_float(lo & 0xffff0000u);
  v[2] = __uint_as_float(hi << 16); v[3] = __uint_as_float(hi & 0xffff0000u);
}
// the same four channels as two register pairs (operands of FFMA2 / FMUL2)
__device__ __forceinline__ void lds8_unpack2(uint32_t addr, float2 (&v)[2]) {
  float t[4];
  lds8_unpack(addr, t);
  v[0] = make_float2(t[0], t[1]); v[1] = make_float2(t[2], t[3]);
}
__device__ __forceinline__ void stg8_bf16(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

template <bool BWD>
__global__ void __launch_bounds__(DW_THREADS, BWD ? 1 : 2)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap m_a, const __grid_constant__ CUtensorMap m_x, const DwParams p) {
  CRD_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + p.stages * p.stage_bytes, bar_empty = bar_full + 8 * DW_MAX_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = blockIdx.y * DW_CH;
  const int S = p.stages;

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, DW_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&m_a) : "memory");
    if (BWD) asm volatile("prefetch.tensormap [%0];" ::"l"(&m_x) : "memory");
  }
  __syncthreads();

  const int per_b = p.strips * p.rsplit;
  if (warp == DW_CONSUMERS / 32) {
    if (lane == 0) {
      int it = 0;
      const uint32_t tx = (uint32_t)p.stage_bytes;
      for (int i = blockIdx.x; i < p.nwork; i += gridDim.x) {
        const int b = i / per_b, rem = i - b * per_b;
        const int strip = rem / p.rsplit, rs = rem - strip * p.rsplit;
        const int h0 = rs * p.rows_per_split, h1 = min(p.H, h0 + p.rows_per_split);
        const int nb = (h1 - h0 + 2 + DW_RH - 1) / DW_RH;
        const int w0 = strip * p.TW;
        for (int k = 0; k < nb; k++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t bar = bar_full + 8 * s, slot = base + s * p.stage_bytes;
          mbar_expect_tx(bar, tx);
          tma_load_4d(slot, &m_a, bar, c0, w0 - 1, h0 - 1 + DW_RH * k, b);
          if (BWD) tma_load_4d(slot + p.a_bytes, &m_x, bar, c0, w0, h0 - 2 + DW_RH * k, b);
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int cg = tid & 15, col = tid >> 4;
  const int c = c0 + cg * 4;
  const bool col_ok = col < p.TW;
  // four channels = two register pairs: the taps run on FFMA2 (two channels per instruction, same per-lane arithmetic)
  float2 wt[9][2], bs[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    bs[j] = (!BWD && p.bias) ? make_float2(p.bias[c + 2 * j], p.bias[c + 2 * j + 1]) : make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; t++) wt[t][j] = make_float2(p.w[(c + 2 * j) * 9 + t], p.w[(c + 2 * j + 1) * 9 + t]);
  }
  float2 acc[BWD ? 10 : 1][2];
#pragma unroll
  for (int q = 0; q < (BWD ? 10 : 1); q++)
#pragma unroll
    for (int j = 0; j < 2; j++) acc[q][j] = make_float2(0.f, 0.f);

  const uint32_t pix_a = (uint32_t)((p.TW + 2) * 128), pix_x = (uint32_t)(p.TW * 128);
  int it = 0;
  for (int i = blockIdx.x; i < p.nwork; i += gridDim.x) {
    const int b = i / per_b, rem = i - b * per_b;
    const int strip = rem / p.rsplit, rs = rem - strip * p.rsplit;
    const int h0 = rs * p.rows_per_split, h1 = min(p.H, h0 + p.rows_per_split);
    const int nb = (h1 - h0 + 2 + DW_RH - 1) / DW_RH;
    const int wcol = strip * p.TW + col;
    const bool valid = col_ok && wcol < p.W;
    float2 a[2], sh[2], kfull[2], ktop[2], kbot[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const long long bc = (long long)b * p.C + c + 2 * j;
      a[j] = make_float2(p.ab[bc * 2], p.ab[bc * 2 + 2]);
      sh[j] = make_float2(p.ab[bc * 2 + 1], p.ab[bc * 2 + 3]);
      if (!BWD) {
        // constant term: bias + sh * (sum of the taps that land inside the image for this column)
        const bool l = wcol > 0, r = wcol < p.W - 1;
        float2 wr[3];
#pragma unroll
        for (int kh = 0; kh < 3; kh++) {
          const float2 wl = l ? wt[kh * 3][j] : make_float2(0.f, 0.f), wrr = r ? wt[kh * 3 + 2][j] : make_float2(0.f, 0.f);
          wr[kh] = make_float2(wl.x + wt[kh * 3 + 1][j].x + wrr.x, wl.y + wt[kh * 3 + 1][j].y + wrr.y);
        }
        kfull[j] = make_float2(fmaf(sh[j].x, wr[0].x + wr[1].x + wr[2].x, bs[j].x),
                               fmaf(sh[j].y, wr[0].y + wr[1].y + wr[2].y, bs[j].y));
        ktop[j] = make_float2(sh[j].x * wr[0].x, sh[j].y * wr[0].y);
        kbot[j] = make_float2(sh[j].x * wr[2].x, sh[j].y * wr[2].y);
      }
    }
    bf16* ob = p.out + (((long long)b * p.H) * p.W + wcol) * p.C + c;
    float2 win[3][3][2];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int j = 0; j < 2; j++) win[r][d][j] = make_float2(0.f, 0.f);

    for (int k = 0; k < nb; k++, it++) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      mbar_wait(bar_full + 8 * s, ph);
      if (col_ok) {
        const uint32_t sa = base + s * p.stage_bytes + (uint32_t)(col * 128 + cg * 8);
        const uint32_t sx = base + s * p.stage_bytes + p.a_bytes + (uint32_t)(col * 128 + cg * 8);
#pragma unroll
        for (int r = 0; r < DW_RH; r++) {
          // newest row goes to window slot r % 3 (DW_RH is a multiple of 3, so the slot is static)
#pragma unroll
          for (int d = 0; d < 3; d++) lds8_unpack2(sa + r * pix_a + d * 128, win[r % 3][d]);
          const int ii = DW_RH * k + r;
          const int h = h0 + ii - 2;
          if (ii >= 2 && h < h1 && valid) {
            const float2 (&top)[3][2] = win[(r + 1) % 3];
            const float2 (&mid)[3][2] = win[(r + 2) % 3];
            const float2 (&bot)[3][2] = win[r % 3];
            float2 o[2];
            if (!BWD) {
#pragma unroll
              for (int j = 0; j < 2; j++) {
                float2 t = make_float2(0.f, 0.f);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                  t = __ffma2_rn(wt[d][j], top[d][j], t);
                  t = __ffma2_rn(wt[3 + d][j], mid[d][j], t);
                  t = __ffma2_rn(wt[6 + d][j], bot[d][j], t);
                }
                float2 kk = kfull[j];
                if (h == 0) { kk.x -= ktop[j].x; kk.y -= ktop[j].y; }
                if (h == p.H - 1) { kk.x -= kbot[j].x; kk.y -= kbot[j].y; }
                o[j] = __ffma2_rn(a[j], t, kk);
              }
            } else {
              float2 xv[2];
              lds8_unpack2(sx + r * pix_x, xv);
#pragma unroll
              for (int j = 0; j < 2; j++) {
                const float2 xn = __ffma2_rn(a[j], xv[j], sh[j]);
                float2 t = make_float2(0.f, 0.f);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                  // window (row rr, column d) holds dy[q - tap] for tap (kh, kw) = (2 - rr, 2 - d)
                  t = __ffma2_rn(wt[6 + (2 - d)][j], top[d][j], t);
                  t = __ffma2_rn(wt[3 + (2 - d)][j], mid[d][j], t);
                  t = __ffma2_rn(wt[(2 - d)][j], bot[d][j], t);
                  acc[6 + (2 - d)][j] = __ffma2_rn(top[d][j], xn, acc[6 + (2 - d)][j]);
                  acc[3 + (2 - d)][j] = __ffma2_rn(mid[d][j], xn, acc[3 + (2 - d)][j]);
                  acc[(2 - d)][j] = __ffma2_rn(bot[d][j], xn, acc[(2 - d)][j]);
                }
                acc[9][j] = __fadd2_rn(acc[9][j], mid[1][j]);
                o[j] = t;
              }
            }
            {
              const float ov[4] = {o[0].x, o[0].y, o[1].x, o[1].y};
              stg8_bf16(ob + (long long)h * p.W * p.C, ov);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    }
  }

  if (BWD) {
    // all boxes are consumed: the ring now holds the per-column partials [10][DW_COLS columns][64 channels]
    asm volatile("bar.sync 1, %0;" ::"n"(DW_CONSUMERS) : "memory");
    float* red = reinterpret_cast<float*>(base_ptr);
#pragma unroll
    for (int q = 0; q < 10; q++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        red[(q * DW_COLS + col) * DW_CH + cg * 4 + 2 * j] = acc[q][j].x;
        red[(q * DW_COLS + col) * DW_CH + cg * 4 + 2 * j + 1] = acc[q][j].y;
      }
    asm volatile("bar.sync 1, %0;" ::"n"(DW_CONSUMERS) : "memory");
    for (int e = tid; e < 10 * DW_CH; e += DW_CONSUMERS) {
      const int q = e / DW_CH, ch = e - q * DW_CH;
      float s = 0.f;
#pragma unroll
      for (int cc = 0; cc < DW_COLS; cc++) s += red[(q * DW_COLS + cc) * DW_CH + ch];
      if (q < 9) atomicAdd(p.dw + (long long)(c0 + ch) * 9 + q, s);
      else if (p.db) atomicAdd(p.db + c0 + ch, s);
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
inline bool dw_tma_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("CAMRADEPTH_TMA_DWCONV"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
inline bool dw_tma_eligible(int dtype, int B, int H, int W, int C, const void* p0, const void* p1, const void* p2) {
  if (!dw_tma_enabled() || dtype != CRD_BF16 || C % DW_CH) return false;
  // size rule per SAMPLE (= the old per-tensor rule at batch 32): the kernel choice, hence the arithmetic order, must
  // not depend on the batch a sample sits in (deterministic mode asserts bitwise batch independence)
  if (H < 2 || W < 4 || B < 1 || (long long)H * W * C < (1LL << 14)) return false;
  if (((uintptr_t)p0 & 15) || ((uintptr_t)p1 & 15) || ((uintptr_t)p2 & 7)) return false;
  return true;
}

template <bool BWD>
inline int dw_tma_launch(const void* a_in, const void* x_in, DwParams p, cudaStream_t st) {
  // strip width: fewest loaded columns (strips * (TW + 2)); ties -> wider strips
  int best = DW_COLS; long long best_cost = -1;
  for (int tw = DW_COLS; tw >= 8; tw--) {
    const long long strips = (p.W + tw - 1) / tw, cost = strips * (tw + 2);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = tw; }
  }
  p.TW = best;
  p.strips = (p.W + p.TW - 1) / p.TW;
  const int ctiles = p.C / DW_CH;
  const int slots = (BWD ? 1 : 2) * sm_count();
  int per_tile = slots / ctiles;
  if (per_tile < 1) per_tile = 1;
  // split rows so that every CTA gets several work items (load balance) while row ranges stay >= 6 rows
  p.rsplit = 1;
  while ((long long)p.B * p.strips * p.rsplit < 8LL * per_tile && p.H / (p.rsplit + 1) >= 6) p.rsplit++;
  p.rows_per_split = (p.H + p.rsplit - 1) / p.rsplit;
  p.rsplit = (p.H + p.rows_per_split - 1) / p.rows_per_split;
  p.nwork = p.B * p.strips * p.rsplit;
  if (per_tile > p.nwork) per_tile = p.nwork;
  p.a_bytes = (p.TW + 2) * DW_RH * 128;
  p.stage_bytes = p.a_bytes + (BWD ? p.TW * DW_RH * 128 : 0);
  const int budget = BWD ? 200 * 1024 : 106 * 1024;
  p.stages = budget / p.stage_bytes;
  if (p.stages > DW_MAX_STAGES) p.stages = DW_MAX_STAGES;
  if (BWD && p.stages * p.stage_bytes < 10 * DW_COLS * DW_CH * 4) return -20;   // reduction scratch must fit the ring
  const int smem = p.stages * p.stage_bytes + 16 * DW_MAX_STAGES + 256;
  static unsigned long long attr = 0;
  if (int e = ensure_smem_attr(dwconv_tma_kernel<BWD>, 200 * 1024 + 16 * DW_MAX_STAGES + 256, attr)) return e;
  CUtensorMap m_a, m_x;
  cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  cuuint64_t str[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.W * p.C * 2, (cuuint64_t)p.H * p.W * p.C * 2};
  cuuint32_t box_a[4] = {DW_CH, (cuuint32_t)(p.TW + 2), DW_RH, 1};
  if (int e = make_map(&m_a, a_in, 4, dims, str, box_a, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
  m_x = m_a;
  if (BWD) {
    cuuint32_t box_x[4] = {DW_CH, (cuuint32_t)p.TW, DW_RH, 1};
    if (int e = make_map(&m_x, x_in, 4, dims, str, box_x, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
  }
  crd_launch(dwconv_tma_kernel<BWD>, dim3(dim3(per_tile, ctiles)), dim3(DW_THREADS), smem, st, m_a, m_x, p);
  return 0;
}

}  // namespace
