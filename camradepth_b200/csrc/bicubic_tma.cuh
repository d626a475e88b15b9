// TMA-staged bicubic x2 upsample (forward and backward), bf16 NHWC with arbitrary pixel strides (the output of
// the forward / input of the backward is a channel slice of a decoder concat buffer).
//
// PyTorch upsample_bicubic2d, align_corners=False, scale 2, A=-0.75 (utils.py:241):
//   out[2y]   = sum_i c25[3-i] * in[clamp(y-2+i)]      out[2y+1] = sum_i c25[i] * in[clamp(y-1+i)]
// The kernel is separable.  As in dwconv_tma.cuh a persistent CTA owns a 64-channel tile and walks (sample,
// column strip, row range) work items; a producer warp streams row boxes through an mbarrier ring and every
// consumer thread marches down one image column for 4 channels:
//   forward : per input row 5 shared-memory loads -> two horizontal results (even / odd output column), kept in a
//             5-row register window; per input pixel 24 FMAs per channel produce the 2x2 outputs
//   backward: dx[s] = sum_{t<8} w_s[t] dy[2s-3+t] per axis (border clamping folded into w_s); per dy row one
//             8-tap horizontal result, kept in an 8-row register window, one dx row per two dy rows
// The register-staged kernels issue 16 (forward) / 64 (backward) global loads per output vector.
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "tma_util.cuh"
#include "dwconv_tma.cuh"
#include "../../include/camradepth_b200.h"

namespace {

constexpr int BC_COLS = 14;
constexpr int BC_CONSUMERS = 16 * BC_COLS;       // 7 warps
constexpr int BC_THREADS = BC_CONSUMERS + 32;
constexpr int BC_MAX_STAGES = 8;
constexpr int BC_MAX_TILES = 8;

struct BcParams {
  int B, H, W, C;                 // INPUT-resolution geometry (H x W); the upsampled side is 2H x 2W
  int ld_in, ld_out, accumulate;
  int TW, strips, rsplit, rows_per_split, nwork;
  int stages, stage_bytes;
  int ntiles, cta_begin[BC_MAX_TILES + 1];      // CTA ranges per 64-channel tile (a short last tile gets fewer)
  // A last tile of <= 8 channels (the 136-channel decoder tensors: 128 features + depth + padding) would keep 2 of
  // the 16 channel lanes busy.  It runs "narrow" instead: boxes of 8 channels (16-byte pixels in shared memory),
  // threads = 2 channel lanes x 112 image columns, its own strip geometry.
  int narrow, n_TW, n_strips, n_nwork, n_stage_bytes;
  bf16* out;
};

__device__ __forceinline__ float bc_weight(int o, int s, int n) {
  // weight with which output coordinate o reads input coordinate s (border clamping folded in)
  const float c25[4] = {-0.10546875f, 0.87890625f, 0.26171875f, -0.03515625f};
  if (o < 0 || o >= 2 * n) return 0.f;
  const int base = (o & 1) ? (o >> 1) - 1 : (o >> 1) - 2;
  float w = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int q = min(max(base + i, 0), n - 1);
    w += (q == s) ? ((o & 1) ? c25[i] : c25[3 - i]) : 0.f;
  }
  return w;
}

template <bool BWD>
__global__ void __launch_bounds__(BC_THREADS, 3)
bicubic_tma_kernel(const __grid_constant__ CUtensorMap m_in, const __grid_constant__ CUtensorMap m_nar,
                   const BcParams p) {
  CRD_PDL_ENTRY();
  constexpr int RH = BWD ? 8 : 5;                // rows per box = register-window depth
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t bar_full = base + p.stages * p.stage_bytes, bar_empty = bar_full + 8 * BC_MAX_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  int tile = 0;
  while (tile + 1 < p.ntiles && (int)blockIdx.x >= p.cta_begin[tile + 1]) tile++;
  const int c0 = tile * DW_CH;
  const int first = (int)blockIdx.x - p.cta_begin[tile], step = p.cta_begin[tile + 1] - p.cta_begin[tile];
  const bool nar = p.narrow && tile == p.ntiles - 1;
  const int TW = nar ? p.n_TW : p.TW, strips = nar ? p.n_strips : p.strips, nwork = nar ? p.n_nwork : p.nwork;
  const uint32_t pp = nar ? 16u : 128u;          // bytes per pixel in a shared-memory box row

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, BC_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // (the maps stay kernel parameters: a pointer selected at run time would make the compiler copy them to local memory)
    if (nar) asm volatile("prefetch.tensormap [%0];" ::"l"(&m_nar) : "memory");
    else asm volatile("prefetch.tensormap [%0];" ::"l"(&m_in) : "memory");
  }
  __syncthreads();

  const int per_b = strips * p.rsplit;
  if (warp == BC_CONSUMERS / 32) {
    if (lane == 0) {
      int it = 0;
      const uint32_t tx = (uint32_t)(nar ? p.n_stage_bytes : p.stage_bytes);
      for (int i = first; i < nwork; i += step) {
        const int b = i / per_b, rem = i - b * per_b;
        const int strip = rem / p.rsplit, rs = rem - strip * p.rsplit;
        const int h0 = rs * p.rows_per_split, h1 = min(p.H, h0 + p.rows_per_split);
        const int nrows = BWD ? 2 * (h1 - h0) + 6 : (h1 - h0) + 4;
        const int nb = (nrows + RH - 1) / RH;
        const int w0 = strip * TW;
        for (int k = 0; k < nb; k++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t bar = bar_full + 8 * s;
          mbar_expect_tx(bar, tx);
          const int cw = BWD ? 2 * w0 - 3 : w0 - 2, ch = BWD ? 2 * h0 - 3 + RH * k : h0 - 2 + RH * k;
          if (nar) tma_load_4d(base + s * p.stage_bytes, &m_nar, bar, c0, cw, ch, b);
          else tma_load_4d(base + s * p.stage_bytes, &m_in, bar, c0, cw, ch, b);
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int cg = nar ? (tid & 1) : (tid & 15), col = nar ? (tid >> 1) : (tid >> 4);
  const int c = c0 + cg * 4;
  const bool thr_ok = col < TW && c < p.C;
  const float k0 = -0.10546875f, k1 = 0.87890625f, k2 = 0.26171875f, k3 = -0.03515625f;   // phase .25 taps
  const uint32_t row_bytes = (uint32_t)(BWD ? 2 * TW + 6 : TW + 4) * pp;
  int it = 0;
  for (int i = first; i < nwork; i += step) {
    const int b = i / per_b, rem = i - b * per_b;
    const int strip = rem / p.rsplit, rs = rem - strip * p.rsplit;
    const int h0 = rs * p.rows_per_split, h1 = min(p.H, h0 + p.rows_per_split);
    const int nrows = BWD ? 2 * (h1 - h0) + 6 : (h1 - h0) + 4;
    const int nb = (nrows + RH - 1) / RH;
    const int w0 = strip * TW, wcol = w0 + col;
    const bool valid = thr_ok && wcol < p.W;

    if (!BWD) {
      // ---------------- forward: input rows h0-2 .. h1+1, five clamped column offsets per thread
      uint32_t coff[5];
#pragma unroll
      for (int d = 0; d < 5; d++) coff[d] = (uint32_t)(min(max(wcol + d - 2, 0), p.W - 1) - (w0 - 2)) * pp + cg * 8;
      float he[5][4], ho[5][4];
#pragma unroll
      for (int r = 0; r < 5; r++)
#pragma unroll
        for (int j = 0; j < 4; j++) he[r][j] = ho[r][j] = 0.f;
      bf16* ob = p.out + (((long long)b * 2 * p.H) * 2 * p.W + 2 * wcol) * p.ld_out + c;
      for (int k = 0; k < nb; k++, it++) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        if (valid) {
          const uint32_t slot = base + s * p.stage_bytes;
#pragma unroll
          for (int r = 0; r < RH; r++) {
            const int ii = RH * k + r;
            const int row = h0 - 2 + ii;                       // image row held by box row r
            if (row >= p.H) {
              // replicate border: same horizontal results as the last image row (the previous window slot)
#pragma unroll
              for (int j = 0; j < 4; j++) { he[r][j] = he[(r + 4) % 5][j]; ho[r][j] = ho[(r + 4) % 5][j]; }
            } else {
              const uint32_t ra = slot + (uint32_t)(row < 0 ? r - row : r) * row_bytes;   // rows < 0 read image row 0
              float v[5][4];
#pragma unroll
              for (int d = 0; d < 5; d++) lds8_unpack(ra + coff[d], v[d]);
#pragma unroll
              for (int j = 0; j < 4; j++) {
                he[r][j] = fmaf(k3, v[0][j], fmaf(k2, v[1][j], fmaf(k1, v[2][j], k0 * v[3][j])));
                ho[r][j] = fmaf(k0, v[1][j], fmaf(k1, v[2][j], fmaf(k2, v[3][j], k3 * v[4][j])));
              }
            }
            const int y = h0 + ii - 4;
            if (ii >= 4 && y < h1) {
              // window rows y-2 .. y+2 live in slots (r+1)%5 .. (r+5)%5
              float o[4];
              bf16* o0 = ob + (long long)(2 * y) * 2 * p.W * p.ld_out;
              bf16* o1 = o0 + (long long)2 * p.W * p.ld_out;
#pragma unroll
              for (int j = 0; j < 4; j++)
                o[j] = fmaf(k3, he[(r + 1) % 5][j], fmaf(k2, he[(r + 2) % 5][j], fmaf(k1, he[(r + 3) % 5][j], k0 * he[(r + 4) % 5][j])));
              stg8_bf16(o0, o);
#pragma unroll
              for (int j = 0; j < 4; j++)
                o[j] = fmaf(k3, ho[(r + 1) % 5][j], fmaf(k2, ho[(r + 2) % 5][j], fmaf(k1, ho[(r + 3) % 5][j], k0 * ho[(r + 4) % 5][j])));
              stg8_bf16(o0 + p.ld_out, o);
#pragma unroll
              for (int j = 0; j < 4; j++)
                o[j] = fmaf(k0, he[(r + 2) % 5][j], fmaf(k1, he[(r + 3) % 5][j], fmaf(k2, he[(r + 4) % 5][j], k3 * he[r][j])));
              stg8_bf16(o1, o);
#pragma unroll
              for (int j = 0; j < 4; j++)
                o[j] = fmaf(k0, ho[(r + 2) % 5][j], fmaf(k1, ho[(r + 3) % 5][j], fmaf(k2, ho[(r + 4) % 5][j], k3 * ho[r][j])));
              stg8_bf16(o1 + p.ld_out, o);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
      }
    } else {
      // ---------------- backward: dy rows 2*h0-3 .. 2*h1+2; dx row s needs dy rows 2s-3 .. 2s+4
      float wx[8];
#pragma unroll
      for (int t = 0; t < 8; t++) wx[t] = bc_weight(2 * wcol - 3 + t, wcol, p.W);
      float hw[8][4];
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int j = 0; j < 4; j++) hw[r][j] = 0.f;
      bf16* ob = p.out + (((long long)b * p.H) * p.W + wcol) * p.ld_out + c;
      for (int k = 0; k < nb; k++, it++) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        if (valid) {
          const uint32_t ra0 = base + s * p.stage_bytes + (uint32_t)(2 * col) * pp + cg * 8;
#pragma unroll
          for (int r = 0; r < RH; r++) {
            const int ii = RH * k + r;
            const uint32_t ra = ra0 + r * row_bytes;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < 8; t++) {
              float v[4];
              lds8_unpack(ra + t * pp, v);
#pragma unroll
              for (int j = 0; j < 4; j++) acc[j] = fmaf(wx[t], v[j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) hw[r][j] = acc[j];
            if (r & 1) {                                       // ii = 2m + 7 completes dx row h0 + m
              const int m = (ii - 7) >> 1;
              const int sy = h0 + m;
              if (ii >= 7 && sy < h1) {
                float wy[8];
                if (sy >= 2 && sy < p.H - 2) {
                  wy[0] = k3; wy[1] = k0; wy[2] = k2; wy[3] = k1; wy[4] = k1; wy[5] = k2; wy[6] = k0; wy[7] = k3;
                } else {
#pragma unroll
                  for (int t = 0; t < 8; t++) wy[t] = bc_weight(2 * sy - 3 + t, sy, p.H);
                }
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int t = 0; t < 8; t++)
#pragma unroll
                  for (int j = 0; j < 4; j++) o[j] = fmaf(wy[t], hw[(r + 1 + t) % 8][j], o[j]);
                bf16* op = ob + (long long)sy * p.W * p.ld_out;
                if (p.accumulate) {
                  const uint2 old = *reinterpret_cast<const uint2*>(op);
                  o[0] += __uint_as_float(old.x << 16); o[1] += __uint_as_float(old.x & 0xffff0000u);
                  o[2] += __uint_as_float(old.y << 16); o[3] += __uint_as_float(old.y & 0xffff0000u);
                }
                stg8_bf16(op, o);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
inline bool bc_tma_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("CAMRADEPTH_TMA_BICUBIC"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
inline bool bc_tma_eligible(int dtype, int B, int H, int W, int C, const void* in, int ld_in, const void* out,
                            int ld_out) {
  if (!bc_tma_enabled() || dtype != CRD_BF16 || C % 8 || C > DW_CH * BC_MAX_TILES) return false;
  if (H < 3 || W < 4 || B < 1 || (long long)H * W * C < (1LL << 13)) return false;      // per sample, see dwconv_tma.cuh
  if (((uintptr_t)in & 15) || ld_in % 8 || ((uintptr_t)out & 7) || ld_out % 4) return false;
  return true;
}

template <bool BWD>
inline int bc_tma_launch(const void* in, BcParams p, cudaStream_t st) {
  const int halo = BWD ? 6 : 4, wmul = BWD ? 2 : 1;
  int best = BC_COLS; long long best_cost = -1;
  for (int tw = BC_COLS; tw >= 7; tw--) {
    const long long strips = (p.W + tw - 1) / tw, cost = strips * (wmul * tw + halo);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = tw; }
  }
  p.TW = best;
  p.strips = (p.W + p.TW - 1) / p.TW;
  p.ntiles = (p.C + DW_CH - 1) / DW_CH;
  // CTAs per channel tile: proportional to the tile's work (a short last tile still pays the per-box latency)
  // resident CTAs per SM: 3 with a 72 KB ring each (80 registers), 2 with 104 KB (CAMRADEPTH_BICUBIC_OCC)
  static int occ = 0;
  if (!occ) { const char* e = getenv("CAMRADEPTH_BICUBIC_OCC"); occ = (e && e[0] == '2') ? 2 : 3; }
  const int ring_kb = occ == 3 ? 72 : 104;
  const int slots = occ * sm_count();
  float wsum = 0.f, wt[BC_MAX_TILES];
  static int nar_on = -1;
  if (nar_on < 0) { const char* e = getenv("CAMRADEPTH_BICUBIC_NARROW"); nar_on = (e && e[0] == '0') ? 0 : 1; }
  const int tail = p.C - (p.ntiles - 1) * DW_CH;
  p.narrow = (nar_on && p.ntiles > 1 && tail <= 8) ? 1 : 0;
  for (int t = 0; t < p.ntiles; t++) {
    const int valid = p.C - t * DW_CH < DW_CH ? p.C - t * DW_CH : DW_CH;
    wt[t] = valid >= 32 ? 1.f : 0.4f;
    if (p.narrow && t == p.ntiles - 1) wt[t] = 0.16f;       // 1/8 of a full tile's bytes and thread work
    wsum += wt[t];
  }
  const int per_full = (int)(slots / wsum) > 0 ? (int)(slots / wsum) : 1;
  p.rsplit = 1;
  while ((long long)p.B * p.strips * p.rsplit < 6LL * per_full && p.H / (p.rsplit + 1) >= 8) p.rsplit++;
  p.rows_per_split = (p.H + p.rsplit - 1) / p.rsplit;
  p.rsplit = (p.H + p.rows_per_split - 1) / p.rows_per_split;
  p.nwork = p.B * p.strips * p.rsplit;
  const int NAR_COLS = BC_CONSUMERS / 2;
  p.n_strips = (p.W + NAR_COLS - 1) / NAR_COLS;
  p.n_TW = (p.W + p.n_strips - 1) / p.n_strips;
  p.n_strips = (p.W + p.n_TW - 1) / p.n_TW;
  p.n_nwork = p.B * p.n_strips * p.rsplit;
  p.cta_begin[0] = 0;
  for (int t = 0; t < p.ntiles; t++) {
    int n = (int)(slots * wt[t] / wsum);
    const int nw = (p.narrow && t == p.ntiles - 1) ? p.n_nwork : p.nwork;
    if (n < 1) n = 1;
    if (n > nw) n = nw;
    p.cta_begin[t + 1] = p.cta_begin[t] + n;
  }
  const int RH = BWD ? 8 : 5;
  p.stage_bytes = (wmul * p.TW + halo) * RH * 128;
  p.n_stage_bytes = (wmul * p.n_TW + halo) * RH * 16;
  if (p.narrow && p.n_stage_bytes > p.stage_bytes) p.stage_bytes = (p.n_stage_bytes + 127) / 128 * 128;
  p.stages = (ring_kb * 1024) / p.stage_bytes;
  if (p.stages > BC_MAX_STAGES) p.stages = BC_MAX_STAGES;
  if (p.stages < 2) return -21;
  const int smem = p.stages * p.stage_bytes + 16 * BC_MAX_STAGES + 256;
  static unsigned long long attr = 0;
  if (int e = ensure_smem_attr(bicubic_tma_kernel<BWD>, 104 * 1024 + 16 * BC_MAX_STAGES + 256, attr)) return e;
  CUtensorMap m_in;
  const int Hi = BWD ? 2 * p.H : p.H, Wi = BWD ? 2 * p.W : p.W;
  cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)p.B};
  cuuint64_t str[3] = {(cuuint64_t)p.ld_in * 2, (cuuint64_t)Wi * p.ld_in * 2, (cuuint64_t)Hi * Wi * p.ld_in * 2};
  cuuint32_t box[4] = {DW_CH, (cuuint32_t)(wmul * p.TW + halo), (cuuint32_t)RH, 1};
  if (int e = make_map(&m_in, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
  CUtensorMap m_nar = m_in;
  if (p.narrow) {
    cuuint32_t boxn[4] = {8, (cuuint32_t)(wmul * p.n_TW + halo), (cuuint32_t)RH, 1};
    if (int e = make_map(&m_nar, in, 4, dims, str, boxn, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
  }
  crd_launch(bicubic_tma_kernel<BWD>, dim3(p.cta_begin[p.ntiles]), dim3(BC_THREADS), smem, st, m_in, m_nar, p);
  return 0;
}

}  // namespace
