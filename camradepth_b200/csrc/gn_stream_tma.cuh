// TMA-staged variants of the GroupNorm streaming kernels (bf16 NHWC tensors).
//
// The register-staged kernels in norm_act.cu keep their in-flight loads in registers, so bytes in flight are
// capped by occupancy (122-175 registers per thread -> 32-64 KB per SM, measured 2.2-3.9 TB/s).  Here a
// producer warp streams {CW channels x PH pixels} boxes of each input through a ring of 16-KiB shared-memory
// slots with cp.async.bulk.tensor + mbarriers (96 KiB in flight per CTA, two CTAs per SM), and 224 consumer
// threads read their 16-byte pieces from shared memory, so the HBM queue depth no longer depends on registers.
//   grid  = (pixel splits, C / CW, B);  one CTA owns a (sample, channel tile, pixel range) and keeps the
//           per-(b,c) constants of its 8 channels per thread in registers, like the register-staged kernels
//   block = 7 consumer warps laid out (cvec = CW/8, rows) + 1 producer warp
// Out-of-range pixels of the last tile are zero-filled by the TMA unit and masked by the consumers.
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "tma_util.cuh"
#include "../../include/camradepth_b200.h"

namespace {

constexpr int ST_SLOT = 16384;         // bytes per tile slot
constexpr int ST_SLOTS = 6;            // 96 KiB ring
constexpr int ST_CONSUMERS = 224;          // 7 warps + 1 producer warp = 256 threads: 128 registers at 2 CTAs/SM
constexpr int ST_THREADS = ST_CONSUMERS + 32;
constexpr int ST_R = 4;                // pixel rows per consumer thread per tile (independent chains per thread)
constexpr int ST_SMEM = ST_SLOTS * ST_SLOT + 256 + 128;

enum { ST_STATS = 0, ST_AFFINE = 1, ST_BWD_REDUCE = 2, ST_BWD_APPLY = 3 };

struct StParams {
  int B, C, CW, cvec, rows, PH;
  long long N, ntiles, tiles_per_cta;
  int act, ldo;
  const float* ab;        // [B][C][2] affine (a, b)
  const float* post;      // [B][C] or null
  const float* addbc;     // [B][C] or null
  const float* coef;      // [B][C][3] or null
  float* red;             // [B][C][2] reduction output or null
  bf16* out;              // y / dz / dx or null
};

__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

__device__ __forceinline__ float st_act_bwd(float z, int act) {
  if (act == CRD_ACT_GELU) return gelu_grad_f(z);
  if (act == CRD_ACT_SIGMOID) { float s = sigmoid_f(z); return s * (1.f - s); }
  return 1.f;
}
__device__ __forceinline__ float st_act_fwd(float z, int act) {
  if (act == CRD_ACT_GELU) return gelu_f(z);
  if (act == CRD_ACT_SIGMOID) return sigmoid_f(z);
  return z;
}

template <int MODE>
__global__ void __launch_bounds__(ST_THREADS, 2)
gn_stream_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, const StParams p) {
  CRD_PDL_ENTRY();
  constexpr int NIN = (MODE == ST_STATS || MODE == ST_AFFINE) ? 1 : 2;
  constexpr int S = ST_SLOTS / NIN;
  constexpr bool REDUCES = (MODE == ST_STATS || MODE == ST_BWD_REDUCE);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + ST_SLOTS * ST_SLOT, bar_empty = bar_full + 8 * ST_SLOTS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ct = blockIdx.y, b = blockIdx.z;
  const long long t0 = (long long)blockIdx.x * p.tiles_per_cta;
  const long long t1 = min(p.ntiles, t0 + p.tiles_per_cta);
  const int nt = (int)max(0LL, t1 - t0);
  const uint32_t tile_bytes = (uint32_t)(p.PH * p.CW * 2);

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, ST_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&m0) : "memory");
    if (NIN == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&m1) : "memory");
  }
  __syncthreads();

  if (warp == ST_CONSUMERS / 32) {
    if (lane == 0) {
      int pix = (int)(t0 * p.PH);
      for (int it = 0; it < nt; it++) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t bar = bar_full + 8 * s;
        mbar_expect_tx(bar, NIN * tile_bytes);
        tma_load_3d(base + (uint32_t)(s * NIN) * ST_SLOT, &m0, bar, ct * p.CW, pix, b);
        if (NIN == 2) tma_load_3d(base + (uint32_t)(s * NIN + 1) * ST_SLOT, &m1, bar, ct * p.CW, pix, b);
        pix += p.PH;
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int ty = tid / p.cvec, tx = tid - ty * p.cvec;
  const bool active = ty < p.rows;
  const int c = ct * p.CW + tx * 8;
  float a[8], sh[8], k1[8], k0[8], cB[8], cC[8];
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = sh[j] = k1[j] = k0[j] = cB[j] = cC[j] = 0.f;
    s0[j] = s1[j] = 0.f;
  }
  if (active && MODE != ST_STATS) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const long long bc = (long long)b * p.C + c + j;
      a[j] = p.ab[bc * 2];
      sh[j] = p.ab[bc * 2 + 1];
      const float ps = p.post ? p.post[bc] : 1.f;
      const float ad = p.addbc ? p.addbc[bc] : 0.f;
      // dz = (g + ad) * ps * act'(z) = (k1 * g + k0) * act'(z)
      k1[j] = ps;
      k0[j] = ad * ps;
      if (MODE == ST_BWD_APPLY) {
        const float cA = p.coef[bc * 3];
        k1[j] *= cA; k0[j] *= cA;                 // dx = cA*dz + cB*x + cC
        cB[j] = p.coef[bc * 3 + 1];
        cC[j] = p.coef[bc * 3 + 2];
      }
    }
  }
  const uint32_t toff = (uint32_t)tid * 16u;                   // (ty * CW + tx * 8) * 2 bytes: linear in tid
  const uint32_t rstep = (uint32_t)(p.rows * p.CW * 2);
  bf16* outb = p.out ? p.out + (long long)b * p.N * p.ldo + c : nullptr;
  long long pix0 = t0 * p.PH;
  for (int it = 0; it < nt; it++, pix0 += p.PH) {
    const int s = it % S;
    const uint32_t ph = (it / S) & 1;
    mbar_wait(bar_full + 8 * s, ph);
    if (active) {
      const uint32_t sl = base + (uint32_t)(s * NIN) * ST_SLOT + toff;
      uint4 r0[ST_R], r1[ST_R];
#pragma unroll
      for (int r = 0; r < ST_R; r++) {
        r0[r] = lds16(sl + r * rstep);
        if (NIN == 2) r1[r] = lds16(sl + ST_SLOT + r * rstep);
      }
#pragma unroll
      for (int r = 0; r < ST_R; r++) {
        const long long q = pix0 + ty + r * p.rows;
        if (q >= p.N) continue;
        float g[8], v[8];
        unpack8(r0[r], g);
        if (MODE == ST_STATS) {
#pragma unroll
          for (int j = 0; j < 8; j++) { s0[j] += g[j]; s1[j] = fmaf(g[j], g[j], s1[j]); }
        } else if (MODE == ST_AFFINE) {
          if (p.act == CRD_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {       // two channels per FFMA2 / FMUL2
              const float2 z = __ffma2_rn(make_float2(a[j], a[j + 1]), make_float2(g[j], g[j + 1]),
                                          make_float2(sh[j], sh[j + 1]));
              const float2 y = __fmul2_rn(gelu2(z), make_float2(k1[j], k1[j + 1]));
              g[j] = y.x; g[j + 1] = y.y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] = st_act_fwd(fmaf(a[j], g[j], sh[j]), p.act) * k1[j];
          }
          store8(outb + q * p.ldo, g);
        } else {
          unpack8(r1[r], v);
          if (p.act == CRD_ACT_GELU) {
            // the activation derivative two channels at a time; g then holds dz and the loop below runs with it
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const float2 dz0 = __ffma2_rn(make_float2(k1[j], k1[j + 1]), make_float2(g[j], g[j + 1]),
                                            make_float2(k0[j], k0[j + 1]));
              const float2 z = __ffma2_rn(make_float2(a[j], a[j + 1]), make_float2(v[j], v[j + 1]),
                                          make_float2(sh[j], sh[j + 1]));
              const float2 dz = __fmul2_rn(dz0, gelu_grad2(z));
              g[j] = dz.x; g[j + 1] = dz.y;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; j++) {
            float dz = p.act == CRD_ACT_GELU ? g[j] : fmaf(k1[j], g[j], k0[j]);
            if (p.act != CRD_ACT_NONE && p.act != CRD_ACT_GELU) dz *= st_act_bwd(fmaf(a[j], v[j], sh[j]), p.act);
            if (MODE == ST_BWD_REDUCE) {
              s0[j] += dz;
              s1[j] = fmaf(dz, v[j], s1[j]);
              g[j] = dz;
            } else {
              g[j] = dz + fmaf(cB[j], v[j], cC[j]);
            }
          }
          if (outb) store8(outb + q * p.ldo, g);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
  }

  if (REDUCES) {
    // every TMA write has been consumed: the ring is free to hold the per-row partial sums [rows][CW][2]
    asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");
    float* r0 = reinterpret_cast<float*>(base_ptr);
    float* r1 = r0 + p.rows * p.CW;
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        r0[ty * p.CW + tx * 8 + j] = s0[j];
        r1[ty * p.CW + tx * 8 + j] = s1[j];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");
    for (int cc = tid; cc < p.CW; cc += ST_CONSUMERS) {
      float a0 = 0.f, a1 = 0.f;
      for (int r = 0; r < p.rows; r++) { a0 += r0[r * p.CW + cc]; a1 += r1[r * p.CW + cc]; }
      float* o = p.red + ((long long)b * p.C + ct * p.CW + cc) * 2;
      atomicAdd(o, a0);
      atomicAdd(o + 1, a1);
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
inline bool st_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("CAMRADEPTH_TMA_STREAM"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

// channel-tile width: whole rows up to 160 channels, else 128- or 64-channel tiles
inline int st_cw(int C) {
  if (C < 32) return 0;
  if (C <= 160) return C;
  if (C % 128 == 0) return 128;
  if (C % 64 == 0) return 64;
  return 0;
}

inline bool st_eligible(int B, long long N, int C, const void* p0, int ld0, const void* p1, int ld1) {
  if (!st_enabled() || st_cw(C) == 0 || C % 8) return false;
  if ((long long)B * N * C < (1LL << 20) || N < 64 || N > 0x7fffffffLL / 2) return false;
  if (((uintptr_t)p0 & 15) || ld0 % 8) return false;
  if (p1 && (((uintptr_t)p1 & 15) || ld1 % 8)) return false;
  return true;
}

struct StLaunch { StParams p; dim3 grid; };

inline StLaunch st_plan(int B, long long N, int C) {
  StLaunch L;
  StParams& p = L.p;
  p.B = B; p.C = C; p.N = N;
  p.CW = st_cw(C);
  p.cvec = p.CW / 8;
  p.rows = ST_CONSUMERS / p.cvec;
  p.PH = ST_R * p.rows;
  p.ntiles = (N + p.PH - 1) / p.PH;
  const long long cols = (long long)B * (C / p.CW);
  // CTAs = cols * splits should fill whole waves of 2 CTAs x 148 SMs; every CTA gets >= 4 tiles
  const long long slots = 2LL * sm_count();
  long long best_s = 1; double best_eff = -1.0;
  const long long max_s = p.ntiles / 4 > 0 ? p.ntiles / 4 : 1;
  for (int waves = 1; waves <= 4; waves++) {
    long long s = waves * slots / cols;
    if (s < 1) s = 1;
    if (s > max_s) s = max_s;
    const long long tpc = (p.ntiles + s - 1) / s;
    const long long ctas = cols * ((p.ntiles + tpc - 1) / tpc);
    const long long w = (ctas + slots - 1) / slots;
    // balance: useful tile slots / (waves * slots * tiles per CTA)
    const double eff = (double)(cols * p.ntiles) / (double)(w * slots * tpc);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_s = s; }
  }
  p.tiles_per_cta = (p.ntiles + best_s - 1) / best_s;
  L.grid = dim3((unsigned)((p.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta), C / p.CW, B);
  p.act = 0; p.ldo = 0;
  p.ab = p.post = p.addbc = p.coef = nullptr; p.red = nullptr; p.out = nullptr;
  return L;
}

inline int st_map(CUtensorMap* m, const void* ptr, int B, long long N, int C, int ld, const StParams& p) {
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)N * ld * 2};
  cuuint32_t box[3] = {(cuuint32_t)p.CW, (cuuint32_t)p.PH, 1};
  return make_map(m, ptr, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

template <int MODE>
inline int st_launch(const StLaunch& L, const CUtensorMap& m0, const CUtensorMap& m1, cudaStream_t s) {
  static unsigned long long attr = 0;
  if (int e = ensure_smem_attr(gn_stream_kernel<MODE>, ST_SMEM, attr)) return e;
  crd_launch(gn_stream_kernel<MODE>, dim3(L.grid), dim3(ST_THREADS), ST_SMEM, s, m0, m1, L.p);
  return 0;
}

}  // namespace
