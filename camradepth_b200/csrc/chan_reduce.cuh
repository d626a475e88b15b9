// Per-(sample, channel) reductions over pixels of an NHWC tensor (shared by several kernels).
#pragma once
#include "common.cuh"

namespace {

// ---- generic per-(b,c) two-quantity reduction over pixels ---------------------------------
// blockDim = (cvec, rows); each thread owns 8 channels and strides over the block's pixel range.
template <typename F>
__device__ __forceinline__ void chan_reduce2(F f, float* out /*[B][C][2]*/, int B, long long N, int C,
                                             long long pix_per_block) {
  extern __shared__ float red[];
  const int cv = threadIdx.x, ry = threadIdx.y, rows = blockDim.y, cvec = blockDim.x;
  const int b = blockIdx.y;
  long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > N) p1 = N;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { s0[j] = 0.f; s1[j] = 0.f; }
#pragma unroll 4
  for (long long p = p0 + ry; p < p1; p += rows) f(b, p, cv * 8, s0, s1);
  // reduce over rows through shared memory
  float* r0 = red;                       // [rows][cvec*8]
  float* r1 = red + (size_t)rows * cvec * 8;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    r0[(ry * cvec + cv) * 8 + j] = s0[j];
    r1[(ry * cvec + cv) * 8 + j] = s1[j];
  }
  __syncthreads();
  const int t = ry * cvec + cv, nt = rows * cvec;
  for (int c = t; c < C; c += nt) {
    float a0 = 0.f, a1 = 0.f;
    for (int r = 0; r < rows; r++) { a0 += r0[r * cvec * 8 + c]; a1 += r1[r * cvec * 8 + c]; }
    atomicAdd(out + ((long long)b * C + c) * 2 + 0, a0);
    atomicAdd(out + ((long long)b * C + c) * 2 + 1, a1);
  }
}

// Tail of every per-(b,c) reduction: rows -> shared memory -> one atomic per (block, channel, quantity).
__device__ __forceinline__ void chan_reduce_finish(float (&s0)[8], float (&s1)[8], float* out, int b, int C) {
  extern __shared__ float red[];
  const int cv = threadIdx.x, ry = threadIdx.y, rows = blockDim.y, cvec = blockDim.x;
  float* r0 = red;
  float* r1 = red + (size_t)rows * cvec * 8;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    r0[(ry * cvec + cv) * 8 + j] = s0[j];
    r1[(ry * cvec + cv) * 8 + j] = s1[j];
  }
  __syncthreads();
  const int t = ry * cvec + cv, nt = rows * cvec;
  for (int c = t; c < C; c += nt) {
    float a0 = 0.f, a1 = 0.f;
    for (int r = 0; r < rows; r++) { a0 += r0[r * cvec * 8 + c]; a1 += r1[r * cvec * 8 + c]; }
    atomicAdd(out + ((long long)b * C + c) * 2 + 0, a0);
    atomicAdd(out + ((long long)b * C + c) * 2 + 1, a1);
  }
}

struct ReduceLaunch { dim3 grid, block; size_t smem; long long ppb; };
inline ReduceLaunch plan_reduce(int B, long long N, int C) {
  ReduceLaunch r;
  int cvec = C / 8;
  int rows = 256 / cvec; if (rows < 1) rows = 1; if (rows > 32) rows = 32;
  // total blocks just UNDER 148 * 12 = a whole number of waves for 2, 3, 4 or 6 resident blocks per SM
  // (896 blocks at 4 blocks/SM was 1.51 waves: a quarter of the time spent in a half-empty tail)
  long long blocks_per_b = (148LL * 12) / B;
  if (blocks_per_b < 1) blocks_per_b = 1;
  long long min_ppb = rows * 4;
  long long ppb = (N + blocks_per_b - 1) / blocks_per_b;
  if (ppb < min_ppb) ppb = min_ppb;
  r.ppb = ppb;
  r.grid = dim3((unsigned)((N + ppb - 1) / ppb), B);
  r.block = dim3(cvec, rows);
  r.smem = (size_t)rows * cvec * 8 * 2 * sizeof(float);
  return r;
}


// streaming (non-reducing) kernels with the same (cvec, rows) thread mapping: more, smaller blocks
inline ReduceLaunch plan_stream(int B, long long N, int C) {
  ReduceLaunch r;
  int cvec = C / 8;
  int rows = 256 / cvec; if (rows < 1) rows = 1; if (rows > 32) rows = 32;
  long long blocks_per_b = (148LL * 12) / B;       // whole waves for 2, 3, 4 or 6 resident blocks per SM
  if (blocks_per_b < 1) blocks_per_b = 1;
  long long min_ppb = rows * 8;
  long long ppb = (N + blocks_per_b - 1) / blocks_per_b;
  if (ppb < min_ppb) ppb = min_ppb;
  r.ppb = ppb;
  r.grid = dim3((unsigned)((N + ppb - 1) / ppb), B);
  r.block = dim3(cvec, rows);
  r.smem = 0;
  return r;
}

inline int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  long long cap = 148LL * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace
