// Generic implicit-GEMM convolution on CUDA cores (fp32 accumulate; fp32 or bf16 operands).
//
// This is the any-shape path: strided / transposed / tiny-channel convolutions, and every
// contraction in the fp32-exact parity mode.  The hot stride-1 3x3 and 1x1 contractions of
// the bf16 training path go through the tcgen05 kernels in igemm_tc.cu instead.
//
// GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin with
// k = (kh*KW + kw)*Cin + c, so an 8-wide k-chunk is 8 contiguous NHWC channels of one tap.
#include "common.cuh"
#include "../../include/camradepth_b200.h"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int APAD = 4;

struct PixCoord { int b, oh, ow; bool ok; };

__device__ __forceinline__ PixCoord decode_pixel(long long m, long long M, int Ho, int Wo) {
  PixCoord p;
  p.ok = m < M;
  long long mm = p.ok ? m : 0;
  int hw = Ho * Wo;
  p.b = (int)(mm / hw);
  int r = (int)(mm - (long long)p.b * hw);
  p.oh = r / Wo;
  p.ow = r - p.oh * Wo;
  return p;
}

// source pixel offset (in pixels) for output pixel p and tap (kh,kw); returns -1 if out of range
__device__ __forceinline__ long long src_pixel(const crd_conv_desc& d, const PixCoord& p, int kh, int kw) {
  int ih, iw;
  if (!d.transposed) {
    ih = p.oh * d.stride - d.pad + kh;
    iw = p.ow * d.stride - d.pad + kw;
  } else {
    int th = p.oh + d.pad - kh, tw = p.ow + d.pad - kw;
    if (th < 0 || tw < 0) return -1;
    if (d.stride > 1) {
      if ((th % d.stride) | (tw % d.stride)) return -1;
      th /= d.stride; tw /= d.stride;
    }
    ih = th; iw = tw;
  }
  if (ih < 0 || iw < 0 || ih >= d.H || iw >= d.W) return -1;
  return ((long long)p.b * d.H + ih) * d.W + iw;
}

template <typename TO>
__device__ __forceinline__ void store_out4(TO* y, const float (&v)[4], bool accumulate) {
  if (accumulate) {
#pragma unroll
    for (int j = 0; j < 4; j++) y[j] = from_f<TO>(v[j] + to_f(y[j]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) y[j] = from_f<TO>(v[j]);
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(NT) conv_fwd_kernel(crd_conv_desc d, const TI* __restrict__ x,
                                                      const TI* __restrict__ w, const float* __restrict__ bias,
                                                      TO* __restrict__ y) {
  CRD_PDL_ENTRY();
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + APAD];
  const int tid = threadIdx.x;
  const long long M = (long long)d.B * d.Ho * d.Wo;
  const int K = d.KH * d.KW * d.Cin;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // loader roles
  const int a_row = tid >> 1, a_kc = tid & 1;
  const PixCoord ap = decode_pixel(m0 + a_row, M, d.Ho, d.Wo);
  const int b_row = (tid & 127) >> 1, b_kc = tid & 1;
  const bool b_loader = tid < 128;
  const bool b_ok = (n0 + b_row) < d.Cout;

  float a_reg[8], b_reg[8];
  auto gload = [&](int k0) {
    // A chunk
    int k = k0 + a_kc * 8;
#pragma unroll
    for (int j = 0; j < 8; j++) a_reg[j] = 0.f;
    if (ap.ok && k < K) {
      int tap = k / d.Cin, c = k - tap * d.Cin;
      int kh = tap / d.KW, kw = tap - kh * d.KW;
      long long sp = src_pixel(d, ap, kh, kw);
      if (sp >= 0) load8(x + sp * d.ldx + c, a_reg);
    }
    if (b_loader) {
      int kb = k0 + b_kc * 8;
#pragma unroll
      for (int j = 0; j < 8; j++) b_reg[j] = 0.f;
      if (b_ok && kb < K) load8(w + (long long)(n0 + b_row) * K + kb, b_reg);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; j++) As[buf][a_kc * 8 + j][a_row] = a_reg[j];
    if (b_loader) {
#pragma unroll
      for (int j = 0; j < 8; j++) Bs[buf][b_kc * 8 + j][b_row] = b_reg[j];
    }
  };

  const int tm = tid >> 4, tn = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int it = 0; it < nk; it++) {
    const int buf = it & 1;
    if (it + 1 < nk) gload((it + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tm * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][tm * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  // epilogue
  const int nb = n0 + tn * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int j = 0; j < 4; j++) if (nb + j < d.Cout) bv[j] = bias[nb + j];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const long long m = m0 + tm * 8 + i;
    if (m >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      v[j] = acc[i][j] + bv[j];
      if (d.act == CRD_ACT_SIGMOID) v[j] = sigmoid_f(v[j]);
    }
    if (d.out_nchw) {
      const PixCoord p = decode_pixel(m, M, d.Ho, d.Wo);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (nb + j < d.Cout) {
          long long o = (((long long)p.b * d.Cout + nb + j) * d.Ho + p.oh) * d.Wo + p.ow;
          y[o] = from_f<TO>(d.accumulate ? v[j] + to_f(y[o]) : v[j]);
        }
      }
    } else {
      TO* yp = y + m * d.ldy + nb;
      if (nb + 3 < d.Cout) {
        store_out4<TO>(yp, v, d.accumulate != 0);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (nb + j < d.Cout) yp[j] = from_f<TO>(d.accumulate ? v[j] + to_f(yp[j]) : v[j]);
      }
    }
  }
}

// ------------------------------------------------------------------ weight gradient
// dw[n][k] += sum_p dy[p][n] * xcol[p][k] ; tile 128 (n) x 64 (k), pixels reduced in chunks of 16,
// split over blockIdx.z with fp32 atomics.
template <typename TI, typename TD>
__global__ void __launch_bounds__(NT) conv_wgrad_kernel(crd_conv_desc d, const TI* __restrict__ x,
                                                        const TD* __restrict__ dy, float* __restrict__ dw,
                                                        long long m_per_split) {
  CRD_PDL_ENTRY();
  __shared__ __align__(16) float Ds[2][BK][BM + APAD];   // [pixel][n]
  __shared__ __align__(16) float Xs[2][BK][BN + APAD];   // [pixel][k]
  const int tid = threadIdx.x;
  const long long M = (long long)d.B * d.Ho * d.Wo;
  const int K = d.KH * d.KW * d.Cin;
  const int k0 = blockIdx.x * BN;
  const int n0 = blockIdx.y * BM;
  const long long m_begin = (long long)blockIdx.z * m_per_split;
  long long m_end = m_begin + m_per_split;
  if (m_end > M) m_end = M;

  // dy loader: 16 pixels x 128 n = 256 chunks
  const int d_px = tid >> 4, d_nc = tid & 15;
  const bool d_ok_n = (n0 + d_nc * 8) < d.Cout;     // Cout padded to 8 in the dy buffer (ldy >= )
  // x loader: 16 pixels x 64 k = 128 chunks
  const bool x_loader = tid < 128;
  const int x_px = (tid & 127) >> 3, x_kc = tid & 7;
  const int xk = k0 + x_kc * 8;
  const bool x_ok_k = xk < K;
  int x_kh = 0, x_kw = 0, x_c = 0;
  if (x_ok_k) { int tap = xk / d.Cin; x_c = xk - tap * d.Cin; x_kh = tap / d.KW; x_kw = tap - x_kh * d.KW; }

  float d_reg[8], x_reg[8];
  auto gload = [&](long long mb) {
#pragma unroll
    for (int j = 0; j < 8; j++) d_reg[j] = 0.f;
    long long m = mb + d_px;
    if (d_ok_n && m < m_end) load8(dy + m * d.ldy + n0 + d_nc * 8, d_reg);
    if (x_loader) {
#pragma unroll
      for (int j = 0; j < 8; j++) x_reg[j] = 0.f;
      long long mx = mb + x_px;
      if (x_ok_k && mx < m_end) {
        PixCoord p = decode_pixel(mx, M, d.Ho, d.Wo);
        long long sp = src_pixel(d, p, x_kh, x_kw);
        if (sp >= 0) load8(x + sp * d.ldx + x_c, x_reg);
      }
    }
  };
  auto sstore = [&](int buf) {
    *reinterpret_cast<float4*>(&Ds[buf][d_px][d_nc * 8]) = make_float4(d_reg[0], d_reg[1], d_reg[2], d_reg[3]);
    *reinterpret_cast<float4*>(&Ds[buf][d_px][d_nc * 8 + 4]) = make_float4(d_reg[4], d_reg[5], d_reg[6], d_reg[7]);
    if (x_loader) {
      *reinterpret_cast<float4*>(&Xs[buf][x_px][x_kc * 8]) = make_float4(x_reg[0], x_reg[1], x_reg[2], x_reg[3]);
      *reinterpret_cast<float4*>(&Xs[buf][x_px][x_kc * 8 + 4]) = make_float4(x_reg[4], x_reg[5], x_reg[6], x_reg[7]);
    }
  };

  const int tm = tid >> 4, tn = tid & 15;     // tm: n rows (8), tn: k cols (4)
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  if (m_begin >= m_end) return;
  const int nit = (int)((m_end - m_begin + BK - 1) / BK);
  gload(m_begin);
  sstore(0);
  __syncthreads();
  for (int it = 0; it < nit; it++) {
    const int buf = it & 1;
    if (it + 1 < nit) gload(m_begin + (long long)(it + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float4 a0 = *reinterpret_cast<const float4*>(&Ds[buf][kk][tm * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&Ds[buf][kk][tm * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Xs[buf][kk][tn * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < nit) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int n = n0 + tm * 8 + i;
    if (n >= d.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int k = k0 + tn * 4 + j;
      if (k < K) atomicAdd(dw + (long long)n * K + k, acc[i][j]);
    }
  }
}

}  // namespace

extern "C" int crd_conv_fwd(const crd_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                            crd_stream_t stream) {
  CRD_REQUIRE(d && x && w && y);
  CRD_REQUIRE(d->Cin % 8 == 0 && d->ldx % 8 == 0);
  CRD_REQUIRE(d->out_nchw || d->ldy % 4 == 0);
  CRD_REQUIRE(d->stride >= 1 && d->w_tap_stride == 0 && d->w_koff == 0);
  const long long M = (long long)d->B * d->Ho * d->Wo;
  if (M == 0) return 0;
  dim3 grid(crd_div_up(M, BM), crd_div_up(d->Cout, BN));
  cudaStream_t s = (cudaStream_t)stream;
  if (d->in_dtype == CRD_F32 && d->out_dtype == CRD_F32)
    crd_launch(conv_fwd_kernel<float, float>, dim3(grid), dim3(NT), 0, s, *d, (const float*)x, (const float*)w, bias, (float*)y);
  else if (d->in_dtype == CRD_BF16 && d->out_dtype == CRD_BF16)
    crd_launch(conv_fwd_kernel<bf16, bf16>, dim3(grid), dim3(NT), 0, s, *d, (const bf16*)x, (const bf16*)w, bias, (bf16*)y);
  else if (d->in_dtype == CRD_BF16 && d->out_dtype == CRD_F32)
    crd_launch(conv_fwd_kernel<bf16, float>, dim3(grid), dim3(NT), 0, s, *d, (const bf16*)x, (const bf16*)w, bias, (float*)y);
  else
    return -2;
  CRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int crd_conv_wgrad(const crd_conv_desc* d, const void* x, const void* dy, float* dw,
                              crd_stream_t stream) {
  CRD_REQUIRE(d && x && dy && dw);
  CRD_REQUIRE(d->Cin % 8 == 0 && d->ldx % 8 == 0 && d->ldy % 8 == 0 && !d->transposed);
  const long long M = (long long)d->B * d->Ho * d->Wo;
  if (M == 0) return 0;
  const int K = d->KH * d->KW * d->Cin;
  const int gx = crd_div_up(K, BN), gy = crd_div_up(d->Cout, BM);
  // enough blocks to fill 148 SMs a few times over; each split handles a multiple of BK pixels
  long long want = (148LL * 4 + (long long)gx * gy - 1) / ((long long)gx * gy);
  long long max_splits = (M + 511) / 512;
  long long splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
  long long mps = ((M + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (M + mps - 1) / mps;
  dim3 grid(gx, gy, (unsigned)splits);
  cudaStream_t s = (cudaStream_t)stream;
  if (d->in_dtype == CRD_F32 && d->out_dtype == CRD_F32)
    crd_launch(conv_wgrad_kernel<float, float>, dim3(grid), dim3(NT), 0, s, *d, (const float*)x, (const float*)dy, dw, mps);
  else if (d->in_dtype == CRD_BF16 && d->out_dtype == CRD_BF16)
    crd_launch(conv_wgrad_kernel<bf16, bf16>, dim3(grid), dim3(NT), 0, s, *d, (const bf16*)x, (const bf16*)dy, dw, mps);
  else if (d->in_dtype == CRD_BF16 && d->out_dtype == CRD_F32)
    crd_launch(conv_wgrad_kernel<bf16, float>, dim3(grid), dim3(NT), 0, s, *d, (const bf16*)x, (const float*)dy, dw, mps);
  else
    return -2;
  CRD_LAUNCH_CHECK();
  return 0;
}
