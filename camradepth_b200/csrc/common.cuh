// Shared device helpers for the camradepth_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define CRD_F32 0
#define CRD_BF16 1

extern unsigned long long g_crd_launches;   // counted by every launcher (crd_launch_count)

#define CRD_LAUNCH_CHECK()                                   \
  do {                                                       \
    g_crd_launches++;                                        \
    cudaError_t e__ = cudaPeekAtLastError();                 \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

#define CRD_REQUIRE(cond) do { if (!(cond)) return -1000 - __LINE__; } while (0)

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// The training step is ~1800 mostly tiny, strictly ordered kernels.  A kernel launched through crd_launch carries the
// programmatic-stream-serialization attribute: its CTAs may be scheduled (and run their prologue: barrier init,
// tensor-map prefetch, TMEM allocation) while the previous kernel in the stream is still draining, and block in
// pdl_wait() until that kernel has completed and its memory is visible.  Every such kernel calls
// pdl_launch_dependents() first (lets ITS successor be scheduled early; always safe because the successor waits for
// full completion) and pdl_wait() before its first global-memory access.  Without the attribute both are no-ops.
// In a captured CUDA graph the attribute becomes a programmatic dependency edge.  Off unless CAMRADEPTH_PDL=1
// (measured neutral here, see lib.cu).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define CRD_PDL_ENTRY() do { pdl_launch_dependents(); pdl_wait(); } while (0)

bool crd_pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t crd_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = crd_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename T> struct TypeTag;
template <> struct TypeTag<float> { static constexpr int id = CRD_F32; };
template <> struct TypeTag<bf16> { static constexpr int id = CRD_BF16; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8-element vector load/store (channels are always padded to multiples of 8)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

// raw 8-element loads (kept packed so several can be in flight before any is unpacked)
struct f32x8 { float4 lo, hi; };
__device__ __forceinline__ uint4 ldg16(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ f32x8 ldg16(const float* p) {
  f32x8 r;
  r.lo = *reinterpret_cast<const float4*>(p);
  r.hi = *reinterpret_cast<const float4*>(p + 4);
  return r;
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void unpack8(const f32x8& r, float (&v)[8]) {
  v[0] = r.lo.x; v[1] = r.lo.y; v[2] = r.lo.z; v[3] = r.lo.w; v[4] = r.hi.x; v[5] = r.hi.y; v[6] = r.hi.z; v[7] = r.hi.w;
}
template <typename T> struct Raw8;
template <> struct Raw8<bf16> { typedef uint4 type; };
template <> struct Raw8<float> { typedef f32x8 type; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exact-erf GELU (nn.GELU default).  erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below fp32
// GroupNorm noise) so value and derivative share ONE exponential: exp(-x^2/2) is both the erf tail and the
// Gaussian pdf.  ~13 FP32 ops + 1 MUFU.EX2 + 1 MUFU.RCP instead of the ~40-op erff().
// MUFU.RCP without the IEEE-rounding fix-up (and its slow-path call) of __frcp_rn / operator/: 1 ulp
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// h = 0.5 * erfc(|x| / sqrt2) (the Gaussian tail beyond |x|) and e = exp(-x^2 / 2)
__device__ __forceinline__ void gelu_tail(float x, float& h, float& e) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_fast(fmaf(0.3275911f, z, 1.0f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));   // exp(-x^2/2)
  float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  poly = fmaf(t, poly, 0.5f * 1.421413741f);
  poly = fmaf(t, poly, 0.5f * -0.284496736f);
  poly = fmaf(t, poly, 0.5f * 0.254829592f);
  h = poly * t * e;
}
// gelu(x) = x * Phi(x) = relu(x) - |x| * h: no cancellation in either tail
__device__ __forceinline__ float gelu_f(float x) {
  float h, e;
  gelu_tail(x, h, e);
  return fmaxf(x, 0.f) - fabsf(x * h);
}
// gelu'(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_grad_f(float x) {
  float h, e;
  gelu_tail(x, h, e);
  const float cdf = x > 0.f ? 1.0f - h : h;
  return fmaf(x * e, 0.39894228040143267794f, cdf);
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_fast(1.0f + __expf(-x)); }

// Two elements per instruction (sm_100 FFMA2 / FMUL2 / FADD2): a 3-register FFMA issues every second cycle per
// scheduler, so the ~14 fma-pipe operations of GELU made the GELU streaming kernels fma-pipe bound (29 cycles per
// warp-element against a 200 us copy: 273 us measured).  The packed forms run the SAME operation sequence per lane
// (IEEE fma / mul / add), so results are bit-identical to the scalar functions above.
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }
__device__ __forceinline__ void gelu_tail2(float2 x, float2& h, float2& e) {
  const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), f2(0.70710678118654752440f));
  const float2 d = __ffma2_rn(f2(0.3275911f), z, f2(1.0f));
  const float2 t = make_float2(rcp_fast(d.x), rcp_fast(d.y));
  const float2 a = __fmul2_rn(__fmul2_rn(x, x), f2(-0.72134752044448170368f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
  float2 poly = __ffma2_rn(t, f2(0.5f * 1.061405429f), f2(0.5f * -1.453152027f));
  poly = __ffma2_rn(t, poly, f2(0.5f * 1.421413741f));
  poly = __ffma2_rn(t, poly, f2(0.5f * -0.284496736f));
  poly = __ffma2_rn(t, poly, f2(0.5f * 0.254829592f));
  h = __fmul2_rn(__fmul2_rn(poly, t), e);
}
__device__ __forceinline__ float2 gelu2(float2 x) {
  float2 h, e;
  gelu_tail2(x, h, e);
  const float2 xh = __fmul2_rn(x, h);
  return __fadd2_rn(make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)), make_float2(-fabsf(xh.x), -fabsf(xh.y)));
}
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
  float2 h, e;
  gelu_tail2(x, h, e);
  const float2 cdf = make_float2(x.x > 0.f ? 1.0f - h.x : h.x, x.y > 0.f ? 1.0f - h.y : h.y);
  return __ffma2_rn(__fmul2_rn(x, e), f2(0.39894228040143267794f), cdf);
}

static inline int crd_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// dtype dispatch helpers for launchers
#define CRD_DISPATCH_1(dt, T, ...)                          \
  if ((dt) == CRD_F32) { typedef float T; __VA_ARGS__; }    \
  else if ((dt) == CRD_BF16) { typedef bf16 T; __VA_ARGS__; } \
  else return -2;
