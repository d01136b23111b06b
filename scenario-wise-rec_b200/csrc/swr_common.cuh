// swr_common.cuh -- device-side descriptors and helpers shared by every kernel.
//
// Data model (DESIGN.md "Lazy activations"): an activation is stored RAW (the Linear
// output before its BatchNorm) together with the fp64 column moments the producing
// kernel accumulated.  Consumers apply normalise+activate while staging the operand
// (a = act(gamma * (raw - mu) * rstd + beta)), so every layer is exactly one pass over
// its activations.  The backward keeps the same shape: dz = dA * act'(z) is written by
// the producer of dA, the reduced statistics (sum dz, sum dz*xhat) are fp64 atomics, and
// dY = c0*dz + c1*raw + c2 is formed on load by the layer's own dgrad/wgrad kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/swr_b200.h"

namespace swr {

constexpr int kMaxGroups = 32;   // groups per grouped launch (kernel-parameter budget)

struct NormRef {
  const double* stats;   // [n][2]  (sum, sum of squares) interleaved     (SWR_NORM_BATCH)
  const float* rmean;    // [n]                                            (SWR_NORM_RUNNING)
  const float* rvar;
  const float* gamma;    // scale = gamma * gamma2 (gamma2 optional: STAR partitioned norm)
  const float* gamma2;
  const float* beta;     // shift = beta + beta2
  const float* beta2;
  int mode;
  float eps;
  float var_scale;       // variance used = var_scale * biased batch variance: 1 for BatchNorm / STAR,
                         // B/(B-1) for HAMUR's domain norm (tmp_out.var(dim=0), hamur.py:192-195)
};

struct ActDev {
  const float* raw;      // [B, ld]
  float* dz;             // [B, ld]  gradient wrt the pre-activation z (wrt the value if plain)
  double* dstats;        // [n][2]   (sum dz, sum dz * xhat) interleaved
  NormRef norm;
  int ld;
  int n;
  int act;
};

// Per-column coefficients of a lazily normalised activation.
//   forward :  z = (raw - mu) * s + b ,  a = act(z)      with s = gamma * rstd
//   xhat    :  (raw - mu) * r
struct ColCoef { float mu, s, b, r; };
// dY = c0 * dz + c1 * raw + c2  (BatchNorm backward folded into an affine map per column)
struct DyCoef { float c0, c1, c2; };

__device__ __forceinline__ float ld_opt(const float* p, int i, float dflt) { return p ? __ldg(p + i) : dflt; }

__device__ __forceinline__ void col_moments(const NormRef& nr, int c, float inv_count, double& mu, double& var) {
  if (nr.mode == SWR_NORM_BATCH) {
    // moments were accumulated in fp64 by the producer, so E[x^2]-E[x]^2 is safe.
    // __ldcg: written by another kernel's atomics, read through L2.
    const double s1 = __ldcg(nr.stats + 2 * c), s2 = __ldcg(nr.stats + 2 * c + 1);
    mu = s1 * (double)inv_count;
    var = s2 * (double)inv_count - mu * mu;
    if (var < 0.0) var = 0.0;
    var *= (double)nr.var_scale;
  } else {
    mu = (double)__ldg(nr.rmean + c);
    var = (double)__ldg(nr.rvar + c);
  }
}

__device__ __forceinline__ ColCoef col_coef(const NormRef& nr, int c, float inv_count) {
  ColCoef k;
  if (nr.mode == SWR_NORM_NONE) { k.mu = 0.f; k.s = 1.f; k.b = 0.f; k.r = 1.f; return k; }
  double mu, var;
  col_moments(nr, c, inv_count, mu, var);
  // the moments need fp64 (E[x^2] - E[x]^2); the reciprocal standard deviation does not: correctly rounded fp32 sqrt and
  // divide (what the fp32 reference computes), a handful of instructions instead of a double-precision sqrt + divide
  const float r = 1.0f / sqrtf((float)var + nr.eps);
  const float g = ld_opt(nr.gamma, c, 1.f) * ld_opt(nr.gamma2, c, 1.f);
  k.mu = (float)mu; k.r = r; k.s = g * r;
  k.b = ld_opt(nr.beta, c, 0.f) + ld_opt(nr.beta2, c, 0.f);
  return k;
}

// Stage 2 of the BatchNorm backward for the column c of activation `a`:
//   batch  : dY = s * (dz - S1/B - v * xhat * S2/B)  S1 = sum dz, S2 = sum dz*xhat, v = var_scale
//   running: dY = s * dz
//   none   : dY = dz
__device__ __forceinline__ DyCoef dy_coef(const ActDev& a, int c, float inv_count) {
  DyCoef d;
  if (a.norm.mode == SWR_NORM_NONE) { d.c0 = 1.f; d.c1 = 0.f; d.c2 = 0.f; return d; }
  const ColCoef k = col_coef(a.norm, c, inv_count);
  d.c0 = k.s;
  if (a.norm.mode == SWR_NORM_RUNNING) { d.c1 = 0.f; d.c2 = 0.f; return d; }
  // the sums were reduced in fp64; the per-column coefficients are formed in fp32
  const float S1 = (float)__ldcg(a.dstats + 2 * c), S2 = (float)__ldcg(a.dstats + 2 * c + 1);
  const float ib = inv_count * a.norm.var_scale;
  const float t = k.s * k.r * S2 * ib;
  d.c1 = -t;
  d.c2 = fmaf(t, k.mu, -k.s * S1 * inv_count);
  return d;
}

__device__ __forceinline__ float act_fwd(float z, int act) {
  switch (act) {
    case SWR_ACT_RELU: return fmaxf(z, 0.f);
    case SWR_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case SWR_ACT_LEAKY: return z > 0.f ? z : 0.1f * z;
    default: return z;
  }
}
// d act / d z given z
__device__ __forceinline__ float act_grad(float z, int act) {
  switch (act) {
    case SWR_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case SWR_ACT_SIGMOID: { const float a = 1.f / (1.f + expf(-z)); return a * (1.f - a); }
    case SWR_ACT_LEAKY: return z > 0.f ? 1.f : 0.1f;
    default: return 1.f;
  }
}
// forward value of a lazily normalised element
__device__ __forceinline__ float act_value(float raw, const ColCoef& k, int act) {
  return act_fwd(fmaf(raw - k.mu, k.s, k.b), act);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int64_t load_index(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case SWR_I8: return (int64_t) reinterpret_cast<const int8_t*>(p)[i];
    case SWR_U8: return (int64_t) reinterpret_cast<const uint8_t*>(p)[i];
    case SWR_I16: return (int64_t) reinterpret_cast<const int16_t*>(p)[i];
    case SWR_I32: return (int64_t) reinterpret_cast<const int32_t*>(p)[i];
    default: return reinterpret_cast<const int64_t*>(p)[i];
  }
}
__device__ __forceinline__ float load_scalar(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case SWR_F32: return reinterpret_cast<const float*>(p)[i];
    case SWR_F16: return __half2float(reinterpret_cast<const __half*>(p)[i]);
    case SWR_BF16: return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    case SWR_F64: return (float) reinterpret_cast<const double*>(p)[i];
    default: return (float) load_index(p, dtype, i);
  }
}

__device__ __forceinline__ bool is_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// rows [.., R) x cols [c, c+4) of a row-major matrix, zero outside
__device__ __forceinline__ float4 load4_guard(const float* __restrict__ base, int64_t ld, int r, int c, int R, int C, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < R && c < C) {
    const float* p = base + (int64_t)r * ld + c;
    if (vec && c + 3 < C) {
      v = *reinterpret_cast<const float4*>(p);
    } else {
      v.x = p[0];
      if (c + 1 < C) v.y = p[1];
      if (c + 2 < C) v.z = p[2];
      if (c + 3 < C) v.w = p[3];
    }
  }
  return v;
}

// ---- host side -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
#define SWR_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      swr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SWR_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)
#define SWR_LAUNCH_OK(name)                                                            \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      swr::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));         \
      return SWR_ERR_CUDA;                                                             \
    }                                                                                  \
    swr::count_launch();                                                               \
  } while (0)

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace swr
