// swr_train.cu -- the trainer-side ops that let one CUDA graph hold a whole training step
// (reference loop: trainers/ctr_trainer.py:69-73): BCELoss forward + gradient, and torch.optim.Adam
// (dense, L2 weight decay) over the flat parameter / gradient / moment arenas.  Both are HBM-bound
// streaming kernels; Adam moves 28 B per element (read p, g, m, v; write p, m, v) with 16-byte
// vector accesses over a grid sized to the SM count.
#include "swr_common.cuh"
#include "swr_launch.h"

namespace swr {

// ---------------------------------------------------------------------------------------
// BCELoss (torch.nn.BCELoss, mean reduction): loss = -mean(y log p + (1-y) log(1-p)), logs clamped at -100;
// gout = (p - y) / max(p (1 - p), 1e-12) / B        (ATen binary_cross_entropy_backward)
// One CTA: B is a few thousand rows and the sum must be deterministic.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bce_kernel(const float* __restrict__ pred, const void* __restrict__ label, int label_dtype,
                                                   float* __restrict__ gout, float* __restrict__ loss_ring,
                                                   const int32_t* __restrict__ ctrl, int ring, int B, float gscale) {
  __shared__ double red[32];
  double acc = 0.0;
  const float inv = gscale / (float)B;   // gscale = 1 / world when data-parallel gradients are summed instead of averaged
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float p = pred[b], y = load_scalar(label, label_dtype, b);
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
    acc -= (double)(y * lp + (1.f - y) * l1p);
    if (gout) gout[b] = (p - y) / fmaxf((1.f - p) * p, 1e-12f) * inv;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    acc = warp_sum(acc);
    if (threadIdx.x == 0 && loss_ring) {
      const int slot = ctrl ? (ctrl[0] % ring + ring) % ring : 0;
      loss_ring[slot] = (float)(acc / (double)B);
    }
  }
}

int launch_bce(const float* pred, const void* label, int label_dtype, float* gout, float* loss_ring, const int32_t* ctrl,
               int ring, int64_t B, float gscale, cudaStream_t st) {
  if (B <= 0) return SWR_OK;
  if (!pred || !label) { set_error("bce: null operand"); return SWR_ERR_INVALID; }
  bce_kernel<<<1, 1024, 0, st>>>(pred, label, label_dtype, gout, loss_ring, ctrl, ring > 0 ? ring : 1, (int)B, gscale);
  SWR_LAUNCH_OK("bce_kernel");
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, amsgrad=False, maximize=False):
//   g += wd * p;  m = lerp(m, g, 1-b1);  v = b2 v + (1-b2) g^2;  p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps)
// hyper (device floats, rewritten by the host every step): lr/bc1, b1, b2, eps, wd, 1/sqrt(bc2)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float step_size, float b1, float b2,
                                          float eps, float wd, float inv_bc2_sqrt) {
  g = fmaf(wd, p, g);
  m = fmaf(1.f - b1, g - m, m);
  v = fmaf(1.f - b2, g * g, b2 * v);
  const float denom = fmaf(sqrtf(v), inv_bc2_sqrt, eps);
  p = fmaf(-step_size, m / denom, p);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const float* __restrict__ hyper, int64_t n, int zero_grad) {
  const float step_size = __ldg(hyper + 0), b1 = __ldg(hyper + 1), b2 = __ldg(hyper + 2), eps = __ldg(hyper + 3),
              wd = __ldg(hyper + 4), ibc2 = __ldg(hyper + 5);
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(p); float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = p4[i], G = g4[i], M = m4[i], V = v4[i];
    adam_elem(P.x, G.x, M.x, V.x, step_size, b1, b2, eps, wd, ibc2);
    adam_elem(P.y, G.y, M.y, V.y, step_size, b1, b2, eps, wd, ibc2);
    adam_elem(P.z, G.z, M.z, V.z, step_size, b1, b2, eps, wd, ibc2);
    adam_elem(P.w, G.w, M.w, V.w, step_size, b1, b2, eps, wd, ibc2);
    p4[i] = P; m4[i] = M; v4[i] = V;
    if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float P = p[i], M = m[i], V = v[i];
    adam_elem(P, g[i], M, V, step_size, b1, b2, eps, wd, ibc2);
    p[i] = P; m[i] = M; v[i] = V;
    if (zero_grad) g[i] = 0.f;
  }
}

int launch_adam(float* p, float* g, float* m, float* v, const float* hyper, int64_t n, int zero_grad, cudaStream_t st) {
  if (n <= 0) return SWR_OK;
  if (!p || !g || !m || !v || !hyper) { set_error("adam: null operand"); return SWR_ERR_INVALID; }
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) {
    set_error("adam: arenas must be 16-byte aligned"); return SWR_ERR_INVALID;
  }
  const int64_t want = (n / 4 + 255) / 256;
  const int grid = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  adam_kernel<<<grid, 256, 0, st>>>(p, g, m, v, hyper, n, zero_grad);
  SWR_LAUNCH_OK("adam_kernel");
  return SWR_OK;
}

}  // namespace swr
