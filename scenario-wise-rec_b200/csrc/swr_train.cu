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

// ---------------------------------------------------------------------------------------
// Row-lazy Adam for embedding tables: the same trajectory as the dense sweep, without touching every row every step.
//
// torch.optim.Adam on a dense table gradient (ctr_trainer.py:50-52,73) updates EVERY row at every step: a row no sample
// looked up has g = 0 and still moves (weight decay pulls it, its moments decay).  Those updates depend on nothing but the
// row's own (p, m, v) and the step's scalars, so they can be postponed and replayed later with bit-identical results:
//   * last[r]   the step up to which row r is current,
//   * hist[s]   the scalars of step s (lr / bias-correction-1, 1 / sqrt(bias-correction-2), weight decay), appended by
//               the update kernel of step s,
//   * catch-up  before the forward gather, every row of the batch is replayed to step t - 1 (g = 0 steps),
//   * update    after the backward scatter, every row of the batch gets step t with its gradient, and its gradient row
//               is zeroed again (the dense gradient arena stays all-zero between steps: no O(vocab) memset),
//   * flush     every row is replayed to the current step: bounds the replay length, and makes the parameters readable
//               by anything outside the fused step (state_dict, evaluation, checkpoints).
// A row is claimed for a phase with atomicMax on claim[r] (2 t for the catch-up, 2 t + 1 for the update), so duplicate
// lookups of one row inside a batch do the work once.  adam_elem is the dense kernel's own update: same instructions,
// same rounding.
// ---------------------------------------------------------------------------------------
constexpr int kLazyMax = 48;
struct LazyParams {
  LazyField f[kLazyMax];
  int n_fields; int B;
  const float* hyper;       // lr/bc1, b1, b2, eps, wd, 1/sqrt(bc2) of the current step
  const int32_t* ctrl;      // ctrl[1] = current step t (1-based), ctrl[2] = first step of the history table
  float4* hist;             // [capacity] scalars per step, index s - ctrl[2]
};

// hist_s: the last kLazyWin steps' scalars staged in shared memory (hist_s[i] = step win0 + i); older steps come from
// global memory (one dependent L2 access per replayed step, which is what made the first version of these kernels slow)
constexpr int kLazyWin = 64;
__device__ __forceinline__ void lazy_replay(float& P, float& M, float& V, int from, int to, const float4* hist, int base,
                                            const float4* hist_s, int win0, float b1, float b2, float eps) {
  for (int s = from; s <= to; ++s) {          // steps the row sat out: zero gradient
    const float4 h = s >= win0 ? hist_s[s - win0] : __ldg(hist + (s - base));
    adam_elem(P, 0.f, M, V, h.x, b1, b2, eps, h.z, h.y);
  }
}
// stage hist[max(base, t - kLazyWin + 1) .. t] (entry t only if `with_t`) ; returns win0
__device__ __forceinline__ int lazy_stage_hist(float4* hist_s, const float4* hist, int base, int t, bool with_t) {
  const int win0 = max(base, t - kLazyWin + 1);
  const int last = with_t ? t : t - 1;
  for (int s = win0 + (int)threadIdx.x; s <= last; s += blockDim.x) hist_s[s - win0] = __ldg(hist + (s - base));
  __syncthreads();
  return win0;
}

// phase 0: catch-up to t - 1;  phase 1: (catch-up +) step t with the gradient row, gradient row zeroed
template <int PHASE>
__global__ void __launch_bounds__(256) adam_rows_kernel(const __grid_constant__ LazyParams p) {
  const int t = p.ctrl[1], base = p.ctrl[2];
  const float b1 = __ldg(p.hyper + 1), b2 = __ldg(p.hyper + 2), eps = __ldg(p.hyper + 3);
  if (PHASE == 1 && blockIdx.x == 0 && threadIdx.x == 0)
    p.hist[t - base] = make_float4(__ldg(p.hyper + 0), __ldg(p.hyper + 5), __ldg(p.hyper + 4), 0.f);
  __shared__ float4 hist_s[kLazyWin];
  const int win0 = lazy_stage_hist(hist_s, p.hist, base, t, false);
  // One grid-stride loop over every (field, batch row) lookup of the launch (a single wave of CTAs).  E <= 32: one row
  // per `E` lanes (E = 16: two lookups per warp pass); wider rows: a warp strides over the row.
  const int E = p.f[0].E, lanes = E < 32 ? E : 32, per_warp = 32 / lanes;      // every field of a launch shares E (host checks)
  const int lane = threadIdx.x & 31, sub = lane / lanes, l = lane - sub * lanes;
  const int64_t total = (int64_t)p.n_fields * p.B;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * per_warp; i0 < total; i0 += n_warps * per_warp) {
    const int64_t i = i0 + sub;
    const bool in = i < total && sub < per_warp;
    const int f = in ? (int)(i / p.B) : 0;
    const LazyField& F = p.f[f];
    int64_t r = -1;
    if (in) {
      r = load_index(F.idx, F.idx_dtype, i - (int64_t)f * p.B);
      if (r < 0 || r >= F.vocab) r = -1;
      else if (F.world > 1) r = (r % F.world == F.rank) ? r / F.world : -1;      // only the owner updates a row
    }
    // everything the owner of the row needs is requested at once; a lookup that turns out not to own the row (a
    // duplicate inside the batch) has read one row in vain
    int old = 0x7fffffff, last = 0;
    float P[2] = {0.f, 0.f}, M[2] = {0.f, 0.f}, V[2] = {0.f, 0.f}, G[2] = {0.f, 0.f};
    if (r >= 0) {
      if (l == 0) old = __ldcg(F.claim + r);
      last = __ldcg(F.last + r);
      if (E <= 64) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int e = l + k * lanes;
          if (e < E) {
            const int64_t o = r * E + e;
            P[k] = F.p[o]; M[k] = F.m[o]; V[k] = F.v[o];
            if (PHASE == 1) G[k] = F.g[o];
          }
        }
      }
      // hot rows (a 3-row table is looked up by the whole batch): only lookups that still see the row unclaimed pay for
      // an atomic on its claim word
      if (l == 0 && old < 2 * t + PHASE) old = atomicMax(F.claim + r, 2 * t + PHASE);
    }
    old = __shfl_sync(0xffffffffu, old, (sub * lanes) & 31);
    const bool own = r >= 0 && old < 2 * t + PHASE;       // in range, and no other lookup of the batch owns the row
    if (own) {
      if (E <= 64) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int e = l + k * lanes;
          if (e < E) {
            const int64_t o = r * E + e;
            lazy_replay(P[k], M[k], V[k], last + 1, t - 1, p.hist, base, hist_s, win0, b1, b2, eps);
            if (PHASE == 1) {
              adam_elem(P[k], G[k], M[k], V[k], __ldg(p.hyper + 0), b1, b2, eps, __ldg(p.hyper + 4), __ldg(p.hyper + 5));
              F.g[o] = 0.f;
            }
            F.p[o] = P[k]; F.m[o] = M[k]; F.v[o] = V[k];
          }
        }
      } else {
        for (int e = l; e < E; e += lanes) {
          const int64_t o = r * E + e;
          float Pe = F.p[o], Me = F.m[o], Ve = F.v[o];
          lazy_replay(Pe, Me, Ve, last + 1, t - 1, p.hist, base, hist_s, win0, b1, b2, eps);
          if (PHASE == 1) {
            adam_elem(Pe, F.g[o], Me, Ve, __ldg(p.hyper + 0), b1, b2, eps, __ldg(p.hyper + 4), __ldg(p.hyper + 5));
            F.g[o] = 0.f;
          }
          F.p[o] = Pe; F.m[o] = Me; F.v[o] = Ve;
        }
      }
    }
    __syncwarp();
    if (own && l == 0) F.last[r] = (PHASE == 1) ? t : t - 1;
  }
}

// every row of one table to step t (ctrl[1]); one thread per element
__global__ void __launch_bounds__(256) adam_flush_kernel(float* __restrict__ P, float* __restrict__ M, float* __restrict__ V,
                                                         int* __restrict__ last, int64_t vocab, int E, const float* __restrict__ hyper,
                                                         const int32_t* __restrict__ ctrl, const float4* __restrict__ hist) {
  const int t = ctrl[1], base = ctrl[2];
  const float b1 = __ldg(hyper + 1), b2 = __ldg(hyper + 2), eps = __ldg(hyper + 3);
  __shared__ float4 hist_s[kLazyWin];
  const int win0 = lazy_stage_hist(hist_s, hist, base, t, true);
  const int64_t n = vocab * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E;
    const int from = last[r] + 1;
    if (from > t) continue;
    float p = P[i], m = M[i], v = V[i];
    lazy_replay(p, m, v, from, t, hist, base, hist_s, win0, b1, b2, eps);
    P[i] = p; M[i] = m; V[i] = v;
  }
}
__global__ void __launch_bounds__(256) adam_flush_mark_kernel(int* __restrict__ last, int64_t vocab, const int32_t* __restrict__ ctrl) {
  const int t = ctrl[1];
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < vocab; r += (int64_t)gridDim.x * blockDim.x)
    if (last[r] < t) last[r] = t;
}

int launch_adam_rows(const LazyField* fields, int n_fields, int64_t B, const float* hyper, const int32_t* ctrl, float4* hist, int phase,
                     cudaStream_t st) {
  if (n_fields <= 0 || B <= 0) return SWR_OK;
  if (!hyper || !ctrl || !hist) { set_error("adam_rows: null operand"); return SWR_ERR_INVALID; }
  int o = 0;
  while (o < n_fields) {       // one launch per run of fields with the same embedding width (normally: one launch)
    LazyParams p{};
    const int E0 = fields[o].E;
    while (o < n_fields && p.n_fields < kLazyMax && fields[o].E == E0) {
      p.f[p.n_fields] = fields[o++];
      const LazyField& F = p.f[p.n_fields++];
      if (!F.p || !F.g || !F.m || !F.v || !F.last || !F.claim || !F.idx || F.E <= 0) { set_error("adam_rows: bad field operand"); return SWR_ERR_INVALID; }
    }
    p.B = (int)B; p.hyper = hyper; p.ctrl = ctrl; p.hist = hist;
    const int64_t per_block = 8 * (E0 < 32 ? 32 / E0 : 1);
    int64_t gx = ((int64_t)p.n_fields * B + per_block - 1) / per_block;
    if (gx > 148 * 8) gx = 148 * 8;      // one wave (256 threads, 8 CTAs per SM)
    if (phase == 0) adam_rows_kernel<0><<<(int)gx, 256, 0, st>>>(p);
    else adam_rows_kernel<1><<<(int)gx, 256, 0, st>>>(p);
    SWR_LAUNCH_OK("adam_rows_kernel");
  }
  return SWR_OK;
}

int launch_adam_flush(const LazyField* fields, int n_fields, const float* hyper, const int32_t* ctrl, const float4* hist, cudaStream_t st) {
  for (int i = 0; i < n_fields; ++i) {
    const LazyField& F = fields[i];
    const int64_t n = F.vocab * F.E;
    if (n <= 0) continue;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    adam_flush_kernel<<<grid, 256, 0, st>>>(F.p, F.m, F.v, F.last, F.vocab, F.E, hyper, ctrl, hist);
    SWR_LAUNCH_OK("adam_flush_kernel");
    int g2 = (int)((F.vocab + 255) / 256);
    if (g2 > 148 * 8) g2 = 148 * 8;
    adam_flush_mark_kernel<<<g2, 256, 0, st>>>(F.last, F.vocab, ctrl);
    SWR_LAUNCH_OK("adam_flush_mark_kernel");
  }
  return SWR_OK;
}

}  // namespace swr
