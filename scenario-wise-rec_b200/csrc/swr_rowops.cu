// swr_rowops.cu -- the row-local glue of the stack, fused so that no intermediate
// [B, n_expert, H] tensor or per-domain mask ever reaches HBM:
//   pool   gate BatchNorm + softmax + gate-weighted expert pooling (mmoe.py:40-49,
//          ple.py:117-133 of the reference) and its backward
//   head   per-domain Linear(H,1) + sigmoid + domain mask-select (every model's tail)
//   bn     running-statistics update and d gamma / d beta of every BatchNorm of a pass
// One thread owns one (row, gate) / (row, expert) / (row, domain); column reductions are
// warp shuffles followed by one fp64 atomic per warp and column.
#include "swr_common.cuh"
#include "swr_launch.h"

namespace swr {

constexpr int kMaxGates = 16;
constexpr int kMaxExperts = 32;
constexpr int kMaxDomains = 16;
constexpr int kRowThreads = 128;

struct PoolParams {
  PoolGate gates[kMaxGates];
  ActDev experts[kMaxExperts];
  int n_gates, n_experts, B, H;
  float inv_count;
};

// coefficients of expert u, column h -> shared memory [u][h] (mu, s, b, r)
__device__ __forceinline__ void stage_expert_coefs(const PoolParams& p, float4* ec) {
  for (int i = threadIdx.x; i < p.n_experts * p.H; i += blockDim.x) {
    const int u = i / p.H, h = i - u * p.H;
    const ColCoef c = col_coef(p.experts[u].norm, h, p.inv_count);
    ec[i] = make_float4(c.mu, c.s, c.b, c.r);
  }
}

__global__ void __launch_bounds__(kRowThreads) pool_fwd_kernel(const __grid_constant__ PoolParams p) {
  extern __shared__ __align__(16) float4 ec[];       // [n_experts][H]
  __shared__ float4 gc[kMaxPoolExperts];
  const PoolGate& G = p.gates[blockIdx.y];
  stage_expert_coefs(p, ec);
  if (threadIdx.x < G.nE) {
    const ColCoef c = col_coef(G.gate.norm, threadIdx.x, p.inv_count);
    gc[threadIdx.x] = make_float4(c.mu, c.s, c.b, c.r);
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;

  float pr[kMaxPoolExperts];
  float mx = -INFINITY;
#pragma unroll
  for (int e = 0; e < kMaxPoolExperts; ++e) {
    if (e < G.nE) {
      pr[e] = fmaf(G.gate.raw[(int64_t)b * G.gate.ld + e] - gc[e].x, gc[e].y, gc[e].z);
      mx = fmaxf(mx, pr[e]);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int e = 0; e < kMaxPoolExperts; ++e)
    if (e < G.nE) { pr[e] = expf(pr[e] - mx); sum += pr[e]; }
  const float inv = 1.f / sum;
#pragma unroll
  for (int e = 0; e < kMaxPoolExperts; ++e)
    if (e < G.nE) { pr[e] *= inv; G.probs[(int64_t)b * G.nE + e] = pr[e]; }

  for (int h = 0; h < p.H; ++h) {
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < kMaxPoolExperts; ++e) {
      if (e < G.nE) {
        const int u = G.expert[e];
        const ActDev& X = p.experts[u];
        const float4 c = ec[u * p.H + h];
        acc = fmaf(pr[e], act_fwd(fmaf(X.raw[(int64_t)b * X.ld + h] - c.x, c.y, c.z), X.act), acc);
      }
    }
    G.pooled[(int64_t)b * G.ldp + h] = acc;
  }
}

// blockIdx.y < n_gates: gate phase; otherwise expert phase for expert blockIdx.y - n_gates
__global__ void __launch_bounds__(kRowThreads) pool_bwd_kernel(const __grid_constant__ PoolParams p) {
  extern __shared__ __align__(16) float4 ec[];
  __shared__ float4 gc[kMaxPoolExperts];
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = b < p.B;
  stage_expert_coefs(p, ec);

  if ((int)blockIdx.y < p.n_gates) {
    const PoolGate& G = p.gates[blockIdx.y];
    if (threadIdx.x < G.nE) {
      const ColCoef c = col_coef(G.gate.norm, threadIdx.x, p.inv_count);
      gc[threadIdx.x] = make_float4(c.mu, c.s, c.b, c.r);
    }
    __syncthreads();
    float dp[kMaxPoolExperts], pr[kMaxPoolExperts];
#pragma unroll
    for (int e = 0; e < kMaxPoolExperts; ++e) { dp[e] = 0.f; pr[e] = 0.f; }
    if (live) {
      for (int h = 0; h < p.H; ++h) {
        const float g = G.dpooled[(int64_t)b * G.ldp + h];
#pragma unroll
        for (int e = 0; e < kMaxPoolExperts; ++e) {
          if (e < G.nE) {
            const int u = G.expert[e];
            const ActDev& X = p.experts[u];
            const float4 c = ec[u * p.H + h];
            dp[e] = fmaf(g, act_fwd(fmaf(X.raw[(int64_t)b * X.ld + h] - c.x, c.y, c.z), X.act), dp[e]);
          }
        }
      }
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < kMaxPoolExperts; ++e)
        if (e < G.nE) { pr[e] = G.probs[(int64_t)b * G.nE + e]; dot = fmaf(pr[e], dp[e], dot); }
#pragma unroll
      for (int e = 0; e < kMaxPoolExperts; ++e)
        if (e < G.nE) dp[e] = pr[e] * (dp[e] - dot);      // d logits (softmax backward)
    }
#pragma unroll
    for (int e = 0; e < kMaxPoolExperts; ++e) {
      if (e < G.nE) {   // uniform across the warp
        double s1 = 0.0, s2 = 0.0;
        if (live) {
          const int64_t o = (int64_t)b * G.gate.ld + e;
          if (G.gate.dz) G.gate.dz[o] = dp[e];
          s1 = (double)dp[e];
          s2 = (double)dp[e] * (double)((G.gate.raw[o] - gc[e].x) * gc[e].w);
        }
        if (G.gate.dstats && G.gate.norm.mode != SWR_NORM_NONE) {
          s1 = warp_sum(s1); s2 = warp_sum(s2);
          if (lane == 0) { atomicAdd(G.gate.dstats + 2 * e, s1); atomicAdd(G.gate.dstats + 2 * e + 1, s2); }
        }
      }
    }
  } else {
    const int u = blockIdx.y - p.n_gates;
    const ActDev& X = p.experts[u];
    __syncthreads();
    for (int h = 0; h < p.H; ++h) {
      double s1 = 0.0, s2 = 0.0;
      if (live) {
        float dA = 0.f;
        for (int g = 0; g < p.n_gates; ++g) {
          const PoolGate& G = p.gates[g];
          for (int e = 0; e < G.nE; ++e)
            if (G.expert[e] == u) dA = fmaf(G.probs[(int64_t)b * G.nE + e], G.dpooled[(int64_t)b * G.ldp + h], dA);
        }
        const float4 c = ec[u * p.H + h];
        const int64_t o = (int64_t)b * X.ld + h;
        const float raw = X.raw[o];
        const float dz = dA * act_grad(fmaf(raw - c.x, c.y, c.z), X.act);
        if (X.dz) X.dz[o] = dz;
        s1 = (double)dz; s2 = (double)dz * (double)((raw - c.x) * c.w);
      }
      if (X.dstats && X.norm.mode != SWR_NORM_NONE) {
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) { atomicAdd(X.dstats + 2 * h, s1); atomicAdd(X.dstats + 2 * h + 1, s2); }
      }
    }
  }
}

static int fill_pool(const PoolLaunch& l, PoolParams& p) {
  if (l.n_gates <= 0 || l.n_gates > kMaxGates || l.n_experts <= 0 || l.n_experts > kMaxExperts) {
    set_error("pool: %d gates / %d experts unsupported (max %d / %d)", l.n_gates, l.n_experts, kMaxGates, kMaxExperts);
    return SWR_ERR_UNSUPPORTED;
  }
  if (l.H <= 0 || l.H > 1024) { set_error("pool: expert width %d unsupported", l.H); return SWR_ERR_UNSUPPORTED; }
  for (int g = 0; g < l.n_gates; ++g) {
    p.gates[g] = l.gates[g];
    if (l.gates[g].nE <= 0 || l.gates[g].nE > kMaxPoolExperts) { set_error("pool: gate over %d experts unsupported", l.gates[g].nE); return SWR_ERR_UNSUPPORTED; }
    for (int e = 0; e < l.gates[g].nE; ++e)
      if (l.gates[g].expert[e] < 0 || l.gates[g].expert[e] >= l.n_experts) { set_error("pool: bad expert index"); return SWR_ERR_INVALID; }
  }
  for (int u = 0; u < l.n_experts; ++u) p.experts[u] = l.experts[u];
  p.n_gates = l.n_gates; p.n_experts = l.n_experts; p.B = (int)l.B; p.H = l.H; p.inv_count = 1.0f / (float)l.B;
  return SWR_OK;
}

int launch_pool_fwd(const PoolLaunch& l, cudaStream_t st) {
  if (l.B <= 0) return SWR_OK;
  PoolParams p{};
  int rc = fill_pool(l, p);
  if (rc) return rc;
  dim3 grid(ceil_div(l.B, kRowThreads), l.n_gates);
  pool_fwd_kernel<<<grid, kRowThreads, sizeof(float4) * l.n_experts * l.H, st>>>(p);
  SWR_LAUNCH_OK("pool_fwd_kernel");
  return SWR_OK;
}

int launch_pool_bwd(const PoolLaunch& l, cudaStream_t st) {
  if (l.B <= 0) return SWR_OK;
  PoolParams p{};
  int rc = fill_pool(l, p);
  if (rc) return rc;
  dim3 grid(ceil_div(l.B, kRowThreads), l.n_gates + l.n_experts);
  pool_bwd_kernel<<<grid, kRowThreads, sizeof(float4) * l.n_experts * l.H, st>>>(p);
  SWR_LAUNCH_OK("pool_bwd_kernel");
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// head
// ---------------------------------------------------------------------------------------
struct HeadParams {
  HeadDomain dom[kMaxDomains];
  int n_domains, dom_dtype, sig_before_select, B;
  int hmax;                        // widest tower hidden layer
  const void* domain_id;
  float* out; const float* gout; const float* add; float* dadd;
  int ld_add;
  float inv_count;
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// Per-column coefficients of the lazily normalised tower activations are staged once per CTA: col_coef() is a
// dependent chain of an L2 load and fp64 rsqrt, far too slow to repeat per row and column.
__global__ void __launch_bounds__(kRowThreads) head_fwd_kernel(const __grid_constant__ HeadParams p) {
  extern __shared__ __align__(16) float4 hc[];   // [n_domains][hmax]: mu, s, b, r
  const int Hs = p.hmax;
  for (int i = threadIdx.x; i < p.n_domains * Hs; i += blockDim.x) {
    const int d = i / Hs, h = i - d * Hs;
    ColCoef c = {0.f, 1.f, 0.f, 1.f};
    if (h < p.dom[d].A.n) c = col_coef(p.dom[d].A.norm, h, p.inv_count);
    hc[i] = make_float4(c.mu, c.s, c.b, c.r);
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int64_t d = p.sig_before_select == SWR_HEAD_NO_SELECT ? 0 : load_index(p.domain_id, p.dom_dtype, b);
  float v = 0.f;
  const bool sel = d >= 0 && d < p.n_domains;
  if (sel) {
    const HeadDomain& D = p.dom[d];
    const int H = D.A.n;
    const float4* cd = hc + d * Hs;
    if (D.w) {
      v = ld_opt(D.bias, 0, 0.f);
      for (int h = 0; h < H; ++h) {
        const float4 c = cd[h];
        v = fmaf(act_fwd(fmaf(D.A.raw[(int64_t)b * D.A.ld + h] - c.x, c.y, c.z), D.A.act), __ldg(D.w + h), v);
      }
    } else {
      const float4 c = cd[0];
      v = act_fwd(fmaf(D.A.raw[(int64_t)b * D.A.ld] - c.x, c.y, c.z), D.A.act);
    }
  }
  float y;
  if (p.sig_before_select != SWR_HEAD_SIG_SELECT_ADD) y = sel ? sigmoidf_(v) : 0.f;
  else y = sigmoidf_(v + (p.add ? p.add[(int64_t)b * p.ld_add] : 0.f));
  p.out[b] = y;
}

// grid.y = domain.  Column sums leave the CTA as one atomic per column: warps combine in shared memory first.
__global__ void __launch_bounds__(kRowThreads) head_bwd_kernel(const __grid_constant__ HeadParams p) {
  extern __shared__ __align__(16) float4 hc[];   // [H] coefficients, then [H] dw (float), [H][2] stats (double)
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = b < p.B;
  const int d = blockIdx.y;
  const HeadDomain& D = p.dom[d];
  const int H = D.A.n;
  double* sst = reinterpret_cast<double*>(hc + H);
  float* sdw = reinterpret_cast<float*>(sst + 2 * H);
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const ColCoef c = col_coef(D.A.norm, h, p.inv_count);
    hc[h] = make_float4(c.mu, c.s, c.b, c.r);
    sst[2 * h] = 0.0; sst[2 * h + 1] = 0.0; sdw[h] = 0.f;
  }
  __syncthreads();
  float dv = 0.f, dsig = 0.f;
  if (live) {
    const int64_t di = p.sig_before_select == SWR_HEAD_NO_SELECT ? 0 : load_index(p.domain_id, p.dom_dtype, b);
    const float y = p.out[b];
    dsig = p.gout[b] * y * (1.f - y);
    if (di == d) dv = dsig;
    if (d == 0 && p.sig_before_select == SWR_HEAD_SIG_SELECT_ADD && p.dadd) p.dadd[(int64_t)b * p.ld_add] = dsig;
  }
  const bool stats = D.A.dstats && D.A.norm.mode != SWR_NORM_NONE;
  const bool wgrad = D.w && D.dw;
  for (int h = 0; h < H; ++h) {
    const float4 c = hc[h];
    float dw = 0.f; double s1 = 0.0, s2 = 0.0;
    if (live) {
      const int64_t o = (int64_t)b * D.A.ld + h;
      const float raw = D.A.raw[o];
      const float z = fmaf(raw - c.x, c.y, c.z);
      dw = dv * act_fwd(z, D.A.act);
      const float dA = D.w ? dv * __ldg(D.w + h) : dv;
      const float dz = dA * act_grad(z, D.A.act);
      if (D.A.dz) D.A.dz[o] = dz;
      s1 = (double)dz; s2 = (double)dz * (double)((raw - c.x) * c.w);
    }
    if (wgrad) { dw = warp_sum(dw); if (lane == 0 && dw != 0.f) atomicAdd(sdw + h, dw); }
    if (stats) {
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { atomicAdd(sst + 2 * h, s1); atomicAdd(sst + 2 * h + 1, s2); }
    }
  }
  if (D.w && D.dbias) { const float t = warp_sum(dv); if (lane == 0 && t != 0.f) atomicAdd(D.dbias, t); }
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    if (wgrad && sdw[h] != 0.f) atomicAdd(D.dw + h, sdw[h]);
    if (stats) { atomicAdd(D.A.dstats + 2 * h, sst[2 * h]); atomicAdd(D.A.dstats + 2 * h + 1, sst[2 * h + 1]); }
  }
}

static int fill_head(const HeadLaunch& l, HeadParams& p) {
  if (l.n_domains <= 0 || l.n_domains > kMaxDomains) { set_error("head: %d domains unsupported (max %d)", l.n_domains, kMaxDomains); return SWR_ERR_UNSUPPORTED; }
  for (int d = 0; d < l.n_domains; ++d) p.dom[d] = l.dom[d];
  p.n_domains = l.n_domains; p.dom_dtype = l.dom_dtype; p.sig_before_select = l.sig_before_select; p.B = (int)l.B;
  p.domain_id = l.domain_id; p.out = l.out; p.gout = l.gout; p.add = l.add; p.dadd = l.dadd; p.ld_add = l.ld_add;
  p.inv_count = 1.0f / (float)l.B;
  p.hmax = 1;
  for (int d = 0; d < l.n_domains; ++d) p.hmax = max(p.hmax, l.dom[d].A.n);
  return SWR_OK;
}

int launch_head_fwd(const HeadLaunch& l, cudaStream_t st) {
  if (l.B <= 0) return SWR_OK;
  HeadParams p{};
  int rc = fill_head(l, p);
  if (rc) return rc;
  head_fwd_kernel<<<ceil_div(l.B, kRowThreads), kRowThreads, sizeof(float4) * p.n_domains * p.hmax, st>>>(p);
  SWR_LAUNCH_OK("head_fwd_kernel");
  return SWR_OK;
}

int launch_head_bwd(const HeadLaunch& l, cudaStream_t st) {
  if (l.B <= 0) return SWR_OK;
  HeadParams p{};
  int rc = fill_head(l, p);
  if (rc) return rc;
  dim3 grid(ceil_div(l.B, kRowThreads), l.n_domains);
  head_bwd_kernel<<<grid, kRowThreads, (sizeof(float4) + 2 * sizeof(double) + sizeof(float)) * p.hmax, st>>>(p);
  SWR_LAUNCH_OK("head_bwd_kernel");
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// BatchNorm bookkeeping for every normalised activation of a pass (one launch)
// ---------------------------------------------------------------------------------------
constexpr int kBnPerLaunch = 48;
struct BnParams { BnLayer l[kBnPerLaunch]; int n; int B; float momentum; float inv_count; };

__global__ void __launch_bounds__(128) bn_update_kernel(const __grid_constant__ BnParams p) {
  const BnLayer& L = p.l[blockIdx.x];
  const double unbias = p.B > 1 ? (double)p.B / (double)(p.B - 1) : 1.0;
  for (int c = threadIdx.x; c < L.A.n; c += blockDim.x) {
    double mu, var;
    col_moments(L.A.norm, c, p.inv_count, mu, var);
    // torch: running = (1 - momentum) * running + momentum * batch_stat, unbiased variance
    const int rep = L.repeat > 0 ? L.repeat : 1;   // a module evaluated `rep` times on the same batch
    float rm = L.rmean ? L.rmean[c] : 0.f, rv = L.rvar ? L.rvar[c] : 0.f;
    for (int t = 0; t < rep; ++t) {
      rm = (1.f - p.momentum) * rm + p.momentum * (float)mu;
      rv = (1.f - p.momentum) * rv + p.momentum * (float)(var * unbias);
    }
    if (L.rmean) L.rmean[c] = rm;
    if (L.rvar) L.rvar[c] = rv;
  }
  if (threadIdx.x == 0 && L.nbt) *L.nbt += (L.repeat > 0 ? L.repeat : 1);
}

__global__ void __launch_bounds__(128) bn_pgrad_kernel(const __grid_constant__ BnParams p) {
  const BnLayer& L = p.l[blockIdx.x];
  for (int c = threadIdx.x; c < L.A.n; c += blockDim.x) {
    const float s1 = (float)L.A.dstats[2 * c], s2 = (float)L.A.dstats[2 * c + 1];
    const float g1 = ld_opt(L.A.norm.gamma, c, 1.f), g2 = ld_opt(L.A.norm.gamma2, c, 1.f);
    if (L.dgamma) atomicAdd(L.dgamma + c, s2 * g2);
    if (L.dgamma2) atomicAdd(L.dgamma2 + c, s2 * g1);
    if (L.dbeta) atomicAdd(L.dbeta + c, s1);
    if (L.dbeta2) atomicAdd(L.dbeta2 + c, s1);
  }
}

int launch_bn_update(const BnLayer* layers, int n, int64_t B, float momentum, cudaStream_t st) {
  for (int o = 0; o < n; o += kBnPerLaunch) {
    BnParams p{};
    p.n = min(kBnPerLaunch, n - o); p.B = (int)B; p.momentum = momentum; p.inv_count = 1.0f / (float)B;
    for (int i = 0; i < p.n; ++i) p.l[i] = layers[o + i];
    bn_update_kernel<<<p.n, 128, 0, st>>>(p);
    SWR_LAUNCH_OK("bn_update_kernel");
  }
  return SWR_OK;
}

int launch_bn_pgrad(const BnLayer* layers, int n, int64_t B, cudaStream_t st) {
  for (int o = 0; o < n; o += kBnPerLaunch) {
    BnParams p{};
    p.n = min(kBnPerLaunch, n - o); p.B = (int)B; p.momentum = 0.f; p.inv_count = 1.0f / (float)B;
    for (int i = 0; i < p.n; ++i) p.l[i] = layers[o + i];
    bn_pgrad_kernel<<<p.n, 128, 0, st>>>(p);
    SWR_LAUNCH_OK("bn_pgrad_kernel");
  }
  return SWR_OK;
}

}  // namespace swr
