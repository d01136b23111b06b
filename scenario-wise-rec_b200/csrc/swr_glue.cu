// swr_glue.cu -- element-wise and row-local ops between the grouped FC launches of the
// STAR / PPNet / EPNet / M3oE / HAMUR stacks.  Every op is HBM-bound: one coalesced pass over
// its [B, n] operands, normalise/activate applied on load (lazy activations), the column sums
// a BatchNorm backward needs reduced in shared memory and added with one fp64 atomic per column
// and CTA.
//
// "Column-tile" kernels: 256 threads = 32 columns x 8 row lanes; grid.x = column tiles,
// grid.y = row chunks, grid.z = group.  A warp reads 128 contiguous bytes of a row.
#include "swr_common.cuh"
#include "swr_launch.h"

namespace swr {

constexpr int kEwGroups = 24;
constexpr int kRowsPerCta = 128;

__host__ __device__ __forceinline__ bool is_lazy(const ActDev& a) { return a.norm.mode != SWR_NORM_NONE || a.act != SWR_ACT_NONE; }

// stage 1 of the backward of a lazy activation for one element: dz = dA * act'(z), column sums
__device__ __forceinline__ void grad_store(const ActDev& A, const ColCoef& k, bool lazy, int64_t o, float dA, bool accumulate,
                                           double& s1, double& s2) {
  float dz = dA;
  if (lazy) {
    const float raw = A.raw[o];
    dz *= act_grad(fmaf(raw - k.mu, k.s, k.b), A.act);
    s1 += (double)dz; s2 += (double)dz * (double)((raw - k.mu) * k.r);
  }
  if (A.dz) A.dz[o] = accumulate ? A.dz[o] + dz : dz;
}

// reduce (s1, s2) over the 8 row lanes of the CTA and add them to dstats[c]
__device__ __forceinline__ void block_col_atomic(double s1, double s2, double* dstats, int c, bool valid) {
  __shared__ double red[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  __syncthreads();
  red[0][ry][cx] = s1; red[1][ry][cx] = s2;
  __syncthreads();
  if (ry == 0 && valid && dstats) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { s1 += red[0][k][cx]; s2 += red[1][k][cx]; }
    atomicAdd(dstats + 2 * c, s1);
    atomicAdd(dstats + 2 * c + 1, s2);
  }
}

// ---------------------------------------------------------------------------------------
// EW: out = value(A) * value(C) * scale | value(A) + value(C) | value(A)
// ---------------------------------------------------------------------------------------
struct EwParams { EwGroup g[kEwGroups]; int n_groups; int B; float inv_count; };

__global__ void __launch_bounds__(256) ew_fwd_kernel(const __grid_constant__ EwParams p) {
  const EwGroup& G = p.g[blockIdx.z];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  if (c >= G.A.n) return;
  const int r0 = blockIdx.y * kRowsPerCta, r1 = min(p.B, r0 + kRowsPerCta);
  const ColCoef ka = col_coef(G.A.norm, c, p.inv_count);
  ColCoef kc = {0.f, 1.f, 0.f, 1.f};
  if (G.mode != SWR_EW_COPY) kc = col_coef(G.C.norm, c, p.inv_count);
  for (int r = r0 + ry; r < r1; r += 8) {
    const float a = act_value(G.A.raw[(int64_t)r * G.A.ld + c], ka, G.A.act);
    float v = a;
    if (G.mode != SWR_EW_COPY) {
      const float b = act_value(G.C.raw[(int64_t)r * G.C.ld + c], kc, G.C.act);
      v = (G.mode == SWR_EW_MUL) ? a * b * G.scale : a + b;
    }
    G.out[(int64_t)r * G.ld_out + c] = v;
  }
}

__global__ void __launch_bounds__(256) ew_bwd_kernel(const __grid_constant__ EwParams p) {
  const EwGroup& G = p.g[blockIdx.z];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool valid = c < G.A.n;
  const int r0 = blockIdx.y * kRowsPerCta, r1 = min(p.B, r0 + kRowsPerCta);
  const bool gradA = (G.flags & EW_A_GRAD) != 0, gradC = (G.flags & EW_C_GRAD) != 0 && G.mode != SWR_EW_COPY;
  const bool lazyA = is_lazy(G.A), lazyC = G.mode != SWR_EW_COPY && is_lazy(G.C);
  double a1 = 0.0, a2 = 0.0, c1 = 0.0, c2 = 0.0;
  if (valid) {
    const ColCoef ka = col_coef(G.A.norm, c, p.inv_count);
    ColCoef kc = {0.f, 1.f, 0.f, 1.f};
    if (G.mode != SWR_EW_COPY) kc = col_coef(G.C.norm, c, p.inv_count);
    for (int r = r0 + ry; r < r1; r += 8) {
      const float g = G.dout[(int64_t)r * G.ld_out + c];
      const int64_t oa = (int64_t)r * G.A.ld + c, oc = (int64_t)r * G.C.ld + c;
      float dA = g, dC = g;
      if (G.mode == SWR_EW_MUL) {
        const float a = act_value(G.A.raw[oa], ka, G.A.act), b = act_value(G.C.raw[oc], kc, G.C.act);
        dA = g * b * G.scale; dC = g * a * G.scale;
      }
      if (gradA) grad_store(G.A, ka, lazyA, oa, dA, (G.flags & EW_A_ACC) != 0, a1, a2);
      if (gradC) grad_store(G.C, kc, lazyC, oc, dC, (G.flags & EW_C_ACC) != 0, c1, c2);
    }
  }
  if (gradA && G.A.norm.mode != SWR_NORM_NONE) block_col_atomic(a1, a2, G.A.dstats, c, valid);
  if (gradC && G.C.norm.mode != SWR_NORM_NONE) block_col_atomic(c1, c2, G.C.dstats, c, valid);
}

static int fill_ew(const EwGroup* g, int n, int64_t B, EwParams& p, int& nmax) {
  if (n <= 0 || n > kEwGroups) { set_error("ew: %d groups (max %d per launch)", n, kEwGroups); return SWR_ERR_INVALID; }
  nmax = 0;
  for (int i = 0; i < n; ++i) {
    p.g[i] = g[i];
    if (!g[i].A.raw || !g[i].out || (g[i].mode != SWR_EW_COPY && !g[i].C.raw)) { set_error("ew: null operand in group %d", i); return SWR_ERR_INVALID; }
    if (g[i].mode != SWR_EW_COPY && g[i].A.n != g[i].C.n) { set_error("ew: operand widths differ in group %d", i); return SWR_ERR_INVALID; }
    nmax = max(nmax, g[i].A.n);
  }
  p.n_groups = n; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  return SWR_OK;
}

int launch_ew_fwd(const EwGroup* g, int n, int64_t B, cudaStream_t st) {
  if (B <= 0) return SWR_OK;
  for (int o = 0; o < n; o += kEwGroups) {
    EwParams p{}; int nmax;
    const int m = min(kEwGroups, n - o);
    int rc = fill_ew(g + o, m, B, p, nmax);
    if (rc) return rc;
    dim3 grid(ceil_div(nmax, 32), ceil_div(B, kRowsPerCta), m);
    ew_fwd_kernel<<<grid, 256, 0, st>>>(p);
    SWR_LAUNCH_OK("ew_fwd_kernel");
  }
  return SWR_OK;
}

int launch_ew_bwd(const EwGroup* g, int n, int64_t B, cudaStream_t st) {
  if (B <= 0) return SWR_OK;
  for (int o = 0; o < n; o += kEwGroups) {
    EwParams p{}; int nmax;
    const int m = min(kEwGroups, n - o);
    int rc = fill_ew(g + o, m, B, p, nmax);
    if (rc) return rc;
    for (int i = 0; i < m; ++i) if (!p.g[i].dout) { set_error("ew_bwd: group %d has no output gradient", i); return SWR_ERR_INVALID; }
    dim3 grid(ceil_div(nmax, 32), ceil_div(B, kRowsPerCta), m);
    ew_bwd_kernel<<<grid, 256, 0, st>>>(p);
    SWR_LAUNCH_OK("ew_bwd_kernel");
  }
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// SUMGRAD: dst.dz (+)= sum_v (c0_v * dz_v + c1_v * raw + c2_v)     (views share dst's raw tensor)
// ---------------------------------------------------------------------------------------
struct SumGradParams { ActDev dst; ActDev v[kMaxViews]; int n_views; int accumulate; int B; float inv_count; };

__global__ void __launch_bounds__(256) sumgrad_kernel(const __grid_constant__ SumGradParams p) {
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  if (c >= p.dst.n) return;
  const int r0 = blockIdx.y * kRowsPerCta, r1 = min(p.B, r0 + kRowsPerCta);
  float c0[kMaxViews], c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int v = 0; v < kMaxViews; ++v) {
    c0[v] = 0.f;
    if (v < p.n_views) { const DyCoef d = dy_coef(p.v[v], c, p.inv_count); c0[v] = d.c0; c1 += d.c1; c2 += d.c2; }
  }
  for (int r = r0 + ry; r < r1; r += 8) {
    const int64_t o = (int64_t)r * p.dst.ld + c;
    float g = fmaf(c1, p.dst.raw[o], c2);
#pragma unroll
    for (int v = 0; v < kMaxViews; ++v)
      if (v < p.n_views) g = fmaf(c0[v], p.v[v].dz[(int64_t)r * p.v[v].ld + c], g);
    p.dst.dz[o] = p.accumulate ? p.dst.dz[o] + g : g;
  }
}

int launch_sumgrad(const SumGradLaunch& s, cudaStream_t st) {
  if (s.B <= 0) return SWR_OK;
  if (s.n_views <= 0 || s.n_views > kMaxViews) { set_error("sumgrad: %d views (max %d)", s.n_views, kMaxViews); return SWR_ERR_UNSUPPORTED; }
  if (!s.dst.raw || !s.dst.dz || is_lazy(s.dst)) { set_error("sumgrad: destination must be a plain activation with a gradient buffer"); return SWR_ERR_INVALID; }
  SumGradParams p{};
  p.dst = s.dst; p.n_views = s.n_views; p.accumulate = s.accumulate; p.B = (int)s.B; p.inv_count = 1.0f / (float)s.B;
  for (int v = 0; v < s.n_views; ++v) {
    p.v[v] = s.views[v];
    if (!s.views[v].dz || s.views[v].n != s.dst.n) { set_error("sumgrad: view %d malformed", v); return SWR_ERR_INVALID; }
  }
  dim3 grid(ceil_div(s.dst.n, 32), ceil_div(s.B, kRowsPerCta));
  sumgrad_kernel<<<grid, 256, 0, st>>>(p);
  SWR_LAUNCH_OK("sumgrad_kernel");
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// SELECT: out[b, :] = value(Y_{domain[b]})[b, :] (zero when the id is outside [0, D))
// ---------------------------------------------------------------------------------------
struct SelectParams { SelectLaunch l; float inv_count; };

__global__ void __launch_bounds__(256) select_fwd_kernel(const __grid_constant__ SelectParams p) {
  const SelectLaunch& L = p.l;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  if (c >= L.n) return;
  const int r0 = blockIdx.y * kRowsPerCta, r1 = min((int)L.B, r0 + kRowsPerCta);
  for (int r = r0 + ry; r < r1; r += 8) {
    const int64_t d = load_index(L.domain_id, L.dom_dtype, r);
    float v = 0.f;
    if (d >= 0 && d < L.n_domains) {
      const ActDev& Y = L.Y[d];
      v = act_value(Y.raw[(int64_t)r * Y.ld + c], col_coef(Y.norm, c, p.inv_count), Y.act);
    }
    L.out[(int64_t)r * L.ld_out + c] = v;
  }
}

// grid.z = domain
__global__ void __launch_bounds__(256) select_bwd_kernel(const __grid_constant__ SelectParams p) {
  const SelectLaunch& L = p.l;
  const int d = blockIdx.z;
  const ActDev& Y = L.Y[d];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool valid = c < L.n;
  const int r0 = blockIdx.y * kRowsPerCta, r1 = min((int)L.B, r0 + kRowsPerCta);
  const bool lazy = is_lazy(Y);
  double s1 = 0.0, s2 = 0.0;
  if (valid) {
    const ColCoef k = col_coef(Y.norm, c, p.inv_count);
    for (int r = r0 + ry; r < r1; r += 8) {
      const float g = (load_index(L.domain_id, L.dom_dtype, r) == d) ? L.dout[(int64_t)r * L.ld_out + c] : 0.f;
      grad_store(Y, k, lazy, (int64_t)r * Y.ld + c, g, false, s1, s2);
    }
  }
  if (Y.norm.mode != SWR_NORM_NONE) block_col_atomic(s1, s2, Y.dstats, c, valid);
}

static int check_select(const SelectLaunch& s) {
  if (s.n_domains <= 0 || s.n_domains > 16) { set_error("select: %d domains unsupported", s.n_domains); return SWR_ERR_UNSUPPORTED; }
  if (!s.out || !s.domain_id) { set_error("select: null operand"); return SWR_ERR_INVALID; }
  return SWR_OK;
}
int launch_select_fwd(const SelectLaunch& s, cudaStream_t st) {
  if (s.B <= 0) return SWR_OK;
  int rc = check_select(s); if (rc) return rc;
  SelectParams p{s, 1.0f / (float)s.B};
  dim3 grid(ceil_div(s.n, 32), ceil_div(s.B, kRowsPerCta));
  select_fwd_kernel<<<grid, 256, 0, st>>>(p);
  SWR_LAUNCH_OK("select_fwd_kernel");
  return SWR_OK;
}
int launch_select_bwd(const SelectLaunch& s, cudaStream_t st) {
  if (s.B <= 0) return SWR_OK;
  int rc = check_select(s); if (rc) return rc;
  if (!s.dout) { set_error("select_bwd: no output gradient"); return SWR_ERR_INVALID; }
  SelectParams p{s, 1.0f / (float)s.B};
  dim3 grid(ceil_div(s.n, 32), ceil_div(s.B, kRowsPerCta), s.n_domains);
  select_bwd_kernel<<<grid, 256, 0, st>>>(p);
  SWR_LAUNCH_OK("select_bwd_kernel");
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// LayerNorm (+ activation), one warp per row; n <= 512
// ---------------------------------------------------------------------------------------
constexpr int kLnGroups = 16;
constexpr int kLnMaxPerLane = 16;
struct LnParams { LnGroup g[kLnGroups]; int n_groups; int B; };

__global__ void __launch_bounds__(256) ln_fwd_kernel(const __grid_constant__ LnParams p) {
  const LnGroup& G = p.g[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= p.B) return;
  const float* y = G.y + (int64_t)row * G.ld_y;
  float v[kLnMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) { const int c = lane + 32 * i; v[i] = c < G.n ? y[c] : 0.f; s += v[i]; }
  const float mean = warp_sum(s) / (float)G.n;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) { const int c = lane + 32 * i; const float d = c < G.n ? v[i] - mean : 0.f; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) / (float)G.n + G.eps);
  if (lane == 0 && G.rowstats) { G.rowstats[2 * (int64_t)row] = mean; G.rowstats[2 * (int64_t)row + 1] = rstd; }
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    if (c < G.n) G.out[(int64_t)row * G.ld_out + c] = act_fwd(fmaf((v[i] - mean) * rstd, __ldg(G.gamma + c), __ldg(G.beta + c)), G.act);
  }
}

// each warp walks rows (stride = warps in the grid) and keeps per-lane column partials of d gamma / d beta
__global__ void __launch_bounds__(256) ln_bwd_kernel(const __grid_constant__ LnParams p) {
  const LnGroup& G = p.g[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * 8;
  float pg[kLnMaxPerLane], pb[kLnMaxPerLane], gam[kLnMaxPerLane], bet[kLnMaxPerLane];
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    pg[i] = 0.f; pb[i] = 0.f;
    gam[i] = c < G.n ? __ldg(G.gamma + c) : 0.f; bet[i] = c < G.n ? __ldg(G.beta + c) : 0.f;
  }
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < p.B; row += warps) {
    const float mean = G.rowstats[2 * (int64_t)row], rstd = G.rowstats[2 * (int64_t)row + 1];
    float xh[kLnMaxPerLane], g[kLnMaxPerLane];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      xh[i] = 0.f; g[i] = 0.f;
      if (c < G.n) {
        xh[i] = (G.y[(int64_t)row * G.ld_y + c] - mean) * rstd;
        const float dz = G.dout[(int64_t)row * G.ld_out + c] * act_grad(fmaf(xh[i], gam[i], bet[i]), G.act);
        pg[i] = fmaf(dz, xh[i], pg[i]); pb[i] += dz;
        g[i] = dz * gam[i];
        m1 += g[i]; m2 = fmaf(g[i], xh[i], m2);
      }
    }
    m1 = warp_sum(m1) / (float)G.n; m2 = warp_sum(m2) / (float)G.n;
    if (G.dy) {
#pragma unroll
      for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < G.n) G.dy[(int64_t)row * G.ld_y + c] = rstd * (g[i] - m1 - xh[i] * m2);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    if (c < G.n) {
      if (G.dgamma && pg[i] != 0.f) atomicAdd(G.dgamma + c, pg[i]);
      if (G.dbeta && pb[i] != 0.f) atomicAdd(G.dbeta + c, pb[i]);
    }
  }
}

static int fill_ln(const LnGroup* g, int n, int64_t B, LnParams& p) {
  if (n <= 0 || n > kLnGroups) { set_error("layernorm: %d groups (max %d per launch)", n, kLnGroups); return SWR_ERR_INVALID; }
  for (int i = 0; i < n; ++i) {
    p.g[i] = g[i];
    if (g[i].n <= 0 || g[i].n > 32 * kLnMaxPerLane) { set_error("layernorm: width %d unsupported (max %d)", g[i].n, 32 * kLnMaxPerLane); return SWR_ERR_UNSUPPORTED; }
    if (!g[i].y || !g[i].out || !g[i].gamma || !g[i].beta || !g[i].rowstats) { set_error("layernorm: null operand in group %d", i); return SWR_ERR_INVALID; }
  }
  p.n_groups = n; p.B = (int)B;
  return SWR_OK;
}
int launch_ln_fwd(const LnGroup* g, int n, int64_t B, cudaStream_t st) {
  if (B <= 0) return SWR_OK;
  for (int o = 0; o < n; o += kLnGroups) {
    LnParams p{};
    const int m = min(kLnGroups, n - o);
    int rc = fill_ln(g + o, m, B, p); if (rc) return rc;
    dim3 grid(ceil_div(B, 8), m);
    ln_fwd_kernel<<<grid, 256, 0, st>>>(p);
    SWR_LAUNCH_OK("ln_fwd_kernel");
  }
  return SWR_OK;
}
int launch_ln_bwd(const LnGroup* g, int n, int64_t B, cudaStream_t st) {
  if (B <= 0) return SWR_OK;
  for (int o = 0; o < n; o += kLnGroups) {
    LnParams p{};
    const int m = min(kLnGroups, n - o);
    int rc = fill_ln(g + o, m, B, p); if (rc) return rc;
    for (int i = 0; i < m; ++i) if (!p.g[i].dout) { set_error("layernorm_bwd: group %d has no output gradient", i); return SWR_ERR_INVALID; }
    dim3 grid(max(1, min(ceil_div(B, 8 * 16), 4 * 148)), m);
    ln_bwd_kernel<<<grid, 256, 0, st>>>(p);
    SWR_LAUNCH_OK("ln_bwd_kernel");
  }
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// MIX (m3oe.py:168-187): out_d += se * (sb * X_d + (1 - sb) / (D - 1) * sum_{j != d} X_j)
//   se = sigmoid(w_exp), sb = sigmoid(w_bal)
// ---------------------------------------------------------------------------------------
struct MixParams { MixLaunch m; };

__device__ __forceinline__ float sigm(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void __launch_bounds__(256) mix_fwd_kernel(const __grid_constant__ MixParams p) {
  const MixLaunch& M = p.m;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.B * M.H) return;
  const int64_t b = i / M.H; const int h = (int)(i - b * M.H);
  const float se = sigm(__ldg(M.w_exp)), sb = sigm(__ldg(M.w_bal));
  const float cc = M.D > 1 ? (1.f - sb) / (float)(M.D - 1) : 0.f;
  float x[16], sum = 0.f;
#pragma unroll
  for (int d = 0; d < 16; ++d) { x[d] = d < M.D ? M.X[d][b * M.ldx + h] : 0.f; sum += x[d]; }
#pragma unroll
  for (int d = 0; d < 16; ++d)
    if (d < M.D) M.out[d][b * M.ldo + h] += se * (sb * x[d] + cc * (sum - x[d]));
}

__global__ void __launch_bounds__(256) mix_bwd_kernel(const __grid_constant__ MixParams p) {
  const MixLaunch& M = p.m;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < M.B * M.H;
  const float se = sigm(__ldg(M.w_exp)), sb = sigm(__ldg(M.w_bal));
  const float cc = M.D > 1 ? (1.f - sb) / (float)(M.D - 1) : 0.f;
  double p1 = 0.0, p2 = 0.0;
  if (live) {
    const int64_t b = i / M.H; const int h = (int)(i - b * M.H);
    float x[16], g[16], sx = 0.f, sg = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      x[d] = d < M.D ? M.X[d][b * M.ldx + h] : 0.f; g[d] = d < M.D ? M.dout[d][b * M.ldo + h] : 0.f;
      sx += x[d]; sg += g[d];
    }
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      if (d < M.D) {
        if (M.dX[d]) M.dX[d][b * M.ldx + h] = se * (sb * g[d] + cc * (sg - g[d]));
        p1 += (double)g[d] * (double)x[d]; p2 += (double)g[d] * (double)(sx - x[d]);
      }
    }
  }
  if (M.red) {
    p1 = warp_sum(p1); p2 = warp_sum(p2);
    if ((threadIdx.x & 31) == 0) { atomicAdd(M.red, p1); atomicAdd(M.red + 1, p2); }
  }
}

__global__ void mix_finish_kernel(const __grid_constant__ MixParams p) {
  const MixLaunch& M = p.m;
  const float se = sigm(*M.w_exp), sb = sigm(*M.w_bal);
  const double P1 = M.red[0], P2 = M.red[1];
  const double cc = M.D > 1 ? (1.0 - sb) / (double)(M.D - 1) : 0.0;
  const double dse = sb * P1 + cc * P2;
  const double dsb = se * (P1 - (M.D > 1 ? P2 / (double)(M.D - 1) : 0.0));
  if (M.dw_exp) atomicAdd(M.dw_exp, (float)(dse * se * (1.f - se)));
  if (M.dw_bal) atomicAdd(M.dw_bal, (float)(dsb * sb * (1.f - sb)));
}

static int check_mix(const MixLaunch& m) {
  if (m.D <= 0 || m.D > 16) { set_error("mix: %d domains unsupported", m.D); return SWR_ERR_UNSUPPORTED; }
  if (!m.w_exp || !m.w_bal) { set_error("mix: null weights"); return SWR_ERR_INVALID; }
  return SWR_OK;
}
int launch_mix_fwd(const MixLaunch& m, cudaStream_t st) {
  if (m.B <= 0) return SWR_OK;
  int rc = check_mix(m); if (rc) return rc;
  MixParams p{m};
  mix_fwd_kernel<<<ceil_div(m.B * m.H, 256), 256, 0, st>>>(p);
  SWR_LAUNCH_OK("mix_fwd_kernel");
  return SWR_OK;
}
int launch_mix_bwd(const MixLaunch& m, cudaStream_t st) {
  if (m.B <= 0) return SWR_OK;
  int rc = check_mix(m); if (rc) return rc;
  MixParams p{m};
  mix_bwd_kernel<<<ceil_div(m.B * m.H, 256), 256, 0, st>>>(p);
  SWR_LAUNCH_OK("mix_bwd_kernel");
  if (m.red && (m.dw_exp || m.dw_bal)) {
    mix_finish_kernel<<<1, 1, 0, st>>>(p);
    SWR_LAUNCH_OK("mix_finish_kernel");
  }
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// BMV (hamur.py:177-189 re-associated): q_g[b, :] = p_g[b, :] * H_b for every group g sharing H.
// One CTA walks kBmvRows samples; H_b is staged once in shared memory per sample.
// ---------------------------------------------------------------------------------------
constexpr int kBmvRows = 8;
struct BmvParams { BmvLaunch m; };

__global__ void __launch_bounds__(128) bmv_fwd_kernel(const __grid_constant__ BmvParams p) {
  extern __shared__ __align__(16) float sm[];
  const BmvLaunch& M = p.m;
  const int k = M.k, kk = k * k;
  float* Hs = sm;              // [k*k]
  float* ps = sm + kk;         // [n_groups][k]
  const int64_t b0 = (int64_t)blockIdx.x * kBmvRows;
  for (int rr = 0; rr < kBmvRows; ++rr) {
    const int64_t b = b0 + rr;
    if (b >= M.B) break;
    __syncthreads();
    for (int i = threadIdx.x; i < kk; i += blockDim.x) Hs[i] = M.H[b * M.ldh + i];
    for (int i = threadIdx.x; i < M.n_groups * k; i += blockDim.x) { const int g = i / k, c = i - g * k; ps[i] = M.g[g].p[b * M.g[g].ldp + c]; }
    __syncthreads();
    for (int i = threadIdx.x; i < M.n_groups * k; i += blockDim.x) {
      const int g = i / k, j = i - g * k;
      float acc = 0.f;
      for (int r = 0; r < k; ++r) acc = fmaf(ps[g * k + r], Hs[r * k + j], acc);
      M.g[g].q[b * M.g[g].ldq + j] = acc;
    }
  }
}

// dp_g[b, i] = sum_j dq_g[b, j] H[b, i, j];   dH[b, i, j] (+)= sum_g p_g[b, i] dq_g[b, j]
__global__ void __launch_bounds__(128) bmv_bwd_kernel(const __grid_constant__ BmvParams p) {
  extern __shared__ __align__(16) float sm[];
  const BmvLaunch& M = p.m;
  const int k = M.k, kk = k * k;
  float* Hs = sm;                        // [k*k]
  float* ps = sm + kk;                   // [n_groups][k]
  float* qs = ps + M.n_groups * k;       // [n_groups][k]  (dq)
  const int64_t b0 = (int64_t)blockIdx.x * kBmvRows;
  for (int rr = 0; rr < kBmvRows; ++rr) {
    const int64_t b = b0 + rr;
    if (b >= M.B) break;
    __syncthreads();
    for (int i = threadIdx.x; i < kk; i += blockDim.x) Hs[i] = M.H[b * M.ldh + i];
    for (int i = threadIdx.x; i < M.n_groups * k; i += blockDim.x) {
      const int g = i / k, c = i - g * k;
      ps[i] = M.g[g].p[b * M.g[g].ldp + c];
      qs[i] = M.g[g].dq[b * M.g[g].ldq + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < M.n_groups * k; i += blockDim.x) {
      const int g = i / k, r = i - g * k;
      if (M.g[g].dp) {
        float acc = 0.f;
        for (int j = 0; j < k; ++j) acc = fmaf(qs[g * k + j], Hs[r * k + j], acc);
        M.g[g].dp[b * M.g[g].ldp + r] = acc;
      }
    }
    if (M.dH) {
      for (int i = threadIdx.x; i < kk; i += blockDim.x) {
        const int r = i / k, j = i - r * k;
        float acc = 0.f;
        for (int g = 0; g < M.n_groups; ++g) acc = fmaf(ps[g * k + r], qs[g * k + j], acc);
        float* d = M.dH + b * M.ldh + i;
        *d = M.accumulate_dH ? *d + acc : acc;
      }
    }
  }
}

static int check_bmv(const BmvLaunch& m, size_t& smem, bool bwd) {
  if (m.n_groups <= 0 || m.n_groups > 16) { set_error("bmv: %d groups unsupported", m.n_groups); return SWR_ERR_UNSUPPORTED; }
  if (m.k <= 0 || m.k > 128 || !m.H) { set_error("bmv: k = %d unsupported or null H", m.k); return SWR_ERR_UNSUPPORTED; }
  smem = sizeof(float) * ((size_t)m.k * m.k + (size_t)(bwd ? 2 : 1) * m.n_groups * m.k);
  return SWR_OK;
}
int launch_bmv_fwd(const BmvLaunch& m, cudaStream_t st) {
  if (m.B <= 0) return SWR_OK;
  size_t smem; int rc = check_bmv(m, smem, false); if (rc) return rc;
  if (smem > 48 * 1024) SWR_CUDA_OK(cudaFuncSetAttribute(bmv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BmvParams p{m};
  bmv_fwd_kernel<<<ceil_div(m.B, kBmvRows), 128, smem, st>>>(p);
  SWR_LAUNCH_OK("bmv_fwd_kernel");
  return SWR_OK;
}
int launch_bmv_bwd(const BmvLaunch& m, cudaStream_t st) {
  if (m.B <= 0) return SWR_OK;
  size_t smem; int rc = check_bmv(m, smem, true); if (rc) return rc;
  if (smem > 48 * 1024) SWR_CUDA_OK(cudaFuncSetAttribute(bmv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BmvParams p{m};
  bmv_bwd_kernel<<<ceil_div(m.B, kBmvRows), 128, smem, st>>>(p);
  SWR_LAUNCH_OK("bmv_bwd_kernel");
  return SWR_OK;
}

}  // namespace swr
