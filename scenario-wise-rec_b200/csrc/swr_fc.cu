// swr_fc.cu -- grouped fully-connected kernels of the expert / gate / domain-tower stack,
// fp32 SIMT path (the tcgen05 path for the wide expert layers lives in swr_fc_tc.cu).
//
// Three kernels share one tiled mainloop (BK = 16, 256 threads, register-prefetched
// operands, fp32 FFMA micro-tiles):
//   fc_fwd    Y = act(norm(A)) * Weff^T + beff      epilogue: optional activation, store raw
//                                                    Y, fp64 column moments of Y (BatchNorm)
//   fc_dgrad  dA = sum_g dY_g * Weff_g               A-operand dY = c0*dz + c1*raw + c2 formed
//                                                    on load (BatchNorm backward stage 2);
//                                                    epilogue: dz_dst = dA * act'(z_dst),
//                                                    fp64 column sums (stage 1 of dst's BN)
//   fc_wgrad  dWeff = dY^T * act(norm(A)), db        batch split across CTAs, fp32 atomics
// "Groups" are independent layers executed by one launch (all experts of a level, gates,
// towers); fc_dgrad instead sums its groups (every consumer of one activation).
#include "swr_common.cuh"
#include "swr_launch.h"

namespace swr {

constexpr int BK = 16;

template <int BM_, int BN_, int TM_, int TN_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, TM = TM_, TN = TN_;
  static constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static constexpr int LDA = BM + 4, LDB = BN + 4;
  static_assert(TX == 16 && NT == 256, "column reduction assumes 16 column-threads, 8 warps");
  static constexpr int A_IT = (BM * BK / 4 + NT - 1) / NT;
  static constexpr int B_IT = (BN * BK / 4 + NT - 1) / NT;
  static constexpr int SMEM_TILE = BK * (LDA + LDB);   // floats
};

struct FcParams {
  FcGroup g[kMaxGroups];
  int tile_start[kMaxGroups + 1];  // fwd/wgrad: first CTA tile of a group; dgrad: first k-tile of a group
  int n_groups;
  int B;
  float inv_count;
  int splits;          // wgrad: number of batch splits
  int rows_per_split;  // wgrad
  // dgrad: several destinations per launch; destination d owns groups [dst_group[d], dst_group[d+1]) (its fan-in)
  // and CTA tiles [dst_tile[d], dst_tile[d+1])
  int n_dst;
  int dst_group[kMaxGroups + 1];
  int dst_tile[kMaxGroups + 1];
};

template <class C>
__device__ __forceinline__ void mma_tile(const float* __restrict__ As, const float* __restrict__ Bs, int ty, int tx,
                                         float (&acc)[C::TM][C::TN], float* rowsum) {
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    float a[C::TM], b[C::TN];
#pragma unroll
    for (int i = 0; i < C::TM; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(As + kk * C::LDA + ty * C::TM + i);
      a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
    }
    if constexpr (C::TN % 4 == 0) {
#pragma unroll
      for (int j = 0; j < C::TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(Bs + kk * C::LDB + tx * C::TN + j);
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < C::TN; ++j) b[j] = Bs[kk * C::LDB + tx * C::TN + j];
    }
#pragma unroll
    for (int i = 0; i < C::TM; ++i)
#pragma unroll
      for (int j = 0; j < C::TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    if (rowsum) {
#pragma unroll
      for (int i = 0; i < C::TM; ++i) rowsum[i] += a[i];
    }
  }
}

// contraction-contiguous tile (ROWS outer rows x BK contraction) -> transposed smem [BK][LD]
template <int NT, int ROWS, int LD>
__device__ __forceinline__ void store_transposed(float* __restrict__ S, const float4* r, int tid) {
  constexpr int IT = (ROWS * BK / 4 + NT - 1) / NT;
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int v = it * NT + tid;
    if (v < ROWS * BK / 4) {
      const int row = v / (BK / 4), kq = v % (BK / 4);
      S[(4 * kq + 0) * LD + row] = r[it].x;
      S[(4 * kq + 1) * LD + row] = r[it].y;
      S[(4 * kq + 2) * LD + row] = r[it].z;
      S[(4 * kq + 3) * LD + row] = r[it].w;
    }
  }
}
// output-contiguous tile (BK contraction rows x COLS) -> smem [BK][LD] as is
template <int NT, int COLS, int LD>
__device__ __forceinline__ void store_direct(float* __restrict__ S, const float4* r, int tid) {
  constexpr int IT = (COLS * BK / 4 + NT - 1) / NT;
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int v = it * NT + tid;
    if (v < COLS * BK / 4) {
      const int kk = v / (COLS / 4), cq = v % (COLS / 4);
      *reinterpret_cast<float4*>(S + kk * LD + 4 * cq) = r[it];
    }
  }
}

// Reduce per-thread column partials over the BM rows of the CTA and add them to the global
// fp64 [n][2] statistics.  red: >= 2*8*BN doubles of shared memory (aliases the operand tiles).
template <class C>
__device__ __forceinline__ void col_reduce_atomic(double (&s1)[C::TN], double (&s2)[C::TN], double* red, double* gstats,
                                                  int n0, int N, int tid) {
  const int lane = tid & 31, warp = tid >> 5, tx = tid % C::TX;
#pragma unroll
  for (int j = 0; j < C::TN; ++j) {
    s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
    s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
  }
  __syncthreads();   // every thread is done with the operand tiles
  if (lane < 16) {
#pragma unroll
    for (int j = 0; j < C::TN; ++j) {
      red[(0 * 8 + warp) * C::BN + tx * C::TN + j] = s1[j];
      red[(1 * 8 + warp) * C::BN + tx * C::TN + j] = s2[j];
    }
  }
  __syncthreads();
  if (tid < 2 * C::BN) {
    const int which = tid / C::BN, col = tid % C::BN;
    if (n0 + col < N) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[(which * 8 + w) * C::BN + col];
      atomicAdd(gstats + 2 * (n0 + col) + which, t);
    }
  }
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(256) fc_fwd_kernel(const __grid_constant__ FcParams p) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + BK * C::LDA;
  float* kc = Bs + BK * C::LDB;   // [3][Kpad]: mu, s, b of the input columns
  const int tid = threadIdx.x, tx = tid % C::TX, ty = tid / C::TX;

  int g = 0;
  while (g + 1 < p.n_groups && p.tile_start[g + 1] <= (int)blockIdx.x) ++g;
  const FcGroup& G = p.g[g];
  const int M = p.B, N = G.Y.n, K = G.A.n;
  const int nt_n = (N + C::BN - 1) / C::BN;
  const int local = blockIdx.x - p.tile_start[g];
  const int m0 = (local / nt_n) * C::BM, n0 = (local % nt_n) * C::BN;
  const int Kpad = (K + BK - 1) / BK * BK;
  const bool plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
  const int actA = G.A.act;

  if (!plainA) {
    for (int k = tid; k < Kpad; k += C::NT) {
      ColCoef c = {0.f, 0.f, 0.f, 0.f};
      if (k < K) c = col_coef(G.A.norm, k, p.inv_count);
      kc[k] = c.mu; kc[Kpad + k] = c.s; kc[2 * Kpad + k] = c.b;
    }
    __syncthreads();
  }

  const bool vecA = is_al16(G.A.raw) && (G.A.ld % 4 == 0);
  const bool vecW = is_al16(G.W) && (G.ldw % 4 == 0) && (!G.W2 || is_al16(G.W2));
  const bool kn = (G.w_layout == SWR_W_KN);
  float4 ra[C::A_IT], rb[C::B_IT];

  auto loadA = [&](int k0) {
#pragma unroll
    for (int it = 0; it < C::A_IT; ++it) {
      const int v = it * C::NT + tid;
      if (v < C::BM * BK / 4) {
        const int r = v / (BK / 4), k = k0 + 4 * (v % (BK / 4));
        float4 x = load4_guard(G.A.raw, G.A.ld, m0 + r, k, M, K, vecA);
        if (!plainA) {   // padded k have s = b = mu = 0 -> contribute act(0) * 0 weight; rows >= M are never stored
          x.x = act_fwd(fmaf(x.x - kc[k], kc[Kpad + k], kc[2 * Kpad + k]), actA);
          x.y = act_fwd(fmaf(x.y - kc[k + 1], kc[Kpad + k + 1], kc[2 * Kpad + k + 1]), actA);
          x.z = act_fwd(fmaf(x.z - kc[k + 2], kc[Kpad + k + 2], kc[2 * Kpad + k + 2]), actA);
          x.w = act_fwd(fmaf(x.w - kc[k + 3], kc[Kpad + k + 3], kc[2 * Kpad + k + 3]), actA);
        }
        ra[it] = x;
      }
    }
  };
  auto loadB = [&](int k0) {
#pragma unroll
    for (int it = 0; it < C::B_IT; ++it) {
      const int v = it * C::NT + tid;
      if (v < C::BN * BK / 4) {
        float4 w;
        if (!kn) {   // nn.Linear [N, K]: contraction-contiguous
          const int r = v / (BK / 4), k = k0 + 4 * (v % (BK / 4));
          w = load4_guard(G.W, G.ldw, n0 + r, k, N, K, vecW);
          if (G.W2) { const float4 u = load4_guard(G.W2, G.ldw, n0 + r, k, N, K, vecW); w.x *= u.x; w.y *= u.y; w.z *= u.z; w.w *= u.w; }
        } else {     // STAR [K, N]: output-contiguous
          const int kk = v / (C::BN / 4), n = n0 + 4 * (v % (C::BN / 4));
          w = load4_guard(G.W, G.ldw, k0 + kk, n, K, N, vecW);
          if (G.W2) { const float4 u = load4_guard(G.W2, G.ldw, k0 + kk, n, K, N, vecW); w.x *= u.x; w.y *= u.y; w.z *= u.z; w.w *= u.w; }
        }
        rb[it] = w;
      }
    }
  };
  auto storeAB = [&]() {
    store_transposed<C::NT, C::BM, C::LDA>(As, ra, tid);
    if (!kn) store_transposed<C::NT, C::BN, C::LDB>(Bs, rb, tid);
    else store_direct<C::NT, C::BN, C::LDB>(Bs, rb, tid);
  };

  float acc[C::TM][C::TN];
#pragma unroll
  for (int i = 0; i < C::TM; ++i)
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j] = 0.f;

  const int nk = Kpad / BK;
  loadA(0); loadB(0);
  storeAB();
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) { loadA((kt + 1) * BK); loadB((kt + 1) * BK); }
    mma_tile<C>(As, Bs, ty, tx, acc, nullptr);
    __syncthreads();
    if (kt + 1 < nk) { storeAB(); __syncthreads(); }
  }

  // epilogue: bias, optional activation, store, column moments
  double s1[C::TN], s2[C::TN];
#pragma unroll
  for (int j = 0; j < C::TN; ++j) { s1[j] = 0.0; s2[j] = 0.0; }
  float bias[C::TN];
#pragma unroll
  for (int j = 0; j < C::TN; ++j) {
    const int n = n0 + tx * C::TN + j;
    bias[j] = (n < N) ? (ld_opt(G.bias, n, 0.f) + ld_opt(G.bias2, n, 0.f)) : 0.f;
  }
  float* Y = const_cast<float*>(G.Y.raw);
#pragma unroll
  for (int i = 0; i < C::TM; ++i) {
    const int m = m0 + ty * C::TM + i;
    if (m < M) {
#pragma unroll
      for (int j = 0; j < C::TN; ++j) {
        const int n = n0 + tx * C::TN + j;
        if (n < N) {
          float y = acc[i][j] + bias[j];
          if (G.e_act != SWR_ACT_NONE) y = act_fwd(y, G.e_act) * G.e_scale;
          Y[(int64_t)m * G.Y.ld + n] = y;
          s1[j] += (double)y; s2[j] += (double)y * (double)y;
        }
      }
    }
  }
  if (G.stats_out) col_reduce_atomic<C>(s1, s2, reinterpret_cast<double*>(smem), G.stats_out, n0, N, tid);
}

// ---------------------------------------------------------------------------------------
// data gradient (fan-in over groups)
// ---------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(256) fc_dgrad_kernel(const __grid_constant__ FcParams p) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + BK * C::LDA;
  float* dc = Bs + BK * C::LDB;   // [3][Kc]: c0, c1, c2 over the concatenated (BK-padded) group columns of this destination
  const int tid = threadIdx.x, tx = tid % C::TX, ty = tid / C::TX;
  int d = 0;
  while (d + 1 < p.n_dst && p.dst_tile[d + 1] <= (int)blockIdx.x) ++d;
  const int gs = p.dst_group[d], ge = p.dst_group[d + 1];
  const ActDev& D = p.g[gs].A;    // destination activation (shared by the groups of this fan-in)
  const int M = p.B, Kd = D.n;
  const int nt_n = (Kd + C::BN - 1) / C::BN;
  const int local = blockIdx.x - p.dst_tile[d];
  const int m0 = (local / nt_n) * C::BM, j0 = (local % nt_n) * C::BN;
  const int kt0 = p.tile_start[gs];
  const int nk = p.tile_start[ge] - kt0;
  const int Kc = nk * BK;

  for (int g = gs; g < ge; ++g) {
    const FcGroup& G = p.g[g];
    const int base = (p.tile_start[g] - kt0) * BK, span = (p.tile_start[g + 1] - p.tile_start[g]) * BK;
    for (int n = tid; n < span; n += C::NT) {
      DyCoef c = {0.f, 0.f, 0.f};
      if (n < G.Y.n) c = dy_coef(G.Y, n, p.inv_count);
      dc[base + n] = c.c0; dc[Kc + base + n] = c.c1; dc[2 * Kc + base + n] = c.c2;
    }
  }
  __syncthreads();

  float4 ra[C::A_IT], rb[C::B_IT];
  int cur_g = gs;
  bool kn = false;
  auto loadAB = [&](int kt) {     // kt: k-tile index inside this destination's fan-in
    while (cur_g + 1 < ge && p.tile_start[cur_g + 1] - kt0 <= kt) ++cur_g;
    const FcGroup& G = p.g[cur_g];
    const int N = G.Y.n;
    const int nl0 = (kt - (p.tile_start[cur_g] - kt0)) * BK;   // first column of this k-tile inside the group
    const int cb = kt * BK;                                    // same position in the coefficient arrays
    const bool vecY = is_al16(G.Y.dz) && is_al16(G.Y.raw) && (G.Y.ld % 4 == 0);
    const bool vecW = is_al16(G.W) && (G.ldw % 4 == 0) && (!G.W2 || is_al16(G.W2));
    const bool need_raw = (G.Y.norm.mode == SWR_NORM_BATCH);
    kn = (G.w_layout == SWR_W_KN);
#pragma unroll
    for (int it = 0; it < C::A_IT; ++it) {
      const int v = it * C::NT + tid;
      if (v < C::BM * BK / 4) {
        const int r = v / (BK / 4), q = 4 * (v % (BK / 4));
        const float4 dz = load4_guard(G.Y.dz, G.Y.ld, m0 + r, nl0 + q, M, N, vecY);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + r < M) {   // c2 != 0, so rows outside the batch must stay exactly zero
          float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
          if (need_raw) raw = load4_guard(G.Y.raw, G.Y.ld, m0 + r, nl0 + q, M, N, vecY);
          const float* c0 = dc + cb + q; const float* c1 = c0 + Kc; const float* c2 = c1 + Kc;
          x.x = fmaf(c0[0], dz.x, fmaf(c1[0], raw.x, c2[0]));
          x.y = fmaf(c0[1], dz.y, fmaf(c1[1], raw.y, c2[1]));
          x.z = fmaf(c0[2], dz.z, fmaf(c1[2], raw.z, c2[2]));
          x.w = fmaf(c0[3], dz.w, fmaf(c1[3], raw.w, c2[3]));
        }
        ra[it] = x;
      }
    }
#pragma unroll
    for (int it = 0; it < C::B_IT; ++it) {
      const int v = it * C::NT + tid;
      if (v < C::BN * BK / 4) {
        float4 w;
        if (!kn) {   // W[n, j]: output(j)-contiguous
          const int kk = v / (C::BN / 4), j = j0 + 4 * (v % (C::BN / 4));
          w = load4_guard(G.W, G.ldw, nl0 + kk, j, N, Kd, vecW);
          if (G.W2) { const float4 u = load4_guard(G.W2, G.ldw, nl0 + kk, j, N, Kd, vecW); w.x *= u.x; w.y *= u.y; w.z *= u.z; w.w *= u.w; }
        } else {     // W[j, n]: contraction(n)-contiguous
          const int r = v / (BK / 4), q = 4 * (v % (BK / 4));
          w = load4_guard(G.W, G.ldw, j0 + r, nl0 + q, Kd, N, vecW);
          if (G.W2) { const float4 u = load4_guard(G.W2, G.ldw, j0 + r, nl0 + q, Kd, N, vecW); w.x *= u.x; w.y *= u.y; w.z *= u.z; w.w *= u.w; }
        }
        rb[it] = w;
      }
    }
  };
  auto storeAB = [&]() {
    store_transposed<C::NT, C::BM, C::LDA>(As, ra, tid);
    if (!kn) store_direct<C::NT, C::BN, C::LDB>(Bs, rb, tid);
    else store_transposed<C::NT, C::BN, C::LDB>(Bs, rb, tid);
  };

  float acc[C::TM][C::TN];
#pragma unroll
  for (int i = 0; i < C::TM; ++i)
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j] = 0.f;

  loadAB(0);
  storeAB();
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) loadAB(kt + 1);
    mma_tile<C>(As, Bs, ty, tx, acc, nullptr);
    __syncthreads();
    if (kt + 1 < nk) { storeAB(); __syncthreads(); }
  }

  // epilogue: backward through the destination's activation / norm (stage 1), store dz
  const bool accumulate = (p.g[gs].flags & FC_A_ACCUMULATE) != 0;
  const bool has_norm = D.norm.mode != SWR_NORM_NONE;
  const bool plainD = !has_norm && D.act == SWR_ACT_NONE;
  double s1[C::TN], s2[C::TN];
  ColCoef cc[C::TN];
#pragma unroll
  for (int j = 0; j < C::TN; ++j) {
    s1[j] = 0.0; s2[j] = 0.0;
    const int col = j0 + tx * C::TN + j;
    cc[j] = ColCoef{0.f, 1.f, 0.f, 1.f};
    if (!plainD && col < Kd) cc[j] = col_coef(D.norm, col, p.inv_count);
  }
#pragma unroll
  for (int i = 0; i < C::TM; ++i) {
    const int m = m0 + ty * C::TM + i;
    if (m < M) {
#pragma unroll
      for (int j = 0; j < C::TN; ++j) {
        const int col = j0 + tx * C::TN + j;
        if (col < Kd) {
          const int64_t o = (int64_t)m * D.ld + col;
          float dz = acc[i][j];
          if (!plainD) {
            const float raw = D.raw[o];
            dz *= act_grad(fmaf(raw - cc[j].mu, cc[j].s, cc[j].b), D.act);
            s1[j] += (double)dz; s2[j] += (double)dz * (double)((raw - cc[j].mu) * cc[j].r);
          }
          if (accumulate) dz += D.dz[o];
          D.dz[o] = dz;
        }
      }
    }
  }
  if (has_norm && D.dstats) col_reduce_atomic<C>(s1, s2, reinterpret_cast<double*>(smem), D.dstats, j0, Kd, tid);
}

// ---------------------------------------------------------------------------------------
// weight / bias gradient
// ---------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(256) fc_wgrad_kernel(const __grid_constant__ FcParams p) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + BK * C::LDA;
  float* mc = Bs + BK * C::LDB;   // [3][BM]: c0,c1,c2 of the dY rows (output features n)
  float* nc = mc + 3 * C::BM;     // [3][BN]: mu,s,b of the input columns j
  const int tid = threadIdx.x, tx = tid % C::TX, ty = tid / C::TX;

  int g = 0;
  while (g + 1 < p.n_groups && p.tile_start[g + 1] <= (int)blockIdx.x) ++g;
  const FcGroup& G = p.g[g];
  const int N = G.Y.n, K = G.A.n;
  const int nt_m = (N + C::BM - 1) / C::BM, nt_n = (K + C::BN - 1) / C::BN;
  int local = blockIdx.x - p.tile_start[g];
  const int split = local / (nt_m * nt_n);
  local -= split * nt_m * nt_n;
  const int m0 = (local / nt_n) * C::BM, j0 = (local % nt_n) * C::BN;
  const int b_begin = split * p.rows_per_split;
  const int b_end = min(p.B, b_begin + p.rows_per_split);
  if (b_begin >= b_end) return;

  const bool plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
  for (int i = tid; i < C::BM; i += C::NT) {
    DyCoef c = {0.f, 0.f, 0.f};
    if (m0 + i < N) c = dy_coef(G.Y, m0 + i, p.inv_count);
    mc[i] = c.c0; mc[C::BM + i] = c.c1; mc[2 * C::BM + i] = c.c2;
  }
  for (int i = tid; i < C::BN; i += C::NT) {
    ColCoef c = {0.f, 0.f, 0.f, 0.f};
    if (j0 + i < K) c = plainA ? ColCoef{0.f, 1.f, 0.f, 1.f} : col_coef(G.A.norm, j0 + i, p.inv_count);
    nc[i] = c.mu; nc[C::BN + i] = c.s; nc[2 * C::BN + i] = c.b;
  }
  __syncthreads();

  const bool vecY = is_al16(G.Y.dz) && is_al16(G.Y.raw) && (G.Y.ld % 4 == 0);
  const bool vecA = is_al16(G.A.raw) && (G.A.ld % 4 == 0);
  const bool need_raw = (G.Y.norm.mode == SWR_NORM_BATCH);
  const int actA = G.A.act;
  float4 ra[C::A_IT], rb[C::B_IT];
  auto loadAB = [&](int b0) {
#pragma unroll
    for (int it = 0; it < C::A_IT; ++it) {   // A(n, b) = dY[b, n]: output(n)-contiguous
      const int v = it * C::NT + tid;
      if (v < C::BM * BK / 4) {
        const int kk = v / (C::BM / 4), q = 4 * (v % (C::BM / 4));
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + kk < b_end) {
          const float4 dz = load4_guard(G.Y.dz, G.Y.ld, b0 + kk, m0 + q, b_end, N, vecY);
          float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
          if (need_raw) raw = load4_guard(G.Y.raw, G.Y.ld, b0 + kk, m0 + q, b_end, N, vecY);
          const float* c0 = mc + q; const float* c1 = c0 + C::BM; const float* c2 = c1 + C::BM;
          x.x = fmaf(c0[0], dz.x, fmaf(c1[0], raw.x, c2[0]));
          x.y = fmaf(c0[1], dz.y, fmaf(c1[1], raw.y, c2[1]));
          x.z = fmaf(c0[2], dz.z, fmaf(c1[2], raw.z, c2[2]));
          x.w = fmaf(c0[3], dz.w, fmaf(c1[3], raw.w, c2[3]));
        }
        ra[it] = x;
      }
    }
#pragma unroll
    for (int it = 0; it < C::B_IT; ++it) {   // B(b, j) = act(norm(A))[b, j]: output(j)-contiguous
      const int v = it * C::NT + tid;
      if (v < C::BN * BK / 4) {
        const int kk = v / (C::BN / 4), q = 4 * (v % (C::BN / 4));
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + kk < b_end) {
          x = load4_guard(G.A.raw, G.A.ld, b0 + kk, j0 + q, b_end, K, vecA);
          if (!plainA) {   // columns >= K have s = b = 0 and act(0) = 0 for relu/none; masked at the store anyway
            const float* mu = nc + q; const float* s = mu + C::BN; const float* b = s + C::BN;
            x.x = act_fwd(fmaf(x.x - mu[0], s[0], b[0]), actA);
            x.y = act_fwd(fmaf(x.y - mu[1], s[1], b[1]), actA);
            x.z = act_fwd(fmaf(x.z - mu[2], s[2], b[2]), actA);
            x.w = act_fwd(fmaf(x.w - mu[3], s[3], b[3]), actA);
          }
        }
        rb[it] = x;
      }
    }
  };
  auto storeAB = [&]() {
    store_direct<C::NT, C::BM, C::LDA>(As, ra, tid);
    store_direct<C::NT, C::BN, C::LDB>(Bs, rb, tid);
  };

  float acc[C::TM][C::TN];
  float rowsum[C::TM];
#pragma unroll
  for (int i = 0; i < C::TM; ++i) {
    rowsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j] = 0.f;
  }
  const bool do_bias = (j0 == 0) && (tx == 0) && (G.dbias || G.dbias2);

  loadAB(b_begin);
  storeAB();
  __syncthreads();
  for (int b0 = b_begin; b0 < b_end; b0 += BK) {
    const bool more = b0 + BK < b_end;
    if (more) loadAB(b0 + BK);
    mma_tile<C>(As, Bs, ty, tx, acc, do_bias ? rowsum : nullptr);
    __syncthreads();
    if (more) { storeAB(); __syncthreads(); }
  }

  const bool kn = (G.w_layout == SWR_W_KN);
#pragma unroll
  for (int i = 0; i < C::TM; ++i) {
    const int n = m0 + ty * C::TM + i;
    if (n < N) {
#pragma unroll
      for (int j = 0; j < C::TN; ++j) {
        const int col = j0 + tx * C::TN + j;
        if (col < K) {
          const int64_t o = kn ? ((int64_t)col * G.ldw + n) : ((int64_t)n * G.ldw + col);
          const float v = acc[i][j];
          if (G.W2) {
            if (G.dW) atomicAdd(G.dW + o, v * __ldg(G.W2 + o));
            if (G.dW2) atomicAdd(G.dW2 + o, v * __ldg(G.W + o));
          } else if (G.dW) {
            atomicAdd(G.dW + o, v);
          }
        }
      }
      if (do_bias) {
        if (G.dbias) atomicAdd(G.dbias + n, rowsum[i]);
        if (G.dbias2) atomicAdd(G.dbias2 + n, rowsum[i]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
using CfgWide = TileCfg<128, 64, 8, 4>;
using CfgMid = TileCfg<64, 64, 4, 4>;
using CfgNarrow = TileCfg<128, 16, 8, 1>;
using CfgNarrowS = TileCfg<64, 16, 4, 1>;

template <class K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 200 * 1024) { set_error("fc: %zu bytes of shared memory needed", bytes); return SWR_ERR_UNSUPPORTED; }
    SWR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return SWR_OK;
}

static int check_groups(const FcGroup* groups, int n_groups, const char* what) {
  if (n_groups <= 0 || n_groups > kMaxGroups) { set_error("%s: %d groups (max %d per launch)", what, n_groups, kMaxGroups); return SWR_ERR_INVALID; }
  for (int g = 0; g < n_groups; ++g) {
    if (!groups[g].A.raw || !groups[g].Y.raw || !groups[g].W) { set_error("%s: null operand in group %d", what, g); return SWR_ERR_INVALID; }
    if (groups[g].A.n <= 0 || groups[g].Y.n <= 0) { set_error("%s: empty layer in group %d", what, g); return SWR_ERR_INVALID; }
  }
  return SWR_OK;
}

template <class C>
static int run_fwd(FcParams& p, cudaStream_t st) {
  int tiles = 0, kmax = 0;
  for (int g = 0; g < p.n_groups; ++g) {
    p.tile_start[g] = tiles;
    tiles += ceil_div(p.B, C::BM) * ceil_div(p.g[g].Y.n, C::BN);
    kmax = max(kmax, p.g[g].A.n);
  }
  p.tile_start[p.n_groups] = tiles;
  const size_t sm = sizeof(float) * (C::SMEM_TILE + 3 * (size_t)((kmax + BK - 1) / BK * BK));
  const size_t need = max(sm, sizeof(double) * 16 * C::BN);
  int rc = set_smem(fc_fwd_kernel<C>, need);
  if (rc) return rc;
  fc_fwd_kernel<C><<<tiles, 256, need, st>>>(p);
  SWR_LAUNCH_OK("fc_fwd_kernel");
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------
// skinny layers (N <= 8 outputs: the MMoE / PLE gates, Linear(IN, n_expert), mmoe.py:28, ple.py:98-105)
// A tiled GEMM wastes its tile on 4 output columns and pays a latency-bound k-loop per CTA; this forward kernel
// reads each activation row once with 16-byte loads and keeps the whole (effective) weight in shared memory.
// (A matching wgrad was tried and lost to the tiled kernel: 128 CTAs' atomics per weight element.)
// ---------------------------------------------------------------------------------------
constexpr int kSkinnyN = 8;
constexpr int kSkinnyK = 1024;
constexpr int kSkinnyRows = 32;     // forward: rows per CTA (4 per warp)

__global__ void __launch_bounds__(256) fc_skinny_fwd_kernel(const __grid_constant__ FcParams p) {
  extern __shared__ __align__(16) float sk[];
  const FcGroup& G = p.g[blockIdx.y];
  const int M = p.B, N = G.Y.n, K = G.A.n, Kp = (K + 3) & ~3;
  float* Ws = sk;                    // [N][Kp] effective weight
  float* kc = Ws + kSkinnyN * Kp;    // [3][Kp] mu, s, b of the input columns
  __shared__ double sst[2 * kSkinnyN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
  const bool kn = (G.w_layout == SWR_W_KN);
  for (int i = tid; i < N * Kp; i += 256) {
    const int n = i / Kp, k = i - n * Kp;
    float w = 0.f;
    if (k < K) {
      const int64_t o = kn ? ((int64_t)k * G.ldw + n) : ((int64_t)n * G.ldw + k);
      w = __ldg(G.W + o);
      if (G.W2) w *= __ldg(G.W2 + o);
    }
    Ws[i] = w;
  }
  if (!plainA)
    for (int k = tid; k < Kp; k += 256) {
      ColCoef c = {0.f, 0.f, 0.f, 0.f};
      if (k < K) c = col_coef(G.A.norm, k, p.inv_count);
      kc[k] = c.mu; kc[Kp + k] = c.s; kc[2 * Kp + k] = c.b;
    }
  if (tid < 2 * kSkinnyN) sst[tid] = 0.0;
  __syncthreads();
  const bool vecA = is_al16(G.A.raw) && (G.A.ld % 4 == 0);
  const int actA = G.A.act;
  float bias = 0.f;
  if (lane < N) bias = ld_opt(G.bias, lane, 0.f) + ld_opt(G.bias2, lane, 0.f);
  float s1 = 0.f, s2 = 0.f;          // lane n: moments of output column n over this warp's rows
  float* Y = const_cast<float*>(G.Y.raw);
  for (int j = 0; j < kSkinnyRows / 8; ++j) {
    const int m = blockIdx.x * kSkinnyRows + warp * (kSkinnyRows / 8) + j;
    if (m >= M) break;               // warp-uniform
    float acc[kSkinnyN];
#pragma unroll
    for (int n = 0; n < kSkinnyN; ++n) acc[n] = 0.f;
    for (int k4 = 4 * lane; k4 < Kp; k4 += 128) {
      float4 x = load4_guard(G.A.raw, G.A.ld, m, k4, M, K, vecA);
      if (!plainA) {
        x.x = act_fwd(fmaf(x.x - kc[k4], kc[Kp + k4], kc[2 * Kp + k4]), actA);
        x.y = act_fwd(fmaf(x.y - kc[k4 + 1], kc[Kp + k4 + 1], kc[2 * Kp + k4 + 1]), actA);
        x.z = act_fwd(fmaf(x.z - kc[k4 + 2], kc[Kp + k4 + 2], kc[2 * Kp + k4 + 2]), actA);
        x.w = act_fwd(fmaf(x.w - kc[k4 + 3], kc[Kp + k4 + 3], kc[2 * Kp + k4 + 3]), actA);
      }
#pragma unroll
      for (int n = 0; n < kSkinnyN; ++n)
        if (n < N) {   // padded k carry a zero weight
          const float4 w = *reinterpret_cast<const float4*>(Ws + n * Kp + k4);
          acc[n] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[n]))));
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < kSkinnyN; ++n)
      if (n < N) {
        const float t = warp_sum(acc[n]);
        if (lane == n) mine = t;
      }
    if (lane < N) {
      float y = mine + bias;
      if (G.e_act != SWR_ACT_NONE) y = act_fwd(y, G.e_act) * G.e_scale;
      Y[(int64_t)m * G.Y.ld + lane] = y;
      s1 += y; s2 = fmaf(y, y, s2);
    }
  }
  if (G.stats_out) {
    if (lane < N) { atomicAdd(&sst[2 * lane], (double)s1); atomicAdd(&sst[2 * lane + 1], (double)s2); }
    __syncthreads();
    if (tid < 2 * N) atomicAdd(G.stats_out + tid, sst[tid]);
  }
}

static bool skinny_ok(const FcGroup* groups, int n_groups) {
  for (int g = 0; g < n_groups; ++g)
    if (groups[g].Y.n > kSkinnyN || groups[g].A.n > kSkinnyK) return false;
  return n_groups > 0;
}

static int launch_fc_skinny_fwd(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  FcParams p{};
  int kmax = 0;
  for (int g = 0; g < n_groups; ++g) { p.g[g] = groups[g]; kmax = max(kmax, groups[g].A.n); }
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  const size_t sm = sizeof(float) * (size_t)((kmax + 3) & ~3) * (kSkinnyN + 3);
  int rc = set_smem(fc_skinny_fwd_kernel, sm);
  if (rc) return rc;
  fc_skinny_fwd_kernel<<<dim3(ceil_div(B, kSkinnyRows), n_groups), 256, sm, st>>>(p);
  SWR_LAUNCH_OK("fc_skinny_fwd_kernel");
  return SWR_OK;
}

static int launch_fc_fwd_simt(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st);
static int launch_fc_wgrad_simt(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st);

int launch_fc_fwd(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  if (B <= 0) return SWR_OK;
  int rc = check_groups(groups, n_groups, "fc_fwd");
  if (rc) return rc;
  if (fc_tc_wanted(groups, n_groups, B) && fc_tc2_usable(groups, n_groups, 0)) {
    rc = launch_fc_tc2_fwd(groups, n_groups, B, st);     // every group of the level, narrow ones included, in one launch
    if (rc != SWR_ERR_UNSUPPORTED) return rc;
  }
  // shapes the tensor-core kernels do not take (unaligned column sub-views): FFMA serves them
  // (the skinny kernel pays off for gate-shaped layers: few outputs over a long contraction)
  bool gate_like = skinny_ok(groups, n_groups);
  for (int g = 0; g < n_groups; ++g) gate_like = gate_like && groups[g].A.n >= 128;
  return gate_like ? launch_fc_skinny_fwd(groups, n_groups, B, st) : launch_fc_fwd_simt(groups, n_groups, B, st);
}

static int launch_fc_fwd_simt(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  FcParams p{};
  int nmax = 0; int64_t cols = 0;
  for (int g = 0; g < n_groups; ++g) { p.g[g] = groups[g]; nmax = max(nmax, groups[g].Y.n); cols += groups[g].Y.n; }
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  if (nmax > 16) {
    const int64_t tiles128 = (int64_t)ceil_div(B, 128) * ((cols + 63) / 64);
    return tiles128 >= 296 ? run_fwd<CfgWide>(p, st) : run_fwd<CfgMid>(p, st);
  }
  return ceil_div(B, 128) * n_groups >= 296 ? run_fwd<CfgNarrow>(p, st) : run_fwd<CfgNarrowS>(p, st);
}

template <class C>
static int run_dgrad(FcParams& p, cudaStream_t st) {
  int kt = 0;
  for (int g = 0; g < p.n_groups; ++g) { p.tile_start[g] = kt; kt += ceil_div(p.g[g].Y.n, BK); }
  p.tile_start[p.n_groups] = kt;
  int tiles = 0, kmax = 0;
  for (int d = 0; d < p.n_dst; ++d) {
    p.dst_tile[d] = tiles;
    tiles += ceil_div(p.B, C::BM) * ceil_div(p.g[p.dst_group[d]].A.n, C::BN);
    kmax = max(kmax, p.tile_start[p.dst_group[d + 1]] - p.tile_start[p.dst_group[d]]);
  }
  p.dst_tile[p.n_dst] = tiles;
  const size_t sm = sizeof(float) * (C::SMEM_TILE + 3 * (size_t)kmax * BK);
  const size_t need = max(sm, sizeof(double) * 16 * C::BN);
  int rc = set_smem(fc_dgrad_kernel<C>, need);
  if (rc) return rc;
  fc_dgrad_kernel<C><<<tiles, 256, need, st>>>(p);
  SWR_LAUNCH_OK("fc_dgrad_kernel");
  return SWR_OK;
}

// groups must be sorted by destination: dst_of[g] is non-decreasing, groups of one destination share A
int launch_fc_dgrad(const FcGroup* groups, const int* dst_of, int n_groups, int64_t B, cudaStream_t st) {
  if (B <= 0) return SWR_OK;
  int rc = check_groups(groups, n_groups, "fc_dgrad");
  if (rc) return rc;
  FcParams p{};
  int n_dst = 0, kd_max = 0;
  int64_t tiles128 = 0;
  for (int g = 0; g < n_groups; ++g) {
    p.g[g] = groups[g];
    if (g == 0 || dst_of[g] != dst_of[g - 1]) {
      if (g > 0 && dst_of[g] < dst_of[g - 1]) { set_error("fc_dgrad: groups are not sorted by destination"); return SWR_ERR_INVALID; }
      p.dst_group[n_dst++] = g;
      if (!groups[g].A.dz) { set_error("fc_dgrad: destination has no gradient buffer"); return SWR_ERR_INVALID; }
      kd_max = max(kd_max, groups[g].A.n);
      tiles128 += (int64_t)ceil_div(B, 128) * ((groups[g].A.n + 63) / 64);
    } else if (groups[g].A.raw != groups[g - 1].A.raw || groups[g].A.n != groups[g - 1].A.n) {
      set_error("fc_dgrad: groups of one destination must share their input activation"); return SWR_ERR_INVALID;
    }
    if (!groups[g].Y.dz) { set_error("fc_dgrad: group %d has no output gradient buffer", g); return SWR_ERR_INVALID; }
  }
  p.dst_group[n_dst] = n_groups;
  p.n_dst = n_dst;
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  if (fc_tc_wanted(groups, n_groups, B) && fc_tc2_usable(groups, n_groups, 1)) {
    rc = launch_fc_tc2_dgrad(groups, p.dst_group, n_dst, n_groups, B, st);
    if (rc != SWR_ERR_UNSUPPORTED) return rc;
  }
  if (kd_max > 16) return tiles128 >= 296 ? run_dgrad<CfgWide>(p, st) : run_dgrad<CfgMid>(p, st);
  return run_dgrad<CfgNarrowS>(p, st);
}

template <class C>
static int run_wgrad(FcParams& p, cudaStream_t st) {
  int base = 0;
  for (int g = 0; g < p.n_groups; ++g) base += ceil_div(p.g[g].Y.n, C::BM) * ceil_div(p.g[g].A.n, C::BN);
  // split the batch so the grid covers the machine about six times; keep >= 64 rows (4 k-steps) per split: the
  // narrow layers this kernel serves are latency-bound per k-step, so short chains on many CTAs win
  int splits = max(1, min((6 * 148 + base - 1) / base, (p.B + 63) / 64));
  int rows = (p.B + splits - 1) / splits;
  rows = (rows + BK - 1) / BK * BK;
  splits = (p.B + rows - 1) / rows;
  p.splits = splits; p.rows_per_split = rows;
  int tiles = 0;
  for (int g = 0; g < p.n_groups; ++g) {
    p.tile_start[g] = tiles;
    tiles += ceil_div(p.g[g].Y.n, C::BM) * ceil_div(p.g[g].A.n, C::BN) * splits;
  }
  p.tile_start[p.n_groups] = tiles;
  const size_t sm = sizeof(float) * (C::SMEM_TILE + 3 * C::BM + 3 * C::BN);
  fc_wgrad_kernel<C><<<tiles, 256, sm, st>>>(p);
  SWR_LAUNCH_OK("fc_wgrad_kernel");
  return SWR_OK;
}

int launch_fc_wgrad(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  if (B <= 0) return SWR_OK;
  int rc = check_groups(groups, n_groups, "fc_wgrad");
  if (rc) return rc;
  for (int g = 0; g < n_groups; ++g)
    if (!groups[g].Y.dz) { set_error("fc_wgrad: group %d has no output gradient buffer", g); return SWR_ERR_INVALID; }
  if (fc_tc_wanted(groups, n_groups, B) && fc_tc2_usable(groups, n_groups, 2)) {
    rc = launch_fc_tc2_wgrad(groups, n_groups, B, st);
    if (rc != SWR_ERR_UNSUPPORTED) return rc;
  }
  return launch_fc_wgrad_simt(groups, n_groups, B, st);
}

static int launch_fc_wgrad_simt(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  FcParams p{};
  int kmax = 0;
  for (int g = 0; g < n_groups; ++g) {
    p.g[g] = groups[g]; kmax = max(kmax, groups[g].A.n);
    if (!groups[g].Y.dz) { set_error("fc_wgrad: group %d has no output gradient buffer", g); return SWR_ERR_INVALID; }
  }
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  return kmax > 16 ? run_wgrad<CfgMid>(p, st) : run_wgrad<CfgNarrowS>(p, st);
}

}  // namespace swr
