// swr_fc_tc2.cu -- grouped fully-connected kernels on the sm_100a tensor cores, second generation:
// persistent, warp-specialised, TMA-fed.
//
// Same contract as the FFMA kernels in swr_fc.cu (fc_fwd / fc_dgrad / fc_wgrad over FcGroup lists; reference:
// basic/layers.py:253-258 Linear -> BatchNorm1d -> act, star.py:103-110, ppnet.py:21-29, hamur.py adapters,
// m3oe.py:45-68).  Arithmetic: 3xTF32 (x = hi + lo, three tcgen05.mma.kind::tf32 per k-step) as before.
//
// What changed against the first generation (swr_fc_tc_v1.cu):
//   * Weights are split ONCE per forward pass by fc_presplit_kernel into "images": the effective weight
//     W (.) W2 as hi / lo TF32 planes, in both orientations (contraction-contiguous for the forward and for
//     the data gradient).  The kernels bring weight tiles in with cp.async.bulk.tensor (TMA, SWIZZLE_128B):
//     no thread touches a weight any more.
//   * Warp roles.  warps 0-7 stage the activation-side operand (lazy BatchNorm + activation, or the
//     BatchNorm-backward affine map, + hi/lo split) from registers into TMEM (tcgen05.st); warps 8-15 are the
//     epilogue; warp 16 lane 0 issues TMA; warp 17 lane 0 issues every tcgen05.mma.  Five mbarrier rings connect
//     them (weight stage full, operand stage full, stage free, accumulator full, accumulator free).
//   * Persistent CTAs: each CTA walks a contiguous range of output tiles; two accumulator buffers in TMEM let
//     the epilogue of one tile overlap the MMAs of the next.
//   * Accumulator flushes.  tcgen05 accumulates with round-toward-zero, a bias that grows with the length of
//     the accumulation chain.  The MMA warp therefore switches accumulator buffer
//     every `flush` k-blocks and the epilogue warps add the partial tiles in fp32 registers with
//     round-to-nearest: the truncated chain is at most flush * 32 * 3 products long instead of 3 K.
#include "swr_common.cuh"
#include "swr_launch.h"
#include "swr_tc.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace swr {
using namespace tc;

constexpr int T2_BM = 128;                 // accumulator rows (TMEM lanes) per tile
constexpr int T2_NSTAGER = 8;              // warps 0..7
constexpr int T2_NEPI = 8;                 // warps 8..15
constexpr int T2_W_TMA = 16, T2_W_MMA = 17;
constexpr int T2_THREADS = 18 * 32;
constexpr int T2_MAX_STAGES = 4;
constexpr uint32_t T2_ACC_COLS = 128;      // columns per accumulator buffer; buffers at TMEM columns 0 and 128
constexpr uint32_t T2_A_COL0 = 256;        // TMEM-resident operand: stage s at columns 256 + 64 s (32 hi + 32 lo)
constexpr uint32_t T2_TMEM_COLS = 512;
constexpr int T2_OT_LD = 68;               // epilogue transpose tile: 128 rows x 64 columns (+4 pad)
constexpr int T2_OT_BYTES = T2_BM * T2_OT_LD * 4;
constexpr int T2_RED_BYTES = 2 * T2_NEPI * 64 * 8;
constexpr int T2_BAR_STAGE = 1, T2_BAR_EPI = 2;   // named barriers of the stager / epilogue warps (256 threads each)
enum { T2_FWD = 0, T2_DGRAD = 1 };

struct alignas(64) Tc2Params {
  CUtensorMap tm[kMaxGroups];      // weight image of group g for this pass (forward: img_f, data gradient: img_d)
  FcGroup g[kMaxGroups];
  int tile_start[kMaxGroups + 1];  // fwd / wgrad: first tile of group g;  dgrad: first k-block of group g
  int nt[kMaxGroups];              // accumulator columns per tile: per group (fwd, wgrad) / per destination (dgrad)
  int n_groups, B;
  float inv_count;
  int stages, flush, n_tiles;
  int stage_bytes;                 // stride of the weight-stage ring (sized for the widest tile of the launch)
  int off_coef, off_ccs, off_ot, off_red;   // shared-memory byte offsets (from the 1024-aligned base)
  int n_dst;
  int dst_group[kMaxGroups + 1];
  int dst_tile[kMaxGroups + 1];
  unsigned dst_atomic;             // bit d: destination entry d is one of several partial fan-ins: add atomically
  int splits, rows_per_split;      // wgrad
};

struct Tc2Shared {
  uint64_t full_b[T2_MAX_STAGES];  // TMA landed the weight tile of the stage
  uint64_t full_a[T2_MAX_STAGES];  // the stager warps wrote their operand(s) of the stage
  uint64_t empty[T2_MAX_STAGES];   // tcgen05.commit: the MMAs that read the stage are done
  uint64_t acc_full[2];            // tcgen05.commit: a partial accumulator is complete
  uint64_t acc_empty[2];           // the epilogue warps have drained it
  uint32_t tmem_base;
};

__device__ __forceinline__ uint8_t* t2_align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

struct T2Ring {                    // position in a ring of mbarrier-guarded stages
  int s; uint32_t ph;
  __device__ __forceinline__ void init() { s = 0; ph = 0; }
  __device__ __forceinline__ void next(int S) { if (++s == S) { s = 0; ph ^= 1u; } }
};

struct T2Tile {
  int g;            // fwd / wgrad: group;  dgrad: first group of the destination
  int ge;           // dgrad: one past the last group of the destination
  int d;            // dgrad: destination entry
  int m0, n0;       // first accumulator row (batch row; wgrad: output feature) / column of the tile
  int NT;           // accumulator columns
  int nkb;          // k-blocks of 32 contraction elements
  int kb0;          // dgrad: first k-block of the destination in the launch-wide numbering
  int b_begin, b_end;   // wgrad: batch rows of this split
};

template <int MODE>
__device__ __forceinline__ T2Tile t2_decode(const Tc2Params& p, int t) {
  T2Tile T{};
  if (MODE == T2_FWD) {
    int g = 0;
    while (g + 1 < p.n_groups && p.tile_start[g + 1] <= t) ++g;
    const int N = p.g[g].Y.n, K = p.g[g].A.n;
    T.g = g; T.NT = p.nt[g];
    const int nt_n = (N + T.NT - 1) / T.NT, local = t - p.tile_start[g];
    T.m0 = (local / nt_n) * T2_BM; T.n0 = (local % nt_n) * T.NT;
    T.nkb = (K + KBLK - 1) / KBLK;
  } else {
    int d = 0;
    while (d + 1 < p.n_dst && p.dst_tile[d + 1] <= t) ++d;
    T.d = d; T.g = p.dst_group[d]; T.ge = p.dst_group[d + 1];
    const int Kd = p.g[T.g].A.n;
    T.NT = p.nt[d];
    const int nt_n = (Kd + T.NT - 1) / T.NT, local = t - p.dst_tile[d];
    T.m0 = (local / nt_n) * T2_BM; T.n0 = (local % nt_n) * T.NT;
    T.kb0 = p.tile_start[T.g];
    T.nkb = p.tile_start[T.ge] - T.kb0;
  }
  return T;
}

// 12 MMAs of one k-block: A (hi / lo, 32 + 32 TMEM columns) x B (hi / lo tiles in shared memory)
__device__ __forceinline__ void t2_issue(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_saddr, uint32_t b_bytes, bool b_mn, uint32_t idesc, bool first) {
  const uint64_t dbh0 = b_mn ? mnmajor_desc(b_saddr, 0) : kmajor_desc(b_saddr, 0);
  const uint64_t b_lo_d = (uint64_t)(b_bytes >> 4);
  const uint64_t b_step = b_mn ? (1024u >> 4) : ((UMMA_K * 4) >> 4);
#pragma unroll
  for (int ks = 0; ks < KBLK / UMMA_K; ++ks) {
    const uint64_t dbh = dbh0 + ks * b_step;
    mma_tf32_ts(d_tmem, a_tmem + 32 + 8 * ks, dbh, idesc, (first && ks == 0) ? 0u : 1u);   // lo * hi
    mma_tf32_ts(d_tmem, a_tmem + 8 * ks, dbh + b_lo_d, idesc, 1u);                          // hi * lo
    mma_tf32_ts(d_tmem, a_tmem + 8 * ks, dbh, idesc, 1u);                                   // hi * hi
  }
}

__device__ __forceinline__ void t2_setup(Tc2Shared& sh, int S, int tid, int warp) {
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&sh.full_b[s], 1); mbar_init(&sh.full_a[s], T2_NSTAGER); mbar_init(&sh.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.acc_full[b], 1); mbar_init(&sh.acc_empty[b], T2_NEPI); }
    fence_mbar_init();
  }
  if (warp == T2_W_MMA) tmem_alloc(&sh.tmem_base, T2_TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
}

__device__ __forceinline__ void t2_range(int n_tiles, int& t_begin, int& t_end) {
  const int per = n_tiles / (int)gridDim.x, rem = n_tiles % (int)gridDim.x, b = (int)blockIdx.x;
  t_begin = b * per + min(b, rem);
  t_end = t_begin + per + (b < rem ? 1 : 0);
}

__device__ __forceinline__ float4 t2_ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 t2_zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void t2_red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 4 floats at p (16-byte aligned, readable), components >= nvalid zeroed
__device__ __forceinline__ float4 t2_ld4_mask(const float* p, int nvalid) {
  if (nvalid <= 0) return t2_zero4();
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  if (nvalid < 4) { v.w = 0.f; if (nvalid < 3) v.z = 0.f; if (nvalid < 2) v.y = 0.f; }
  return v;
}
__device__ __forceinline__ float t2_slope(int act) { return act == SWR_ACT_RELU ? 0.f : (act == SWR_ACT_LEAKY ? 0.1f : 1.f); }
__device__ __forceinline__ float t2_act(float z, float slope, bool sig) {
  return sig ? 1.f / (1.f + expf(-z)) : fmaxf(z, slope * z);
}

// split 16 values and store them as this thread's 16 hi + 16 lo columns of an operand stage
__device__ __forceinline__ void t2_store_split16(uint32_t taddr_hi, const float (&v)[16]) {
  float hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) split_tf32(v[i], hi[i], lo[i]);
  tmem_st16(taddr_hi, hi);
  tmem_st16(taddr_hi + 32, lo);
}

// ---- epilogue helpers -------------------------------------------------------------------------------------
// drain this warp's share (lane quarter q, columns [64 ch, 64 ch + 64) of an NT-wide accumulator) and add it to acc
__device__ __forceinline__ void t2_drain_add(uint32_t tmem_acc, int q, int ch, int NT, float (&acc)[64]) {
  const int my = min(max(NT - 64 * ch, 0), 64);
  const uint32_t taddr = tmem_acc + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * ch);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (16 * c < my) {      // warp-uniform
      uint32_t r[16];
      tmem_ld16(taddr + 16 * c, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[16 * c + i] += __uint_as_float(r[i]);
    }
  }
}
// the partial sums of the warps that own column half `pass` -> ot[128][T2_OT_LD]
__device__ __forceinline__ void t2_acc_to_smem(float* ot, int row, int pc, const float (&acc)[64]) {
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (4 * i < pc) *reinterpret_cast<float4*>(ot + (size_t)row * T2_OT_LD + 4 * i) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
}
// red: [2][T2_NEPI][64] doubles -> one fp64 atomic per column and statistic
__device__ __forceinline__ void t2_col_atomics(const double* red, double* gstats, int col0, int nvalid, int etid) {
  for (int i = etid; i < 2 * nvalid; i += 32 * T2_NEPI) {
    const int which = i / nvalid, col = i - which * nvalid;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < T2_NEPI; ++w) t += red[(which * T2_NEPI + w) * 64 + col];
    atomicAdd(gstats + 2 * (col0 + col) + which, t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward:  Y[m, n] = sum_k act(norm(A))[m, k] * Weff[n, k] + beff[n]                               (MODE = T2_FWD)
// data gradient: dA[m, j] = sum_g sum_n dY_g[m, n] * Weff_g[n, j], then the destination's act' / norm stage 1
//                                                                                                  (MODE = T2_DGRAD)
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(T2_THREADS, 1) fc_tc2_kernel(const __grid_constant__ Tc2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Tc2Shared sh;
  uint8_t* smem = t2_align1024(smem_raw);
  const uint32_t smem_s = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages, F = p.flush, M = p.B;
  int t_begin, t_end;
  t2_range(p.n_tiles, t_begin, t_end);
  t2_setup(sh, S, tid, warp);
  const uint32_t tmem = sh.tmem_base;

  if (warp == T2_W_TMA) {
    // ===== TMA producer: weight tiles =====
    if (lane == 0) {
      T2Ring r; r.init();
      for (int t = t_begin; t < t_end; ++t) {
        const T2Tile T = t2_decode<MODE>(p, t);
        const uint32_t bytes = 2u * (uint32_t)T.NT * 128u, stage_bytes = (uint32_t)p.stage_bytes;
        if (MODE == T2_FWD) {
          for (int kb = 0; kb < T.nkb; ++kb, r.next(S)) {
            mbar_wait(&sh.empty[r.s], r.ph ^ 1u);
            mbar_expect_tx(&sh.full_b[r.s], bytes);
            tma_load_3d(smem_s + (uint32_t)r.s * stage_bytes, &p.tm[T.g], kb * KBLK, T.n0, 0, &sh.full_b[r.s]);
          }
        } else {
          for (int g = T.g; g < T.ge; ++g) {
            const int nk = p.tile_start[g + 1] - p.tile_start[g];
            for (int kb = 0; kb < nk; ++kb, r.next(S)) {
              mbar_wait(&sh.empty[r.s], r.ph ^ 1u);
              mbar_expect_tx(&sh.full_b[r.s], bytes);
              tma_load_3d(smem_s + (uint32_t)r.s * stage_bytes, &p.tm[g], kb * KBLK, T.n0, 0, &sh.full_b[r.s]);
            }
          }
        }
      }
    }
  } else if (warp == T2_W_MMA) {
    // ===== MMA issuer =====
    if (lane == 0) {
      T2Ring r; r.init();
      uint32_t acc_it = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const T2Tile T = t2_decode<MODE>(p, t);
        const uint32_t b_bytes = (uint32_t)T.NT * 128u, stage_bytes = (uint32_t)p.stage_bytes;
        const uint32_t idesc = make_idesc_tf32(T2_BM, T.NT, false, false);
        int fpos = 0;
        for (int kb = 0; kb < T.nkb; ++kb, r.next(S)) {
          const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
          if (fpos == 0) { mbar_wait(&sh.acc_empty[buf], aph ^ 1u); fence_after_sync(); }
          mbar_wait(&sh.full_b[r.s], r.ph);
          mbar_wait(&sh.full_a[r.s], r.ph);
          fence_after_sync();
          t2_issue(tmem + buf * T2_ACC_COLS, tmem + T2_A_COL0 + (uint32_t)r.s * 64u, smem_s + (uint32_t)r.s * stage_bytes, b_bytes,
                   false, idesc, fpos == 0);
          mma_commit(&sh.empty[r.s]);
          if (++fpos == F || kb == T.nkb - 1) { mma_commit(&sh.acc_full[buf]); ++acc_it; fpos = 0; }
        }
      }
    }
  } else if (warp < T2_NSTAGER) {
    // ===== stagers: the 128-row operand, registers -> TMEM =====
    const int q = warp & 3, kh = warp >> 2, stid = tid;          // lane quarter, half of the k-block, 0..255
    const int row = 32 * q + lane;
    const uint32_t ta = tmem + T2_A_COL0 + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * kh);
    float* coef = reinterpret_cast<float*>(smem + p.off_coef);
    T2Ring r; r.init();
    int cur_key = -1;
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = t2_decode<MODE>(p, t);
      const int mrow = min(T.m0 + row, M - 1);      // rows past the batch only feed accumulator rows that are never stored
      if (MODE == T2_FWD) {
        const FcGroup& G = p.g[T.g];
        const int K = G.A.n, Kpad = T.nkb * KBLK;
        const bool plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
        if (!plainA && cur_key != T.g) {            // [3][Kpad]: mu, s, b of the input columns (zero beyond K)
          named_bar(T2_BAR_STAGE, 32 * T2_NSTAGER);
          for (int k = stid; k < Kpad; k += 32 * T2_NSTAGER) {
            ColCoef c = {0.f, 0.f, 0.f, 0.f};
            if (k < K) c = col_coef(G.A.norm, k, p.inv_count);
            coef[k] = c.mu; coef[Kpad + k] = c.s; coef[2 * Kpad + k] = c.b;
          }
          named_bar(T2_BAR_STAGE, 32 * T2_NSTAGER);
          cur_key = T.g;
        }
        const float slope = t2_slope(G.A.act);
        const bool sig = G.A.act == SWR_ACT_SIGMOID;
        const float* src = G.A.raw + (int64_t)mrow * G.A.ld + 16 * kh;
        float4 x[4];
        auto load = [&](int kb) {
          const int k = kb * KBLK + 16 * kh;
#pragma unroll
          for (int i = 0; i < 4; ++i) x[i] = t2_ld4_mask(src + kb * KBLK + 4 * i, K - (k + 4 * i));
        };
        load(0);
        for (int kb = 0; kb < T.nkb; ++kb, r.next(S)) {
          mbar_wait(&sh.empty[r.s], r.ph ^ 1u);
          fence_after_sync();
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) { v[4 * i] = x[i].x; v[4 * i + 1] = x[i].y; v[4 * i + 2] = x[i].z; v[4 * i + 3] = x[i].w; }
          if (!plainA) {
            const float* c = coef + kb * KBLK + 16 * kh;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 mu = t2_ld4s(c + 4 * i), sc = t2_ld4s(c + Kpad + 4 * i), bb = t2_ld4s(c + 2 * Kpad + 4 * i);
              v[4 * i] = t2_act(fmaf(v[4 * i] - mu.x, sc.x, bb.x), slope, sig);
              v[4 * i + 1] = t2_act(fmaf(v[4 * i + 1] - mu.y, sc.y, bb.y), slope, sig);
              v[4 * i + 2] = t2_act(fmaf(v[4 * i + 2] - mu.z, sc.z, bb.z), slope, sig);
              v[4 * i + 3] = t2_act(fmaf(v[4 * i + 3] - mu.w, sc.w, bb.w), slope, sig);
            }
          }
          t2_store_split16(ta + (uint32_t)r.s * 64u, v);
          if (kb + 1 < T.nkb) load(kb + 1);
          tmem_st_wait();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh.full_a[r.s]);
        }
      } else {
        // coefficients of dY = c0 * dz + c1 * raw + c2 over the concatenated output columns of the fan-in
        const int Kc = T.nkb * KBLK;
        if (cur_key != T.d) {
          named_bar(T2_BAR_STAGE, 32 * T2_NSTAGER);
          for (int g = T.g; g < T.ge; ++g) {
            const FcGroup& G = p.g[g];
            const int base = (p.tile_start[g] - T.kb0) * KBLK, span = (p.tile_start[g + 1] - p.tile_start[g]) * KBLK;
            for (int n = stid; n < span; n += 32 * T2_NSTAGER) {
              DyCoef c = {0.f, 0.f, 0.f};
              if (n < G.Y.n) c = dy_coef(G.Y, n, p.inv_count);
              coef[base + n] = c.c0; coef[Kc + base + n] = c.c1; coef[2 * Kc + base + n] = c.c2;
            }
          }
          named_bar(T2_BAR_STAGE, 32 * T2_NSTAGER);
          cur_key = T.d;
        }
        int g = T.g, g_kb0 = 0, g_nk = p.tile_start[T.g + 1] - p.tile_start[T.g];
        float4 x[4], w[4];
        // (group, k-block inside it) of the tile's k-block kb; advances g monotonically
        auto load = [&](int kb) {
          while (kb >= g_kb0 + g_nk) { g_kb0 += g_nk; ++g; g_nk = p.tile_start[g + 1] - p.tile_start[g]; }
          const FcGroup& G = p.g[g];
          const int lk = (kb - g_kb0) * KBLK + 16 * kh, N = G.Y.n;
          const int64_t o = (int64_t)mrow * G.Y.ld + lk;
          const bool need_raw = (G.Y.norm.mode == SWR_NORM_BATCH);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            x[i] = t2_ld4_mask(G.Y.dz + o + 4 * i, N - (lk + 4 * i));
            w[i] = need_raw ? t2_ld4_mask(G.Y.raw + o + 4 * i, N - (lk + 4 * i)) : t2_zero4();
          }
        };
        load(0);
        for (int kb = 0; kb < T.nkb; ++kb, r.next(S)) {
          mbar_wait(&sh.empty[r.s], r.ph ^ 1u);
          fence_after_sync();
          float v[16];
          const float* c = coef + kb * KBLK + 16 * kh;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 c0 = t2_ld4s(c + 4 * i), c1 = t2_ld4s(c + Kc + 4 * i), c2 = t2_ld4s(c + 2 * Kc + 4 * i);
            v[4 * i] = fmaf(c0.x, x[i].x, fmaf(c1.x, w[i].x, c2.x));
            v[4 * i + 1] = fmaf(c0.y, x[i].y, fmaf(c1.y, w[i].y, c2.y));
            v[4 * i + 2] = fmaf(c0.z, x[i].z, fmaf(c1.z, w[i].z, c2.z));
            v[4 * i + 3] = fmaf(c0.w, x[i].w, fmaf(c1.w, w[i].w, c2.w));
          }
          t2_store_split16(ta + (uint32_t)r.s * 64u, v);
          if (kb + 1 < T.nkb) load(kb + 1);
          tmem_st_wait();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh.full_a[r.s]);
        }
      }
    }
  } else {
    // ===== epilogue warps =====
    const int e = warp - T2_NSTAGER, q = warp & 3, ch = e >> 2, etid = tid - 32 * T2_NSTAGER;
    const int arow = 32 * q + lane;                  // accumulator row this thread drains
    const int rsub = lane >> 4, c4 = (lane & 15) * 4;   // coalesced pass: 16 lanes per row, 2 rows per warp step
    float* ot = reinterpret_cast<float*>(smem + p.off_ot);
    double* red = reinterpret_cast<double*>(smem + p.off_red);
    float* ccs = reinterpret_cast<float*>(smem + p.off_ccs);
    uint32_t acc_it = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = t2_decode<MODE>(p, t);
      const FcGroup& G = p.g[T.g];
      const ActDev& D = G.A;                          // dgrad: the destination
      const int Nfull = (MODE == T2_FWD) ? G.Y.n : D.n;
      const int Nend = min(Nfull, T.n0 + T.NT);
      bool plainD = true, has_norm = false;
      if (MODE == T2_DGRAD) {
        has_norm = D.norm.mode != SWR_NORM_NONE;
        plainD = !has_norm && D.act == SWR_ACT_NONE;
        // coefficients of the destination's own norm / activation: [4][NT] mu, s, b, r
        named_bar(T2_BAR_EPI, 32 * T2_NEPI);
        for (int c = etid; c < T.NT; c += 32 * T2_NEPI) {
          ColCoef cc = {0.f, 1.f, 0.f, 1.f};
          if (!plainD && T.n0 + c < Nend) cc = col_coef(D.norm, T.n0 + c, p.inv_count);
          ccs[c] = cc.mu; ccs[T.NT + c] = cc.s; ccs[2 * T.NT + c] = cc.b; ccs[3 * T.NT + c] = cc.r;
        }
        named_bar(T2_BAR_EPI, 32 * T2_NEPI);
      }
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      const int nflush = (T.nkb + F - 1) / F;
      for (int f = 0; f < nflush; ++f, ++acc_it) {
        const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        mbar_wait(&sh.acc_full[buf], aph);
        fence_after_sync();
        t2_drain_add(tmem + buf * T2_ACC_COLS, q, ch, T.NT, acc);
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.acc_empty[buf]);
      }
      // ---- final epilogue, 64 accumulator columns per pass ----
      for (int pass = 0; pass < 2; ++pass) {
        const int pc0 = 64 * pass;
        if (pc0 >= T.NT) break;
        const int pc = min(64, T.NT - pc0);
        const int nvalid = min(max(Nend - (T.n0 + pc0), 0), pc);
        named_bar(T2_BAR_EPI, 32 * T2_NEPI);          // everybody is done with the previous contents of ot / red
        if (ch == pass) t2_acc_to_smem(ot, arow, pc, acc);
        named_bar(T2_BAR_EPI, 32 * T2_NEPI);
        const int nv = nvalid - c4;                  // valid components of this lane's column quad (<= 0: none)
        const int n = T.n0 + pc0 + c4;               // first output column of the quad
        double s1d[4] = {0.0, 0.0, 0.0, 0.0}, s2d[4] = {0.0, 0.0, 0.0, 0.0};
        if (MODE == T2_FWD) {
          float* Y = const_cast<float*>(G.Y.raw);
          const bool vec = (nv >= 4) && (G.Y.ld % 4 == 0) && is_al16(Y) && (n % 4 == 0);
          if (nv > 0) {
            float bias[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bias[j] = (j < nv) ? ld_opt(G.bias, n + j, 0.f) + ld_opt(G.bias2, n + j, 0.f) : 0.f;
            // moments of this lane's 8 rows: fp32 sums centred on the first value, widened to fp64 once
            float t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f}, y0[4] = {0.f, 0.f, 0.f, 0.f};
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 16 * e + 2 * i + rsub, m = T.m0 + rr;
              if (m < M) {
                const float4 a4 = t2_ld4s(ot + (size_t)rr * T2_OT_LD + c4);
                float y[4] = {a4.x + bias[0], a4.y + bias[1], a4.z + bias[2], a4.w + bias[3]};
                if (G.e_act != SWR_ACT_NONE) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) y[j] = act_fwd(y[j], G.e_act) * G.e_scale;
                }
                float* dst = Y + (int64_t)m * G.Y.ld + n;
                if (vec) *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
                else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) if (j < nv) dst[j] = y[j];
                }
                if (cnt == 0) { y0[0] = y[0]; y0[1] = y[1]; y0[2] = y[2]; y0[3] = y[3]; }
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float dlt = y[j] - y0[j]; t1[j] += dlt; t2[j] = fmaf(dlt, dlt, t2[j]); }
                ++cnt;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double dy0 = (double)y0[j], dt1 = (double)t1[j];
              s1d[j] = (double)cnt * dy0 + dt1;
              s2d[j] = (double)cnt * dy0 * dy0 + 2.0 * dy0 * dt1 + (double)t2[j];
            }
          }
        } else {
          const bool accumulate = (G.flags & FC_A_ACCUMULATE) != 0;
          const bool atomic_dst = (p.dst_atomic >> T.d) & 1u;
          const bool vec = (nv >= 4) && (D.ld % 4 == 0) && is_al16(D.dz) && is_al16(D.raw) && (n % 4 == 0);
          if (nv > 0) {
            const int cc0 = pc0 + c4;
            const float4 mu4 = t2_ld4s(ccs + cc0), sc4 = t2_ld4s(ccs + T.NT + cc0), bb4 = t2_ld4s(ccs + 2 * T.NT + cc0), rr4 = t2_ld4s(ccs + 3 * T.NT + cc0);
            const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
            const float bb[4] = {bb4.x, bb4.y, bb4.z, bb4.w}, rr_[4] = {rr4.x, rr4.y, rr4.z, rr4.w};
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 16 * e + 2 * i + rsub, m = T.m0 + rr;
              if (m < M) {
                const float4 a4 = t2_ld4s(ot + (size_t)rr * T2_OT_LD + c4);
                float dz[4] = {a4.x, a4.y, a4.z, a4.w};
                const int64_t o = (int64_t)m * D.ld + n;
                if (!plainD) {
                  float raw[4];
                  if (vec) { const float4 r4 = *reinterpret_cast<const float4*>(D.raw + o); raw[0] = r4.x; raw[1] = r4.y; raw[2] = r4.z; raw[3] = r4.w; }
                  else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) raw[j] = (j < nv) ? D.raw[o + j] : 0.f;
                  }
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    dz[j] *= act_grad(fmaf(raw[j] - mu[j], sc[j], bb[j]), D.act);
                    s1[j] += dz[j];
                    s2[j] = fmaf(dz[j], (raw[j] - mu[j]) * rr_[j], s2[j]);
                  }
                }
                float* dst = D.dz + o;
                if (atomic_dst) {                     // plain destination split over its fan-in
                  if (vec) t2_red_add_v4(dst, make_float4(dz[0], dz[1], dz[2], dz[3]));
                  else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < nv) atomicAdd(dst + j, dz[j]);
                  }
                } else if (vec) {
                  float4 o4 = make_float4(dz[0], dz[1], dz[2], dz[3]);
                  if (accumulate) { const float4 old = *reinterpret_cast<const float4*>(dst); o4.x += old.x; o4.y += old.y; o4.z += old.z; o4.w += old.w; }
                  *reinterpret_cast<float4*>(dst) = o4;
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) if (j < nv) dst[j] = accumulate ? dst[j] + dz[j] : dz[j];
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { s1d[j] = (double)s1[j]; s2d[j] = (double)s2[j]; }
          }
        }
        const bool want_stats = (MODE == T2_FWD) ? (G.stats_out != nullptr) : (has_norm && D.dstats != nullptr);
        if (want_stats) {           // uniform over the epilogue warps
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s1d[j] += __shfl_xor_sync(0xffffffffu, s1d[j], 16);
            s2d[j] += __shfl_xor_sync(0xffffffffu, s2d[j], 16);
          }
          if (rsub == 0 && nv > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { red[(0 * T2_NEPI + e) * 64 + c4 + j] = s1d[j]; red[(1 * T2_NEPI + e) * 64 + c4 + j] = s2d[j]; }
          }
          named_bar(T2_BAR_EPI, 32 * T2_NEPI);
          t2_col_atomics(red, (MODE == T2_FWD) ? G.stats_out : D.dstats, T.n0 + pc0, nvalid, etid);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == T2_W_MMA) { __syncwarp(); tmem_dealloc(tmem, T2_TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------------------
// weight / bias gradient: dWeff[n, j] = sum_b dY[b, n] * act(norm(A))[b, j],  db[n] = sum_b dY[b, n]
// accumulator rows = output features n, columns = input features j, contraction = the batch rows of one split.
// Both operands are computed: dY^T goes registers -> TMEM (lane = output feature, coalesced 128-byte reads over
// 32 features per batch row), the activations go to shared memory in the MN-major swizzled layout.  No TMA here.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ T2Tile t2_decode_wgrad(const Tc2Params& p, int t) {
  T2Tile T{};
  int g = 0;
  while (g + 1 < p.n_groups && p.tile_start[g + 1] <= t) ++g;
  const int N = p.g[g].Y.n, K = p.g[g].A.n;
  T.g = g; T.NT = p.nt[g];
  const int nt_m = (N + T2_BM - 1) / T2_BM, nt_n = (K + T.NT - 1) / T.NT;
  int local = t - p.tile_start[g];
  const int split = local / (nt_m * nt_n);
  local -= split * nt_m * nt_n;
  T.m0 = (local / nt_n) * T2_BM; T.n0 = (local % nt_n) * T.NT;
  T.b_begin = split * p.rows_per_split;
  T.b_end = min(p.B, T.b_begin + p.rows_per_split);
  T.nkb = (max(T.b_end - T.b_begin, 0) + KBLK - 1) / KBLK;
  return T;
}

__global__ void __launch_bounds__(T2_THREADS, 1) fc_tc2_wgrad_kernel(const __grid_constant__ Tc2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Tc2Shared sh;
  uint8_t* smem = t2_align1024(smem_raw);
  const uint32_t smem_s = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages, F = p.flush;
  int t_begin, t_end;
  t2_range(p.n_tiles, t_begin, t_end);
  t2_setup(sh, S, tid, warp);
  const uint32_t tmem = sh.tmem_base;

  if (warp == T2_W_MMA) {
    if (lane == 0) {
      T2Ring r; r.init();
      uint32_t acc_it = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const T2Tile T = t2_decode_wgrad(p, t);
        const uint32_t b_bytes = (uint32_t)T.NT * 128u, stage_bytes = (uint32_t)p.stage_bytes;
        const uint32_t idesc = make_idesc_tf32(T2_BM, T.NT, false, true);
        int fpos = 0;
        for (int kb = 0; kb < T.nkb; ++kb, r.next(S)) {
          const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
          if (fpos == 0) { mbar_wait(&sh.acc_empty[buf], aph ^ 1u); fence_after_sync(); }
          mbar_wait(&sh.full_a[r.s], r.ph);
          fence_after_sync();
          t2_issue(tmem + buf * T2_ACC_COLS, tmem + T2_A_COL0 + (uint32_t)r.s * 64u, smem_s + (uint32_t)r.s * stage_bytes, b_bytes,
                   true, idesc, fpos == 0);
          mma_commit(&sh.empty[r.s]);
          if (++fpos == F || kb == T.nkb - 1) { mma_commit(&sh.acc_full[buf]); ++acc_it; fpos = 0; }
        }
      }
    }
  } else if (warp < T2_NSTAGER) {
    const int q = warp & 3, kh = warp >> 2, stid = tid;
    const uint32_t ta = tmem + T2_A_COL0 + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * kh);
    float* coef = reinterpret_cast<float*>(smem + p.off_coef);   // [3][NT]: mu, s, b of the input columns of the tile
    T2Ring r; r.init();
    int cur_g = -1, cur_n0 = -1;
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = t2_decode_wgrad(p, t);
      if (T.nkb == 0) continue;
      const FcGroup& G = p.g[T.g];
      const int N = G.Y.n, K = G.A.n, NT = T.NT, nq = NT >> 2, nit = NT >> 5;   // float4 per thread and k-block of the column operand
      const bool plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
      if (cur_g != T.g || cur_n0 != T.n0) {
        named_bar(T2_BAR_STAGE, 32 * T2_NSTAGER);
        for (int i = stid; i < NT; i += 32 * T2_NSTAGER) {
          ColCoef c = {0.f, 0.f, 0.f, 0.f};
          if (T.n0 + i < K) c = plainA ? ColCoef{0.f, 1.f, 0.f, 1.f} : col_coef(G.A.norm, T.n0 + i, p.inv_count);
          coef[i] = c.mu; coef[NT + i] = c.s; coef[2 * NT + i] = c.b;
        }
        named_bar(T2_BAR_STAGE, 32 * T2_NSTAGER);
        cur_g = T.g; cur_n0 = T.n0;
      }
      const int n = T.m0 + 32 * q + lane;             // the output feature this thread stages
      const bool n_ok = n < N;
      DyCoef dc = {0.f, 0.f, 0.f};
      if (n_ok) dc = dy_coef(G.Y, n, p.inv_count);
      const bool need_raw = (G.Y.norm.mode == SWR_NORM_BATCH);
      const float slope = t2_slope(G.A.act);
      const bool sig = G.A.act == SWR_ACT_SIGMOID;
      const int rows = T.b_end - T.b_begin;
      const float* dzp = G.Y.dz + (int64_t)T.b_begin * G.Y.ld + (n_ok ? n : 0);
      const float* rwp = G.Y.raw + (int64_t)T.b_begin * G.Y.ld + (n_ok ? n : 0);
      const float* ap = G.A.raw + (int64_t)T.b_begin * G.A.ld + T.n0;
      const bool vecA = is_al16(G.A.raw) && (G.A.ld % 4 == 0) && (T.n0 % 4 == 0);
      float rowsum = 0.f;
      float xa[16], xr[16];
      float4 xb[4];
      auto load = [&](int kb) {
        const int b0 = kb * KBLK + 16 * kh;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const bool ok = n_ok && (b0 + i < rows);
          xa[i] = ok ? __ldg(dzp + (int64_t)(b0 + i) * G.Y.ld) : 0.f;
          xr[i] = (ok && need_raw) ? __ldg(rwp + (int64_t)(b0 + i) * G.Y.ld) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          if (it < nit) {
            const int v = it * (32 * T2_NSTAGER) + stid, c = v / nq, qd = v - c * nq;
            const int b = min(kb * KBLK + c, rows - 1);       // rows past the split meet an exactly-zero dY column
            const int jv = K - (T.n0 + 4 * qd);
            const float* src = ap + (int64_t)b * G.A.ld + 4 * qd;
            if (vecA) xb[it] = t2_ld4_mask(src, jv);
            else {
              xb[it] = t2_zero4();
              if (jv > 0) xb[it].x = __ldg(src);
              if (jv > 1) xb[it].y = __ldg(src + 1);
              if (jv > 2) xb[it].z = __ldg(src + 2);
              if (jv > 3) xb[it].w = __ldg(src + 3);
            }
          }
        }
      };
      load(0);
      const uint32_t b_bytes = (uint32_t)NT * 128u, stage_bytes = (uint32_t)p.stage_bytes;
      for (int kb = 0; kb < T.nkb; ++kb, r.next(S)) {
        mbar_wait(&sh.empty[r.s], r.ph ^ 1u);
        fence_after_sync();
        {
          float v[16];
          const int b0 = kb * KBLK + 16 * kh;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool ok = n_ok && (b0 + i < rows);          // c2 != 0: contraction padding must stay exactly zero
            v[i] = ok ? fmaf(dc.c0, xa[i], fmaf(dc.c1, xr[i], dc.c2)) : 0.f;
            rowsum += v[i];
          }
          t2_store_split16(ta + (uint32_t)r.s * 64u, v);
        }
        const uint32_t bh = smem_s + (uint32_t)r.s * stage_bytes, bl = bh + b_bytes;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          if (it < nit) {
            const int v = it * (32 * T2_NSTAGER) + stid, c = v / nq, qd = v - c * nq;
            float4 x = xb[it];
            if (!plainA) {
              const float* cf = coef + 4 * qd;
              const float4 mu = t2_ld4s(cf), sc = t2_ld4s(cf + NT), bb = t2_ld4s(cf + 2 * NT);
              x.x = t2_act(fmaf(x.x - mu.x, sc.x, bb.x), slope, sig); x.y = t2_act(fmaf(x.y - mu.y, sc.y, bb.y), slope, sig);
              x.z = t2_act(fmaf(x.z - mu.z, sc.z, bb.z), slope, sig); x.w = t2_act(fmaf(x.w - mu.w, sc.w, bb.w), slope, sig);
            }
            store_split(bh, bl, mnmajor_off(qd, c), x);
          }
        }
        if (kb + 1 < T.nkb) load(kb + 1);
        fence_proxy_async();      // shared-memory stores -> visible to the MMA unit
        tmem_st_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.full_a[r.s]);
      }
      if (T.n0 == 0 && n_ok) {    // bias gradient: the j-tile 0 CTAs carry it (two threads per output feature)
        if (G.dbias) atomicAdd(G.dbias + n, rowsum);
        if (G.dbias2) atomicAdd(G.dbias2 + n, rowsum);
      }
    }
  } else if (warp < T2_NSTAGER + T2_NEPI) {
    const int e = warp - T2_NSTAGER, q = warp & 3, ch = e >> 2;
    const int arow = 32 * q + lane;
    const int rsub = lane >> 4, c4 = (lane & 15) * 4;
    float* ot = reinterpret_cast<float*>(smem + p.off_ot);
    uint32_t acc_it = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = t2_decode_wgrad(p, t);
      if (T.nkb == 0) continue;
      const FcGroup& G = p.g[T.g];
      const int N = G.Y.n, K = G.A.n;
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      const int nflush = (T.nkb + F - 1) / F;
      for (int f = 0; f < nflush; ++f, ++acc_it) {
        const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        mbar_wait(&sh.acc_full[buf], aph);
        fence_after_sync();
        t2_drain_add(tmem + buf * T2_ACC_COLS, q, ch, T.NT, acc);
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.acc_empty[buf]);
      }
      const bool kn = (G.w_layout == SWR_W_KN);
      const int mvalid = min(N - T.m0, T2_BM);
      for (int pass = 0; pass < 2; ++pass) {
        const int pc0 = 64 * pass;
        if (pc0 >= T.NT) break;
        const int pc = min(64, T.NT - pc0);
        const int nvalid = min(max(K - (T.n0 + pc0), 0), pc);
        named_bar(T2_BAR_EPI, 32 * T2_NEPI);          // everybody is done reading the previous contents of ot
        if (ch == pass) t2_acc_to_smem(ot, arow, pc, acc);
        named_bar(T2_BAR_EPI, 32 * T2_NEPI);
        const int j = T.n0 + pc0 + c4, nv = nvalid - c4;
        if (!kn) {                                    // dW[n, j]: 16 lanes cover 64 consecutive input features of one output feature
          const bool vec = !G.W2 && G.dW && (nv >= 4) && (G.ldw % 4 == 0) && is_al16(G.dW) && (j % 4 == 0);
          if (nv > 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 16 * e + 2 * i + rsub;
              if (rr < mvalid) {
                const float4 a4 = t2_ld4s(ot + (size_t)rr * T2_OT_LD + c4);
                const int64_t o = (int64_t)(T.m0 + rr) * G.ldw + j;
                if (vec) t2_red_add_v4(G.dW + o, a4);
                else {
                  const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) {
                    if (jj < nv) {
                      if (G.W2) {
                        if (G.dW) atomicAdd(G.dW + o + jj, a[jj] * __ldg(G.W2 + o + jj));
                        if (G.dW2) atomicAdd(G.dW2 + o + jj, a[jj] * __ldg(G.W + o + jj));
                      } else if (G.dW) atomicAdd(G.dW + o + jj, a[jj]);
                    }
                  }
                }
              }
            }
          }
        } else {                                      // dW[j, n]: lanes run over output features n
          for (int col = e; col < nvalid; col += T2_NEPI) {
            for (int rb = 0; rb < T2_BM; rb += 32) {
              const int rr = rb + lane;
              if (rr < mvalid) {
                const float a = ot[(size_t)rr * T2_OT_LD + col];
                const int64_t o = (int64_t)(T.n0 + pc0 + col) * G.ldw + T.m0 + rr;
                if (G.W2) {
                  if (G.dW) atomicAdd(G.dW + o, a * __ldg(G.W2 + o));
                  if (G.dW2) atomicAdd(G.dW2 + o, a * __ldg(G.W + o));
                } else if (G.dW) atomicAdd(G.dW + o, a);
              }
            }
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == T2_W_MMA) { __syncwarp(); tmem_dealloc(tmem, T2_TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------------------
// presplit: effective weight W (.) W2 -> hi / lo TF32 planes in both orientations
//   img_f [2][N][Kp]  (Kp = K rounded up to 32): row n, contraction k contiguous      -- forward
//   img_d [2][K][Np]  (Np = N rounded up to 32): row k, contraction n contiguous      -- data gradient
// One 32 x 32 tile per 256-thread block, transposed through shared memory; the padding columns are written as zeros.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPresplitMax = 64;
struct PresplitItem { const float* W; const float* W2; float* img_f; float* img_d; int N, K, ldw, layout, block0; };
struct PresplitParams { PresplitItem it[kPresplitMax]; int n; };

__global__ void __launch_bounds__(256) fc_presplit_kernel(const __grid_constant__ PresplitParams p) {
  __shared__ float tile[32][33];
  int gi = 0;
  while (gi + 1 < p.n && p.it[gi + 1].block0 <= (int)blockIdx.x) ++gi;
  const PresplitItem& I = p.it[gi];
  const int N = I.N, K = I.K, Kp = (K + 31) & ~31, Np = (N + 31) & ~31;
  const int tiles_k = Kp >> 5, local = blockIdx.x - I.block0;
  const int n0 = (local / tiles_k) * 32, k0 = (local % tiles_k) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  // load the effective weight so that reads follow the contiguous dimension of the source
  if (I.layout == SWR_W_NK) {       // W[n][k]: tile[n][k]
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int n = n0 + r, k = k0 + tx;
      float w = 0.f;
      if (n < N && k < K) { const int64_t o = (int64_t)n * I.ldw + k; w = __ldg(I.W + o); if (I.W2) w *= __ldg(I.W2 + o); }
      tile[r][tx] = w;
    }
  } else {                          // W[k][n]: read rows k, store transposed
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int k = k0 + r, n = n0 + tx;
      float w = 0.f;
      if (n < N && k < K) { const int64_t o = (int64_t)k * I.ldw + n; w = __ldg(I.W + o); if (I.W2) w *= __ldg(I.W2 + o); }
      tile[tx][r] = w;
    }
  }
  __syncthreads();
  if (I.img_f) {
    float* hi = I.img_f; float* lo = I.img_f + (int64_t)N * Kp;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int n = n0 + r;
      if (n < N) { float h, l; split_tf32(tile[r][tx], h, l); hi[(int64_t)n * Kp + k0 + tx] = h; lo[(int64_t)n * Kp + k0 + tx] = l; }
    }
  }
  if (I.img_d) {
    float* hi = I.img_d; float* lo = I.img_d + (int64_t)K * Np;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int k = k0 + r;
      if (k < K) { float h, l; split_tf32(tile[tx][r], h, l); hi[(int64_t)k * Np + n0 + tx] = h; lo[(int64_t)k * Np + n0 + tx] = l; }
    }
  }
}

int launch_fc_presplit(const FcGroup* groups, int n_groups, cudaStream_t st) {
  int o = 0;
  while (o < n_groups) {
    PresplitParams p{};
    int blocks = 0;
    while (o < n_groups && p.n < kPresplitMax) {
      const FcGroup& G = groups[o++];
      if (!G.img_f && !G.img_d) continue;
      PresplitItem& I = p.it[p.n++];
      I.W = G.W; I.W2 = G.W2; I.img_f = const_cast<float*>(G.img_f); I.img_d = const_cast<float*>(G.img_d);
      I.N = G.Y.n; I.K = G.A.n; I.ldw = G.ldw; I.layout = G.w_layout; I.block0 = blocks;
      blocks += ceil_div(G.Y.n, 32) * ceil_div(G.A.n, 32);
    }
    if (p.n == 0) continue;
    fc_presplit_kernel<<<blocks, 256, 0, st>>>(p);
    SWR_LAUNCH_OK("fc_presplit_kernel");
  }
  return SWR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
static inline int t2_round_up(int a, int b) { return (a + b - 1) / b * b; }

static int t2_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*t2_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static t2_encode_fn t2_encoder() {
  static t2_encode_fn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<t2_encode_fn>(f);
  }();
  return fn;
}
// image [2][rows][pitch] fp32 -> 3-D map {pitch, rows, 2}, box {32, box_rows, 2}, 128-byte swizzle; rows outside the
// image read as zeros
static int t2_make_map(CUtensorMap* tm, const float* img, int rows, int pitch, int box_rows) {
  t2_encode_fn enc = t2_encoder();
  if (!enc) { set_error("fc_tc2: cuTensorMapEncodeTiled is not available"); return SWR_ERR_UNSUPPORTED; }
  const cuuint64_t gdim[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, 2};
  const cuuint64_t gstr[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)rows * pitch * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 2};
  const cuuint32_t est[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("fc_tc2: cuTensorMapEncodeTiled failed (%d) rows=%d pitch=%d box=%d", (int)r, rows, pitch, box_rows); return SWR_ERR_CUDA; }
  return SWR_OK;
}

static int t2_pick_nt(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }
static int t2_flush_for(int nt) {
  static int forced = [] { const char* e = getenv("SWR_TC_FLUSH"); return e ? atoi(e) : 0; }();
  if (forced > 0) return forced;
  return nt >= 128 ? 1 : 128 / nt;       // >= ~768 MMA cycles between accumulator switches; chain <= 8 k-blocks
}

template <class K>
static int t2_set_smem(K kernel, size_t bytes) {
  constexpr int kMaxDyn = 227 * 1024 - 1024;
  if (bytes > (size_t)kMaxDyn) { set_error("fc_tc2: %zu bytes of shared memory needed", bytes); return SWR_ERR_UNSUPPORTED; }
  static thread_local const void* done[8] = {nullptr};
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < 8; ++i) {
    if (done[i] == key) return SWR_OK;
    if (!done[i]) {
      SWR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn));
      done[i] = key;
      return SWR_OK;
    }
  }
  SWR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn));
  return SWR_OK;
}

// shared-memory plan: [stages][2][NT][128 B] | coef | ccs | ot | red
static int t2_plan_smem(Tc2Params& p, int nt_max, size_t coef_bytes, size_t ccs_bytes, int nkb_max, size_t* smem_bytes) {
  const size_t stage = 2 * (size_t)nt_max * 128;
  const size_t fixed = ((coef_bytes + 15) & ~(size_t)15) + ((ccs_bytes + 15) & ~(size_t)15) + T2_OT_BYTES + T2_RED_BYTES;
  const size_t budget = 224 * 1024 - 1024;
  if (fixed + 2 * stage > budget) { set_error("fc_tc2: tables of %zu bytes do not fit beside the stages", fixed); return SWR_ERR_UNSUPPORTED; }
  int s = (int)((budget - fixed) / stage);
  if (s > T2_MAX_STAGES) s = T2_MAX_STAGES;
  if (s > nkb_max && nkb_max >= 2) s = nkb_max;
  if (s < 2) s = 2;
  p.stages = s;
  p.stage_bytes = (int)stage;
  size_t off = (size_t)s * stage;
  p.off_coef = (int)off; off += (coef_bytes + 15) & ~(size_t)15;
  p.off_ccs = (int)off; off += (ccs_bytes + 15) & ~(size_t)15;
  p.off_ot = (int)off; off += T2_OT_BYTES;
  p.off_red = (int)off; off += T2_RED_BYTES;
  *smem_bytes = off + 1024;
  return SWR_OK;
}

static inline bool is_al16_host(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool fc_tc2_usable(const FcGroup* groups, int n_groups, int pass) {
  static const bool off = [] { const char* e = getenv("SWR_FC_TC_V1"); return e && atoi(e) != 0; }();
  if (off) return false;
  for (int g = 0; g < n_groups; ++g) {
    const FcGroup& G = groups[g];
    const bool al = is_al16_host(G.A.raw) && (G.A.ld % 4 == 0) && is_al16_host(G.Y.raw) && (G.Y.ld % 4 == 0);
    if (!al) return false;
    if (pass == 0 && !G.img_f) return false;
    if (pass == 1 && (!G.img_d || !is_al16_host(G.Y.dz) || !is_al16_host(G.A.dz))) return false;
    if (pass == 2 && !is_al16_host(G.Y.dz)) return false;
  }
  return true;
}

int launch_fc_tc2_fwd(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  Tc2Params p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  const int mtiles = ceil_div(B, T2_BM);
  int tiles = 0, kmax = 0, nt_max = 0, flush = 1 << 30;
  for (int g = 0; g < n_groups; ++g) {
    p.g[g] = groups[g];
    const int N = groups[g].Y.n, K = groups[g].A.n;
    p.nt[g] = t2_pick_nt(N);
    p.tile_start[g] = tiles;
    tiles += mtiles * ceil_div(N, p.nt[g]);
    kmax = max(kmax, K); nt_max = max(nt_max, p.nt[g]);
    flush = min(flush, t2_flush_for(p.nt[g]));
    int rc = t2_make_map(&p.tm[g], groups[g].img_f, N, t2_round_up(K, 32), p.nt[g]);
    if (rc) return rc;
  }
  p.tile_start[n_groups] = tiles; p.n_tiles = tiles; p.flush = flush;
  size_t smem = 0;
  int rc = t2_plan_smem(p, nt_max, 3 * sizeof(float) * (size_t)t2_round_up(kmax, KBLK), 0, ceil_div(kmax, KBLK), &smem);
  if (rc) return rc;
  rc = t2_set_smem(fc_tc2_kernel<T2_FWD>, smem);
  if (rc) return rc;
  fc_tc2_kernel<T2_FWD><<<min(tiles, t2_num_sms()), T2_THREADS, smem, st>>>(p);
  SWR_LAUNCH_OK("fc_tc2_kernel<fwd>");
  return SWR_OK;
}

// dst_group[0..n_dst]: group ranges of the destinations (prepared by launch_fc_dgrad)
int launch_fc_tc2_dgrad(const FcGroup* groups, const int* dst_group_in, int n_dst, int n_groups, int64_t B, cudaStream_t st) {
  Tc2Params p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  int kb = 0;
  for (int g = 0; g < n_groups; ++g) { p.g[g] = groups[g]; p.tile_start[g] = kb; kb += ceil_div(groups[g].Y.n, KBLK); }
  p.tile_start[n_groups] = kb;
  const int mtiles = ceil_div(B, T2_BM);
  int dst_group[kMaxGroups + 1];
  for (int d = 0; d <= n_dst; ++d) dst_group[d] = dst_group_in[d];
  // A single destination with a long fan-in (the embedding output at level 0: every expert and gate) is split into
  // partial fan-ins that add atomically, so that the launch has about two tiles per SM instead of a fraction of a wave.
  // Only a plain destination qualifies (no activation / norm in its epilogue, so partial sums commute).
  if (n_dst == 1 && n_groups >= 2 && kb >= 16) {
    const ActDev& D = groups[0].A;
    const bool plain = D.norm.mode == SWR_NORM_NONE && D.act == SWR_ACT_NONE;
    const int base = mtiles * ceil_div(D.n, t2_pick_nt(D.n));
    int chunks = min(min(n_groups, 31), max(1, (2 * t2_num_sms()) / max(1, base)));
    if (plain && chunks >= 2) {
      if (!(groups[0].flags & FC_A_ACCUMULATE))
        SWR_CUDA_OK(cudaMemsetAsync(D.dz, 0, sizeof(float) * (size_t)B * D.ld, st));
      int d = 0, acc = 0;
      dst_group[0] = 0;
      for (int g = 0; g < n_groups; ++g) {
        acc += p.tile_start[g + 1] - p.tile_start[g];
        const int left_groups = n_groups - 1 - g, left_chunks = chunks - 1 - d;
        const int next = g + 1 < n_groups ? p.tile_start[g + 2] - p.tile_start[g + 1] : 0;
        if (left_chunks > 0 && left_groups >= left_chunks && 2 * acc * chunks + next * chunks >= 2 * kb * (d + 1)) dst_group[++d] = g + 1;
      }
      n_dst = d + 1;
      dst_group[n_dst] = n_groups;
      p.dst_atomic = n_dst >= 32 ? 0xffffffffu : ((1u << n_dst) - 1u);
    }
  }
  p.n_dst = n_dst;
  for (int d = 0; d <= n_dst; ++d) p.dst_group[d] = dst_group[d];
  int tiles = 0, nkb_max = 0, nt_max = 0, flush = 1 << 30;
  for (int d = 0; d < n_dst; ++d) {
    const int kd = groups[dst_group[d]].A.n;
    p.nt[d] = t2_pick_nt(kd);
    p.dst_tile[d] = tiles;
    tiles += mtiles * ceil_div(kd, p.nt[d]);
    nkb_max = max(nkb_max, p.tile_start[dst_group[d + 1]] - p.tile_start[dst_group[d]]);
    nt_max = max(nt_max, p.nt[d]);
    flush = min(flush, t2_flush_for(p.nt[d]));
    for (int g = dst_group[d]; g < dst_group[d + 1]; ++g) {
      int rc = t2_make_map(&p.tm[g], groups[g].img_d, groups[g].A.n, t2_round_up(groups[g].Y.n, 32), p.nt[d]);
      if (rc) return rc;
    }
  }
  p.dst_tile[n_dst] = tiles; p.n_tiles = tiles; p.flush = flush;
  size_t smem = 0;
  int rc = t2_plan_smem(p, nt_max, 3 * sizeof(float) * (size_t)nkb_max * KBLK, 4 * sizeof(float) * (size_t)nt_max, nkb_max, &smem);
  if (rc) return rc;
  rc = t2_set_smem(fc_tc2_kernel<T2_DGRAD>, smem);
  if (rc) return rc;
  fc_tc2_kernel<T2_DGRAD><<<min(tiles, t2_num_sms()), T2_THREADS, smem, st>>>(p);
  SWR_LAUNCH_OK("fc_tc2_kernel<dgrad>");
  return SWR_OK;
}

int launch_fc_tc2_wgrad(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  Tc2Params p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  int base = 0, nt_max = 0, flush = 1 << 30;
  for (int g = 0; g < n_groups; ++g) {
    p.g[g] = groups[g];
    p.nt[g] = max(32, t2_pick_nt(groups[g].A.n));      // the MN-major column operand is staged in 32-wide groups
    base += ceil_div(groups[g].Y.n, T2_BM) * ceil_div(groups[g].A.n, p.nt[g]);
    nt_max = max(nt_max, p.nt[g]);
    flush = min(flush, t2_flush_for(p.nt[g]));
  }
  // split the batch so that the launch has about two tiles per SM; keep >= 4 k-blocks (128 rows) per split
  const int sms = t2_num_sms();
  int splits = max(1, min((2 * sms + base / 2) / max(base, 1), ceil_div(B, 4 * KBLK)));
  int rows = t2_round_up(ceil_div(B, splits), KBLK);
  splits = ceil_div(B, rows);
  p.splits = splits; p.rows_per_split = rows;
  int tiles = 0;
  for (int g = 0; g < n_groups; ++g) {
    p.tile_start[g] = tiles;
    tiles += ceil_div(groups[g].Y.n, T2_BM) * ceil_div(groups[g].A.n, p.nt[g]) * splits;
  }
  p.tile_start[n_groups] = tiles; p.n_tiles = tiles; p.flush = flush;
  size_t smem = 0;
  int rc = t2_plan_smem(p, nt_max, 3 * sizeof(float) * (size_t)nt_max, 0, ceil_div(rows, KBLK), &smem);
  if (rc) return rc;
  rc = t2_set_smem(fc_tc2_wgrad_kernel, smem);
  if (rc) return rc;
  fc_tc2_wgrad_kernel<<<min(tiles, sms), T2_THREADS, smem, st>>>(p);
  SWR_LAUNCH_OK("fc_tc2_wgrad_kernel");
  return SWR_OK;
}

}  // namespace swr
