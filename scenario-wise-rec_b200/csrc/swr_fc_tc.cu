// swr_fc_tc.cu -- grouped fully-connected kernels on the sm_100a tensor cores, second generation:
// persistent, warp-specialised, TMA-fed.
//
// Same contract as the FFMA kernels in swr_fc.cu (fc_fwd / fc_dgrad / fc_wgrad over FcGroup lists; reference:
// basic/layers.py:253-258 Linear -> BatchNorm1d -> act, star.py:103-110, ppnet.py:21-29, hamur.py adapters,
// m3oe.py:45-68).  Arithmetic: 3xTF32 (x = hi + lo, three tcgen05.mma.kind::tf32 per k-step).
//
//   * Weights are split ONCE per forward pass by fc_presplit_kernel into "images": the effective weight
//     W (.) W2 as hi / lo TF32 planes, in both orientations (contraction-contiguous for the forward and for
//     the data gradient).  Weight tiles arrive by TMA (cp.async.bulk.tensor, SWIZZLE_128B): no thread touches a weight.
//   * The activation-side operands (activations, dz, raw outputs) also arrive by TMA, as RAW fp32 tiles; the stager
//     warps read their rows from shared memory, apply the lazy BatchNorm + activation (forward), the BatchNorm-backward
//     affine map (gradients) and the hi / lo split in registers, and write the MMA's row operand into TMEM
//     (tcgen05.st).  No stager issues a global load: a thread-per-row global load touches 32 cache lines per
//     instruction and was what bounded the first generation.
//   * Warp roles: warps 0-7 stagers, warps 8-15 epilogue, warp 16 lane 0 issues every TMA, warp 17 lane 0 issues
//     every tcgen05.mma.  mbarrier rings connect them: weight stages (TMA -> MMA), raw-tile stages (TMA -> stagers),
//     TMEM operand stages (stagers -> MMA), accumulator buffers (MMA -> epilogue).
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
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace swr {
using namespace tc;

constexpr int T2_BM = 128;                 // accumulator rows (TMEM lanes) per tile
constexpr int T2_NSTAGER = 8;              // warps 0..7
constexpr int T2_NEPI = 8;                 // warps 8..15
constexpr int T2_W_TMA = 16, T2_W_MMA = 17;
constexpr int T2_THREADS = 20 * 32;        // five warpgroups: stagers (warps 0-3, 4-7), epilogue (8-11, 12-15), {TMA, MMA, two idle warps}
// Register budget per thread after the role split (setmaxnreg works on whole warpgroups).  The kernel starts with
// 96 registers x 640 threads = 61440; the epilogue warps, which keep a 64-column strip of the accumulator tile in
// registers across the flushes, take what the stagers and the two issuing warps give back:
// 2 x 128 x 80 + 128 x 48 + 2 x 128 x 136 = 61440 <= 61440.
constexpr int T2_REG_STAGER = 80, T2_REG_ISSUE = 48, T2_REG_EPI = 136;
// weight gradient: its stagers do more per k-block (both operands) and its epilogue keeps no prefetched rows:
// 2 x 128 x 96 + 128 x 48 + 2 x 128 x 112 = 59392
constexpr int T2W_REG_STAGER = 96, T2W_REG_EPI = 112;
constexpr int T2_MAX_STAGES = 4;
constexpr int T2_ASTAGES = 4;              // TMEM operand stages
constexpr uint32_t T2_ACC_COLS = 128;      // columns per accumulator buffer; buffers at TMEM columns 0 and 128
constexpr uint32_t T2_A_COL0 = 256;        // TMEM-resident operand: stage s at columns 256 + 64 s (32 hi + 32 lo)
constexpr uint32_t T2_TMEM_COLS = 512;
constexpr int T2_MAX_CTAS = 191;           // upper bound of the persistent grid (one CTA per SM)
constexpr int T2_MAX_PERM = 1024;          // launches with more tiles keep contiguous ranges
constexpr int T2_RAW_BYTES = T2_BM * 128;  // one raw fp32 tile: 128 rows x 32 contraction elements (or 32 batch rows x 128 features)
constexpr int T2_OT_LD = 36;               // epilogue transpose tile of one epilogue group: 128 rows x 32 columns (+4 pad)
constexpr int T2_OT_GROUP = T2_BM * T2_OT_LD;       // floats per group
constexpr int T2_OT_BYTES = 2 * T2_OT_GROUP * 4;
constexpr int T2_RED_GROUP = 2 * 4 * 32;            // [2 statistics][4 warps][32 columns] floats per group
constexpr int T2_RED_BYTES = 2 * T2_RED_GROUP * 4;
// named barriers: stager warps (256 threads), all epilogue warps (256), one epilogue group (128 threads, id 3 + group)
constexpr int T2_BAR_STAGE = 1, T2_BAR_EPI = 2, T2_BAR_EGRP = 3;
enum { T2_FWD = 0, T2_DGRAD = 1 };

struct alignas(64) Tc2Params {
  CUtensorMap tm0[kMaxGroups];     // fwd: weight image img_f;  dgrad: weight image img_d;  wgrad: dz   [B, N] (box 128 x 32)
  CUtensorMap tm1[kMaxGroups];     // fwd: input activation raw [B, K];  dgrad: dz [B, N];  wgrad: raw output [B, N]
  CUtensorMap tm2[kMaxGroups];     // dgrad: raw output [B, N];  wgrad: input activation raw [B, K] (box NT x 32)
  FcGroup g[kMaxGroups];
  int tile_start[kMaxGroups + 1];  // fwd / wgrad: first tile of group g;  dgrad: first k-block of group g
  int nt[kMaxGroups];              // accumulator columns per tile: per group (fwd, wgrad) / per destination (dgrad)
  int n_groups, B;
  float inv_count;
  int sb, sr, flush, n_tiles;      // weight (wgrad: column-operand) stages, raw-tile stages
  int stage_b, stage_r;            // bytes per stage of the two rings (sized for the widest tile of the launch)
  int off_r, off_coef, off_ccs, off_ot, off_red;   // shared-memory byte offsets (from the 1024-aligned base); ring B at 0
  int n_dst;
  int dst_group[kMaxGroups + 1];
  int dst_tile[kMaxGroups + 1];
  unsigned dst_atomic;             // bit d: destination entry d is one of several partial fan-ins: add atomically
  int splits, rows_per_split;      // wgrad
  int cluster;                     // fwd / dgrad: CTAs per cluster (1, 2 or 4) that share every weight tile by TMA multicast;
                                   // a tile index then names `cluster` consecutive row tiles, one per CTA rank
  unsigned short cta_tile[T2_MAX_CTAS + 1];   // balanced != 0: CTA (cluster) b walks positions [cta_tile[b], cta_tile[b + 1])
  unsigned short perm[T2_MAX_PERM];           // balanced == 2: position -> tile (tiles of mixed cost dealt out longest first)
  int balanced;
  long long* dbg;                  // development aid (SWR_TC_DEBUG): clock64 stamps of CTA 0, [role][128]; null in production
};

static_assert(sizeof(Tc2Params) <= 32764, "kernel parameters are limited to 32764 bytes");

// development aid: event `i` of role `role` (0 TMA, 1 MMA, 2 stager warp 0, 3 epilogue warp 8) of CTA 0
#define T2_STAMP(role, i) do { if (p.dbg && blockIdx.x == 0 && (i) < 128) p.dbg[(role) * 128 + (i)] = clock64(); } while (0)



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

constexpr int T2_TILE_CAP = 32;            // tiles one CTA walks (launches that need more fall back to FFMA)

struct Tc2Shared {
  uint64_t full_b[T2_MAX_STAGES];  // fwd/dgrad: TMA landed the weight tile
  uint64_t empty_b[T2_MAX_STAGES]; // tcgen05.commit: the MMAs that read the stage are done
  uint64_t full_r[T2_MAX_STAGES];  // TMA landed the raw tile(s)
  uint64_t empty_r[T2_MAX_STAGES]; // the stager warps have consumed them
  uint64_t full_a[T2_ASTAGES];     // the stager warps wrote the TMEM operand stage (wgrad: and the column-operand stage)
  uint64_t empty_a[T2_ASTAGES];    // tcgen05.commit
  uint64_t acc_full[2];            // tcgen05.commit: a partial accumulator is complete
  uint64_t acc_empty[2];           // the epilogue warps have drained it
  alignas(16) uint32_t tmem_base;
  uint32_t pad_[3];
  T2Tile tile[T2_TILE_CAP];        // this CTA's tiles, decoded once by an otherwise idle warp: every role reads them here (the
                                   // search + divisions of the decode are a few hundred instructions per inlined copy, and
                                   // code that runs once per tile is paid for in instruction-cache misses)
};

__device__ __forceinline__ uint8_t* t2_align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

struct T2Ring {                    // position in a ring of mbarrier-guarded stages
  int s; uint32_t ph;
  __device__ __forceinline__ void init() { s = 0; ph = 0; }
  __device__ __forceinline__ void next(int S) { if (++s == S) { s = 0; ph ^= 1u; } }
};


template <int MODE>
__device__ __forceinline__ T2Tile t2_decode(const Tc2Params& p, int t, int rank) {
  T2Tile T{};
  if (p.balanced == 2) t = p.perm[t];
  if (MODE == T2_FWD) {
    int g = 0;
    while (g + 1 < p.n_groups && p.tile_start[g + 1] <= t) ++g;
    const int N = p.g[g].Y.n, K = p.g[g].A.n;
    T.g = g; T.NT = p.nt[g];
    const int nt_n = (N + T.NT - 1) / T.NT, local = t - p.tile_start[g];
    T.m0 = ((local / nt_n) * p.cluster + rank) * T2_BM; T.n0 = (local % nt_n) * T.NT;
    T.nkb = (K + KBLK - 1) / KBLK;
  } else {
    int d = 0;
    while (d + 1 < p.n_dst && p.dst_tile[d + 1] <= t) ++d;
    T.d = d; T.g = p.dst_group[d]; T.ge = p.dst_group[d + 1];
    const int Kd = p.g[T.g].A.n;
    T.NT = p.nt[d];
    const int nt_n = (Kd + T.NT - 1) / T.NT, local = t - p.dst_tile[d];
    T.m0 = ((local / nt_n) * p.cluster + rank) * T2_BM; T.n0 = (local % nt_n) * T.NT;
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

__device__ __forceinline__ void t2_setup(Tc2Shared& sh, int tid, int warp, int stager_arrivals, int cluster) {
  if (tid == 0) {
    for (int s = 0; s < T2_MAX_STAGES; ++s) {
      mbar_init(&sh.full_b[s], 1); mbar_init(&sh.empty_b[s], cluster);   // every CTA of the cluster releases a multicast stage
      mbar_init(&sh.full_r[s], 1); mbar_init(&sh.empty_r[s], stager_arrivals);
    }
    for (int s = 0; s < T2_ASTAGES; ++s) { mbar_init(&sh.full_a[s], stager_arrivals); mbar_init(&sh.empty_a[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.acc_full[b], 1); mbar_init(&sh.acc_empty[b], T2_NEPI); }
    fence_mbar_init();
  }
  if (warp == T2_W_MMA) tmem_alloc(&sh.tmem_base, T2_TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  if (cluster > 1) cluster_sync_all();     // peers signal our barriers and write our stages: they must be initialised first
  fence_after_sync();
}

__device__ __forceinline__ void t2_range(const Tc2Params& p, int n_tiles, int cluster, int& t_begin, int& t_end) {
  const int nb = (int)gridDim.x / cluster, b = (int)blockIdx.x / cluster;
  if (p.balanced) { t_begin = p.cta_tile[b]; t_end = p.cta_tile[b + 1]; return; }
  const int per = n_tiles / nb, rem = n_tiles % nb;
  t_begin = b * per + min(b, rem);
  t_end = t_begin + per + (b < rem ? 1 : 0);
}

__device__ __forceinline__ float4 t2_ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 t2_zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void t2_red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float t2_slope(int act) { return act == SWR_ACT_RELU ? 0.f : (act == SWR_ACT_LEAKY ? 0.1f : 1.f); }

// split 16 values and store them as this thread's 16 hi + 16 lo columns of an operand stage
__device__ __forceinline__ void t2_store_split16(uint32_t taddr_hi, const float (&v)[16]) {
  float hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) split_tf32(v[i], hi[i], lo[i]);
  tmem_st16(taddr_hi, hi);
  tmem_st16(taddr_hi + 32, lo);
}
// same for elements [o, o + 16) of a 32-element row
__device__ __forceinline__ void t2_store_split16(uint32_t taddr_hi, const float (&v)[32], int o) {
  float hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) split_tf32(v[o + i], hi[i], lo[i]);
  tmem_st16(taddr_hi, hi);
  tmem_st16(taddr_hi + 32, lo);
}
__device__ __noinline__ float4 t2_sigmoid4(float4 z) {
  return make_float4(1.f / (1.f + expf(-z.x)), 1.f / (1.f + expf(-z.y)), 1.f / (1.f + expf(-z.z)), 1.f / (1.f + expf(-z.w)));
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// this thread's 16 contraction elements of row `row` (chunks 4 kh .. 4 kh + 3) of a K-major swizzled raw tile
__device__ __forceinline__ void t2_lds_row16(const uint8_t* tile, int row, int kh, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 x = *reinterpret_cast<const float4*>(tile + kmajor_off(row, 4 * kh + i));
    v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
  }
}

// ---- epilogue helpers -------------------------------------------------------------------------------------
// drain this warp's share (lane quarter q, columns [64 ch, 64 ch + 64) of an NT-wide accumulator) and add it to acc:
// sixteen columns per tcgen05.ld (the epilogue warpgroups run with T2_REG_EPI registers per thread).
__device__ __forceinline__ void t2_drain_add(uint32_t tmem_acc, int q, int ch, int NT, float (&acc)[64]) {
  const int my = min(max(NT - 64 * ch, 0), 64);       // 0, 16, 32, 48 or 64 columns (warp-uniform)
  const uint32_t taddr = tmem_acc + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * ch);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (16 * c < my) {
      uint32_t r[16];
      tmem_ld16(taddr + 16 * c, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[16 * c + i] += __uint_as_float(r[i]);
    }
  }
}
// 32 accumulator columns [32 sp, 32 sp + 32) of this thread's row -> the group's transpose tile (shared-space address)
__device__ __forceinline__ void t2_acc_to_smem(uint32_t ot_s, int row, int sp, int pc, const float (&acc)[64]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (4 * i < pc) {
      const float4 v = sp == 0 ? make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3])
                               : make_float4(acc[32 + 4 * i], acc[32 + 4 * i + 1], acc[32 + 4 * i + 2], acc[32 + 4 * i + 3]);
      sts128(ot_s + (uint32_t)(row * T2_OT_LD + 4 * i) * 4u, v);
    }
}
// bias + bias2 of output columns [n, n + 4) that lie below n_end (zero elsewhere)
__device__ __forceinline__ float4 t2_bias_quad(const FcGroup& G, int n, int n_end) {
  float4 b = t2_zero4();
  if (n < n_end) b.x = ld_opt(G.bias, n, 0.f) + ld_opt(G.bias2, n, 0.f);
  if (n + 1 < n_end) b.y = ld_opt(G.bias, n + 1, 0.f) + ld_opt(G.bias2, n + 1, 0.f);
  if (n + 2 < n_end) b.z = ld_opt(G.bias, n + 2, 0.f) + ld_opt(G.bias2, n + 2, 0.f);
  if (n + 3 < n_end) b.w = ld_opt(G.bias, n + 3, 0.f) + ld_opt(G.bias2, n + 3, 0.f);
  return b;
}
// forward epilogue value of one accumulator element: bias, optional activation of layers without a norm (GateNU)
__device__ __forceinline__ float t2_epi_val(float a, float bias, int e_act, float e_scale) {
  const float y = a + bias;
  return e_act == SWR_ACT_NONE ? y : act_fwd(y, e_act) * e_scale;
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
  const int SB = p.sb, SR = p.sr, F = p.flush, M = p.B;
  int t_begin, t_end;
  const int C = p.cluster, crank = C > 1 ? (int)cluster_ctarank() : 0;
  t2_range(p, p.n_tiles, C, t_begin, t_end);
  if (tid == 0) T2_STAMP(0, 126);
  if (warp == T2_W_MMA + 1)
    for (int i = lane; i < t_end - t_begin; i += 32) sh.tile[i] = t2_decode<MODE>(p, t_begin + i, crank);
  if (warp == T2_W_TMA && lane < p.n_groups) {       // descriptor fetches overlap the barrier / TMEM setup
    tma_prefetch_desc(&p.tm0[lane]); tma_prefetch_desc(&p.tm1[lane]);
    if (MODE == T2_DGRAD && p.g[lane].Y.norm.mode == SWR_NORM_BATCH) tma_prefetch_desc(&p.tm2[lane]);
  }
  t2_setup(sh, tid, warp, T2_NSTAGER / 2, C);
  const uint32_t tmem = sh.tmem_base;
  if (tid == 0) T2_STAMP(0, 127);

  // every role branch opens with its warpgroup's setmaxnreg, so that ptxas allocates the branch under that budget
  if (warp >= T2_NSTAGER + T2_NEPI) {
   reg_dec<T2_REG_ISSUE>();
   if (warp == T2_W_TMA) {
    // ===== TMA producer: weight tiles (ring B) and raw activation-side tiles (ring R) =====
    // The whole warp walks the loop (so every operand is warp-uniform); one elected lane issues.  The two rings
    // advance independently: raw tiles feed the longer chain (stagers -> TMEM -> MMA) and run ahead as far as their
    // ring allows; a cursor is a (tile, group, k-block) position in the CTA's k-block sequence.
    struct Cursor { int t, g, ge, kb, nk, m0, n0, NT; T2Ring r; bool done; };
    auto enter = [&](Cursor& c) {
      const T2Tile T = sh.tile[c.t - t_begin];
      c.g = T.g; c.ge = T.ge; c.m0 = T.m0; c.n0 = T.n0; c.NT = T.NT; c.kb = 0;
      c.nk = (MODE == T2_FWD) ? T.nkb : p.tile_start[c.g + 1] - p.tile_start[c.g];
    };
    auto start = [&](Cursor& c) {
      c.t = t_begin; c.done = (t_begin >= t_end); c.r.init();
      if (!c.done) enter(c);
    };
    auto advance = [&](Cursor& c, int S) {
      c.r.next(S);
      if (++c.kb < c.nk) return;
      c.kb = 0;
      if (MODE == T2_DGRAD && ++c.g < c.ge) { c.nk = p.tile_start[c.g + 1] - p.tile_start[c.g]; return; }
      if (++c.t >= t_end) { c.done = true; return; }
      enter(c);
    };
    Cursor cb, cr;
    start(cb); start(cr);
    int ev = 0;
    while (!cb.done || !cr.done) {
      // raw tiles first; block on a ring only when the other one has nothing to issue either
      const bool r_ready = !cr.done && mbar_test(&sh.empty_r[cr.r.s], cr.r.ph ^ 1u);
      const bool b_ready = !cb.done && mbar_test(&sh.empty_b[cb.r.s], cb.r.ph ^ 1u);
      if (!r_ready && !b_ready) {
        if (!cr.done) mbar_wait(&sh.empty_r[cr.r.s], cr.r.ph ^ 1u);
        else mbar_wait(&sh.empty_b[cb.r.s], cb.r.ph ^ 1u);
        continue;
      }
      if (r_ready) {
        const bool two = (MODE == T2_DGRAD) && (p.g[cr.g].Y.norm.mode == SWR_NORM_BATCH);
        if (elect_one()) {
          mbar_expect_tx(&sh.full_r[cr.r.s], two ? 2u * T2_RAW_BYTES : (uint32_t)T2_RAW_BYTES);
          const uint32_t rdst = smem_s + (uint32_t)(p.off_r + cr.r.s * p.stage_r);
          tma_load_2d(rdst, &p.tm1[cr.g], cr.kb * KBLK, cr.m0, &sh.full_r[cr.r.s]);
          if (two) tma_load_2d(rdst + T2_RAW_BYTES, &p.tm2[cr.g], cr.kb * KBLK, cr.m0, &sh.full_r[cr.r.s]);
        }
        __syncwarp();
        advance(cr, SR);
      }
      if (b_ready) {
        if (elect_one()) {
          mbar_expect_tx(&sh.full_b[cb.r.s], 2u * (uint32_t)cb.NT * 128u);
          const uint32_t bdst = smem_s + (uint32_t)(cb.r.s * p.stage_b);
          if (C == 1) {
            tma_load_3d(bdst, &p.tm0[cb.g], cb.kb * KBLK, cb.n0, 0, &sh.full_b[cb.r.s]);
          } else {      // this CTA fetches rows [crank, crank + 1) * NT / C of both planes for every CTA of the cluster
            const int rows = cb.NT / C;
            const uint32_t so = (uint32_t)(crank * rows) * 128u;
            const uint16_t mask = (uint16_t)((1u << C) - 1u);
            tma_load_3d_mc(bdst + so, &p.tm0[cb.g], cb.kb * KBLK, cb.n0 + crank * rows, 0, &sh.full_b[cb.r.s], mask);
            tma_load_3d_mc(bdst + (uint32_t)cb.NT * 128u + so, &p.tm0[cb.g], cb.kb * KBLK, cb.n0 + crank * rows, 1, &sh.full_b[cb.r.s], mask);
          }
        }
        __syncwarp();
        if (lane == 0) { T2_STAMP(0, ev); } ++ev;
        advance(cb, SB);
      }
    }
   } else if (warp == T2_W_MMA) {
    // ===== MMA issuer: converged warp, one elected lane issues the tcgen05.mma / commit instructions =====
    {
      T2Ring rb, ra; rb.init(); ra.init();
      uint32_t acc_it = 0;
      int ev = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const T2Tile T = sh.tile[t - t_begin];
        const uint32_t b_bytes = (uint32_t)T.NT * 128u;
        const uint32_t idesc = make_idesc_tf32(T2_BM, T.NT, false, false);
        int fpos = 0;
        for (int kb = 0; kb < T.nkb; ++kb, rb.next(SB), ra.next(T2_ASTAGES)) {
          const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
          if (fpos == 0) mbar_wait(&sh.acc_empty[buf], aph ^ 1u);
          mbar_wait(&sh.full_b[rb.s], rb.ph);
          mbar_wait(&sh.full_a[ra.s], ra.ph);
          fence_after_sync();
          const bool last = (fpos + 1 == F) || (kb == T.nkb - 1);
          if (elect_one()) {
            t2_issue(tmem + buf * T2_ACC_COLS, tmem + T2_A_COL0 + (uint32_t)ra.s * 64u, smem_s + (uint32_t)(rb.s * p.stage_b), b_bytes,
                     false, idesc, fpos == 0);
            if (C == 1) mma_commit(&sh.empty_b[rb.s]); else mma_commit_mc(&sh.empty_b[rb.s], (uint16_t)((1u << C) - 1u));
            mma_commit(&sh.empty_a[ra.s]);
            if (last) mma_commit(&sh.acc_full[buf]);
          }
          __syncwarp();
          if (last) { ++acc_it; fpos = 0; } else ++fpos;
          if (lane == 0) { T2_STAMP(1, ev); } ++ev;
        }
      }
    }
   }
  } else if (warp < T2_NSTAGER) {
    reg_dec<T2_REG_STAGER>();
    // ===== stagers: raw tile (shared memory) -> transform -> split -> TMEM =====
    // Two groups of four warps (one warp per TMEM lane quarter) alternate over the k-blocks of the CTA, so that the
    // handshake latencies of consecutive k-blocks overlap; a thread stages the 32 contraction elements of its row.
    const int q = warp & 3, grp = warp >> 2, stid = tid;
    const int row = 32 * q + lane;
    const uint32_t ta = tmem + T2_A_COL0 + ((uint32_t)(32 * q) << 16);
    float* coef = reinterpret_cast<float*>(smem + p.off_coef);
    T2Ring rr, ra; rr.init(); ra.init();
    if (grp) { rr.next(SR); ra.next(T2_ASTAGES); }
    int kbg = 0, cur_key = -1, ev = 0;     // kbg: k-blocks of the CTA before this tile (its parity picks the group)
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = sh.tile[t - t_begin];
      const int Kc = T.nkb * KBLK;         // padded contraction length of the tile (fwd: K; dgrad: the whole fan-in)
      bool plainA = false, sig = false;
      float slope = 1.f;
      if (MODE == T2_FWD) {
        const FcGroup& G = p.g[T.g];
        const int K = G.A.n;
        plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
        slope = t2_slope(G.A.act); sig = G.A.act == SWR_ACT_SIGMOID;
        if (!plainA && cur_key != T.g) {            // [3][Kc]: mu, s, b of the input columns (zero beyond K)
          named_bar(T2_BAR_STAGE, 32 * T2_NSTAGER);
          for (int k = stid; k < Kc; k += 32 * T2_NSTAGER) {
            ColCoef c = {0.f, 0.f, 0.f, 0.f};
            if (k < K) c = col_coef(G.A.norm, k, p.inv_count);
            coef[k] = c.mu; coef[Kc + k] = c.s; coef[2 * Kc + k] = c.b;
          }
          named_bar(T2_BAR_STAGE, 32 * T2_NSTAGER);
          cur_key = T.g;
        }
      } else if (cur_key != T.d) {
        // coefficients of dY = c0 * dz + c1 * raw + c2 over the concatenated output columns of the fan-in
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
      int g = T.g, g_end_kb = (MODE == T2_DGRAD) ? p.tile_start[T.g + 1] - T.kb0 : T.nkb;
      bool two = (MODE == T2_DGRAD) && p.g[g].Y.norm.mode == SWR_NORM_BATCH;
      for (int kb = 0; kb < T.nkb; ++kb) {
        if (MODE == T2_DGRAD) {
          while (kb >= g_end_kb) { ++g; g_end_kb = p.tile_start[g + 1] - T.kb0; two = p.g[g].Y.norm.mode == SWR_NORM_BATCH; }
        }
        if (((kbg + kb) & 1) != grp) continue;
        mbar_wait(&sh.full_r[rr.s], rr.ph);
        mbar_wait(&sh.empty_a[ra.s], ra.ph ^ 1u);
        fence_after_sync();
        if (tid == 0) { T2_STAMP(2, ev); ++ev; }
        const uint32_t rt = smem_s + (uint32_t)(p.off_r + rr.s * p.stage_r);
        const uint32_t tst = ta + (uint32_t)ra.s * 64u;
        const float* c = coef + kb * KBLK;
        // two halves of 16 contraction elements, each taken from shared memory to TMEM before the next one starts
        // (keeps the live registers of the loop near 60)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 x = lds128(rt + kmajor_off(row, 4 * h + i));
            v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
          }
          const float* ch_ = c + 16 * h;
          if (MODE == T2_FWD) {
            if (!plainA) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 mu = t2_ld4s(ch_ + 4 * i), sc = t2_ld4s(ch_ + Kc + 4 * i), bb = t2_ld4s(ch_ + 2 * Kc + 4 * i);
                v[4 * i] = fmaf(v[4 * i] - mu.x, sc.x, bb.x); v[4 * i + 1] = fmaf(v[4 * i + 1] - mu.y, sc.y, bb.y);
                v[4 * i + 2] = fmaf(v[4 * i + 2] - mu.z, sc.z, bb.z); v[4 * i + 3] = fmaf(v[4 * i + 3] - mu.w, sc.w, bb.w);
              }
              if (sig) {          // rare as an input activation: out of line, four values per call
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 y = t2_sigmoid4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                  v[4 * i] = y.x; v[4 * i + 1] = y.y; v[4 * i + 2] = y.z; v[4 * i + 3] = y.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], slope * v[i]);
              }
            }
          } else if (two) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 w = lds128(rt + T2_RAW_BYTES + kmajor_off(row, 4 * h + i));
              const float4 c0 = t2_ld4s(ch_ + 4 * i), c1 = t2_ld4s(ch_ + Kc + 4 * i), c2 = t2_ld4s(ch_ + 2 * Kc + 4 * i);
              v[4 * i] = fmaf(c0.x, v[4 * i], fmaf(c1.x, w.x, c2.x)); v[4 * i + 1] = fmaf(c0.y, v[4 * i + 1], fmaf(c1.y, w.y, c2.y));
              v[4 * i + 2] = fmaf(c0.z, v[4 * i + 2], fmaf(c1.z, w.z, c2.z)); v[4 * i + 3] = fmaf(c0.w, v[4 * i + 3], fmaf(c1.w, w.w, c2.w));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 c0 = t2_ld4s(ch_ + 4 * i), c2 = t2_ld4s(ch_ + 2 * Kc + 4 * i);
              v[4 * i] = fmaf(c0.x, v[4 * i], c2.x); v[4 * i + 1] = fmaf(c0.y, v[4 * i + 1], c2.y);
              v[4 * i + 2] = fmaf(c0.z, v[4 * i + 2], c2.z); v[4 * i + 3] = fmaf(c0.w, v[4 * i + 3], c2.w);
            }
          }
          t2_store_split16(tst + 16 * h, v);
          asm volatile("" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty_r[rr.s]);   // every lane's values left shared memory before its split
        tmem_st_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.full_a[ra.s]);
        if (tid == 0) { T2_STAMP(2, ev); ++ev; }
        rr.next(SR); rr.next(SR); ra.next(T2_ASTAGES); ra.next(T2_ASTAGES);
      }
      kbg += T.nkb;
    }
  } else {
    reg_inc<T2_REG_EPI>();
    // ===== epilogue warps =====
    // Two independent groups of four warps (one per TMEM lane quarter): group ch owns accumulator columns
    // [64 ch, 64 ch + 64), drains them into registers flush by flush, then finishes them in two 32-column passes
    // through its own transpose tile.  What a pass needs from global memory (the destination's raw values in the data
    // gradient, the bias in the forward) is requested before the accumulators are waited for, so that its latency
    // hides behind the MMAs / the previous pass.
    const int e = warp - T2_NSTAGER, q = warp & 3, ch = e >> 2;
    uint32_t acc_it = 0;
    int ev = 0;
    for (int t = t_begin; t < t_end; ++t) {
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      float4 pre[4];     // what a pass needs from global memory: the next raw rows of the destination (dgrad) / pre[0] = bias quad (fwd)
      {
        // Only what the drain loop needs is decoded here (the tile is decoded again below): the running sums take 64
        // registers, and everything else that lives across the loop pushes them into local memory.
        const T2Tile T0 = sh.tile[t - t_begin];
        if (MODE == T2_DGRAD) {
          // coefficients of the destination's own norm / activation: [4][NT] mu, s, b, r
          const ActDev& D0 = p.g[T0.g].A;
          const bool plain0 = D0.norm.mode == SWR_NORM_NONE && D0.act == SWR_ACT_NONE;
          const int Nend0 = min(D0.n, T0.n0 + T0.NT);
          float* ccs = reinterpret_cast<float*>(smem + p.off_ccs);
          named_bar(T2_BAR_EPI, 32 * T2_NEPI);          // every epilogue warp is done with the previous tile's table
          for (int c = tid - 32 * T2_NSTAGER; c < T0.NT; c += 32 * T2_NEPI) {
            ColCoef cc = {0.f, 1.f, 0.f, 1.f};
            if (!plain0 && T0.n0 + c < Nend0) cc = col_coef(D0.norm, T0.n0 + c, p.inv_count);
            ccs[c] = cc.mu; ccs[T0.NT + c] = cc.s; ccs[2 * T0.NT + c] = cc.b; ccs[3 * T0.NT + c] = cc.r;
          }
          named_bar(T2_BAR_EPI, 32 * T2_NEPI);
        } else {
          pre[0] = t2_bias_quad(p.g[T0.g], T0.n0 + 64 * ch + (lane & 7) * 4, min(p.g[T0.g].Y.n, T0.n0 + T0.NT));     // pass 0 of this group
        }
        const int NT0 = T0.NT, nflush = (T0.nkb + F - 1) / F;
        for (int f = 0; f < nflush; ++f, ++acc_it) {
          const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
          mbar_wait(&sh.acc_full[buf], aph);
          fence_after_sync();
          if (tid == 32 * T2_NSTAGER) { T2_STAMP(3, ev); ++ev; }
          t2_drain_add(tmem + buf * T2_ACC_COLS, q, ch, NT0, acc);
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh.acc_empty[buf]);
        }
      }
      if (tid == 32 * T2_NSTAGER) { T2_STAMP(3, ev); ++ev; }
      int t_again = t;
      asm volatile("" : "+r"(t_again));              // keeps the decode below from being carried through the drain loop
      const T2Tile T = sh.tile[t_again - t_begin];
      const int gtid = (tid - 32 * T2_NSTAGER) & 127;
      const int arow = 32 * q + lane;                  // accumulator row this thread drains
      const int rsub = lane >> 3, c4 = (lane & 7) * 4;  // coalesced pass: 8 lanes per row, 4 rows per warp step
      const int bar_g = T2_BAR_EGRP + ch;
      const int row_base = 32 * q + rsub;              // this lane's rows: row_base + 4 i
      float* ot = reinterpret_cast<float*>(smem + p.off_ot) + ch * T2_OT_GROUP;
      const uint32_t ot_s = smem_s + (uint32_t)p.off_ot + (uint32_t)(ch * T2_OT_GROUP) * 4u;
      float* red = reinterpret_cast<float*>(smem + p.off_red) + ch * T2_RED_GROUP;     // [2][4 warps][32]
      float* ccs = reinterpret_cast<float*>(smem + p.off_ccs);
      const uint32_t my_ot = ot_s + (uint32_t)(row_base * T2_OT_LD + c4) * 4u;         // + i * 4 rows
      const FcGroup& G = p.g[T.g];
      const ActDev& D = G.A;                          // dgrad: the destination
      const int Nfull = (MODE == T2_FWD) ? G.Y.n : D.n;
      const int Nend = min(Nfull, T.n0 + T.NT);
      const bool has_norm = (MODE == T2_DGRAD) && D.norm.mode != SWR_NORM_NONE;
      const bool plainD = (MODE != T2_DGRAD) || (!has_norm && D.act == SWR_ACT_NONE);
      const bool atomic_dst = (MODE == T2_DGRAD) && ((p.dst_atomic >> T.d) & 1u);
      const bool accumulate = (MODE == T2_DGRAD) && (G.flags & FC_A_ACCUMULATE) != 0;
      const int cnt_rows = min(T2_BM, M - T.m0);
      const int rows_here = min((max(cnt_rows - row_base, 0) + 3) >> 2, 8);
      const int my_passes = min(max(T.NT - 64 * ch, 0) + 31, 64) >> 5;     // 32-column passes of this group (0, 1 or 2)
      const bool base_al = (MODE == T2_FWD) ? ((G.Y.ld % 4 == 0) && is_al16(G.Y.raw))
                                            : ((D.ld % 4 == 0) && is_al16(D.dz) && is_al16(D.raw));
      // Requests of pass sp, issued ahead of its use: the forward's bias quad (pass 0: before the drain loop, above), the
      // first four raw rows of the data gradient (pass 0: here; pass 1: behind the row loop of pass 0, hidden by its
      // statistics tail); the row loop keeps four rows in flight.
      // The row loops below are deliberately NOT unrolled: this code runs once or twice per tile, so every instruction is
      // an instruction-cache miss the first time round -- straight-line code ran at ~40 cycles per instruction here
      // (stall_no_inst in profiles/r02y), a rolled loop pays that once.
      bool pre_ok = false;
      auto prefetch = [&](int sp) {
        pre_ok = false;
        if (sp >= my_passes) return;
        const int pc0 = 64 * ch + 32 * sp;
        const int n = T.n0 + pc0 + c4, nv = Nend - n;
        if (MODE == T2_FWD) {
          pre[0] = t2_bias_quad(G, n, Nend);
        } else if (!plainD && !atomic_dst && nv >= 4 && base_al && (n % 4 == 0)) {
          const float* src = D.raw + (int64_t)(T.m0 + row_base) * D.ld + n;
          const int64_t ostep = 4 * (int64_t)D.ld;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < rows_here) pre[i] = *reinterpret_cast<const float4*>(src + i * ostep);
          pre_ok = true;
        }
      };
      if (MODE == T2_DGRAD) prefetch(0);
      // ---- final epilogue of this group's columns, 32 per pass ----
      // Column statistics stay in fp32 until one thread per column widens them (fp64 throughput is a small fraction
      // of fp32's): forward moments are sums of deviations from the tile's first row, a centre every thread can read.
#pragma unroll 1
      for (int sp = 0; sp < my_passes; ++sp) {
        const int pc0 = 64 * ch + 32 * sp;
        const int pc = min(32, T.NT - pc0);
        const int nvalid = min(max(Nend - (T.n0 + pc0), 0), pc);
        named_bar(bar_g, 128);                       // the group is done with the previous contents of ot / red
        t2_acc_to_smem(ot_s, arow, sp, pc, acc);
        named_bar(bar_g, 128);
        if (tid == 32 * T2_NSTAGER) { T2_STAMP(3, ev); ++ev; }
        const int nv = nvalid - c4;                  // valid components of this lane's column quad (<= 0: none)
        const int n = T.n0 + pc0 + c4;               // first output column of the quad
        float4 s1 = t2_zero4(), s2 = t2_zero4();
        if (MODE == T2_FWD) {
          float* Y = const_cast<float*>(G.Y.raw);
          const bool vec = (nv >= 4) && base_al && (n % 4 == 0);
          const float4 bias = pre[0];
          if (nv > 0) {
            float* dst = Y + (int64_t)(T.m0 + row_base) * G.Y.ld + n;
            const int64_t dstep = 4 * (int64_t)G.Y.ld;
            const float4 r0 = lds128(ot_s + (uint32_t)c4 * 4u);      // row 0 of the tile: always a valid batch row
            if (G.e_act == SWR_ACT_NONE && vec) {   // the common case, branch-free per row
              const float4 y0 = make_float4(r0.x + bias.x, r0.y + bias.y, r0.z + bias.z, r0.w + bias.w);
#pragma unroll 1
              for (int i = 0; i < rows_here; ++i) {
                const float4 a = lds128(my_ot + (uint32_t)(4 * i * T2_OT_LD) * 4u);
                const float4 y = make_float4(a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w);
                *reinterpret_cast<float4*>(dst + i * dstep) = y;
                const float4 d = make_float4(y.x - y0.x, y.y - y0.y, y.z - y0.z, y.w - y0.w);
                s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
                s2.x = fmaf(d.x, d.x, s2.x); s2.y = fmaf(d.y, d.y, s2.y); s2.z = fmaf(d.z, d.z, s2.z); s2.w = fmaf(d.w, d.w, s2.w);
              }
            } else {
              const float4 y0 = make_float4(t2_epi_val(r0.x, bias.x, G.e_act, G.e_scale), t2_epi_val(r0.y, bias.y, G.e_act, G.e_scale),
                                            t2_epi_val(r0.z, bias.z, G.e_act, G.e_scale), t2_epi_val(r0.w, bias.w, G.e_act, G.e_scale));
#pragma unroll 1
              for (int i = 0; i < rows_here; ++i) {
                const float4 a = lds128(my_ot + (uint32_t)(4 * i * T2_OT_LD) * 4u);
                const float4 y = make_float4(t2_epi_val(a.x, bias.x, G.e_act, G.e_scale), t2_epi_val(a.y, bias.y, G.e_act, G.e_scale),
                                             t2_epi_val(a.z, bias.z, G.e_act, G.e_scale), t2_epi_val(a.w, bias.w, G.e_act, G.e_scale));
                float* dp = dst + i * dstep;
                if (vec) *reinterpret_cast<float4*>(dp) = y;
                else { dp[0] = y.x; if (nv > 1) dp[1] = y.y; if (nv > 2) dp[2] = y.z; if (nv > 3) dp[3] = y.w; }
                const float4 d = make_float4(y.x - y0.x, y.y - y0.y, y.z - y0.z, y.w - y0.w);
                s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
                s2.x = fmaf(d.x, d.x, s2.x); s2.y = fmaf(d.y, d.y, s2.y); s2.z = fmaf(d.z, d.z, s2.z); s2.w = fmaf(d.w, d.w, s2.w);
              }
            }
          }
          prefetch(sp + 1);
        } else {
          const bool vec = (nv >= 4) && base_al && (n % 4 == 0);
          if (nv > 0) {
            const int cc0 = pc0 + c4;
            const int64_t o0 = (int64_t)(T.m0 + row_base) * D.ld + n, ostep = 4 * (int64_t)D.ld;
            const bool full8 = rows_here == 8 && vec && !accumulate;
            if (full8 && atomic_dst && plainD) {
              // partial fan-in of a plain destination (the embedding output at level 0), full tile: transposed row -> one
              // 16-byte reduction, nothing else
              float* dp = D.dz + o0;
#pragma unroll 1
              for (int i = 0; i < 8; ++i, dp += ostep) t2_red_add_v4(dp, lds128(my_ot + (uint32_t)(4 * i * T2_OT_LD) * 4u));
            } else if (full8 && !atomic_dst && pre_ok && D.act != SWR_ACT_SIGMOID) {
              // BatchNorm / ReLU-family destination, full tile: the hot case.  Two trips of four rows without guards; pointer
              // increments instead of per-row address arithmetic (the general loop below spends most of its ~100
              // instructions per row on addressing and divergence bookkeeping, and two warps per scheduler cannot hide it)
              const float4 mu = t2_ld4s(ccs + cc0), sc = t2_ld4s(ccs + T.NT + cc0), bb = t2_ld4s(ccs + 2 * T.NT + cc0), rr4 = t2_ld4s(ccs + 3 * T.NT + cc0);
              const float slope = t2_slope(D.act);
              const float* rp = D.raw + o0 + 4 * ostep;       // rows 4..7 (rows 0..3 are in pre[])
              float* dp = D.dz + o0;
              uint32_t op = my_ot;
#pragma unroll 1
              for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float4 dz = lds128(op);
                  const float4 raw = pre[j];
                  if (h == 0) pre[j] = *reinterpret_cast<const float4*>(rp);
                  const float4 xc = make_float4(raw.x - mu.x, raw.y - mu.y, raw.z - mu.z, raw.w - mu.w);
                  dz.x *= fmaf(xc.x, sc.x, bb.x) > 0.f ? 1.f : slope; dz.y *= fmaf(xc.y, sc.y, bb.y) > 0.f ? 1.f : slope;
                  dz.z *= fmaf(xc.z, sc.z, bb.z) > 0.f ? 1.f : slope; dz.w *= fmaf(xc.w, sc.w, bb.w) > 0.f ? 1.f : slope;
                  s1.x += dz.x; s1.y += dz.y; s1.z += dz.z; s1.w += dz.w;
                  s2.x = fmaf(dz.x, xc.x * rr4.x, s2.x); s2.y = fmaf(dz.y, xc.y * rr4.y, s2.y);
                  s2.z = fmaf(dz.z, xc.z * rr4.z, s2.z); s2.w = fmaf(dz.w, xc.w * rr4.w, s2.w);
                  *reinterpret_cast<float4*>(dp) = dz;
                  dp += ostep; rp += ostep; op += (uint32_t)(4 * T2_OT_LD) * 4u;
                }
              }
            } else if (vec && !atomic_dst) {       // the general aligned case
              const float4 mu = t2_ld4s(ccs + cc0), sc = t2_ld4s(ccs + T.NT + cc0), bb = t2_ld4s(ccs + 2 * T.NT + cc0), rr4 = t2_ld4s(ccs + 3 * T.NT + cc0);
              const float slope = t2_slope(D.act);
              const bool sigD = D.act == SWR_ACT_SIGMOID;
              const float* rsrc = D.raw + o0;
              // four rows per trip, so that the in-flight raw rows keep static register names (a rotation through moves
              // would wait for the youngest load every iteration)
#pragma unroll 1
              for (int i0 = 0; i0 < rows_here; i0 += 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int i = i0 + j;
                  if (i < rows_here) {
                    float4 dz = lds128(my_ot + (uint32_t)(4 * i * T2_OT_LD) * 4u);
                    float* dst = D.dz + o0 + i * ostep;
                    if (!plainD) {
                      float4 raw;
                      if (pre_ok) {       // take row i, ask for row i + 4
                        raw = pre[j];
                        if (i + 4 < rows_here) pre[j] = *reinterpret_cast<const float4*>(rsrc + (i + 4) * ostep);
                      } else {
                        raw = *reinterpret_cast<const float4*>(rsrc + i * ostep);
                      }
                      const float4 xc = make_float4(raw.x - mu.x, raw.y - mu.y, raw.z - mu.z, raw.w - mu.w);
                      const float4 z = make_float4(fmaf(xc.x, sc.x, bb.x), fmaf(xc.y, sc.y, bb.y), fmaf(xc.z, sc.z, bb.z), fmaf(xc.w, sc.w, bb.w));
                      if (sigD) {
                        dz.x *= act_grad(z.x, SWR_ACT_SIGMOID); dz.y *= act_grad(z.y, SWR_ACT_SIGMOID);
                        dz.z *= act_grad(z.z, SWR_ACT_SIGMOID); dz.w *= act_grad(z.w, SWR_ACT_SIGMOID);
                      } else {       // d max(z, slope z) / dz
                        dz.x *= z.x > 0.f ? 1.f : slope; dz.y *= z.y > 0.f ? 1.f : slope;
                        dz.z *= z.z > 0.f ? 1.f : slope; dz.w *= z.w > 0.f ? 1.f : slope;
                      }
                      s1.x += dz.x; s1.y += dz.y; s1.z += dz.z; s1.w += dz.w;
                      s2.x = fmaf(dz.x, xc.x * rr4.x, s2.x); s2.y = fmaf(dz.y, xc.y * rr4.y, s2.y);
                      s2.z = fmaf(dz.z, xc.z * rr4.z, s2.z); s2.w = fmaf(dz.w, xc.w * rr4.w, s2.w);
                    }
                    if (accumulate) { const float4 old = *reinterpret_cast<const float4*>(dst); dz.x += old.x; dz.y += old.y; dz.z += old.z; dz.w += old.w; }
                    *reinterpret_cast<float4*>(dst) = dz;
                  }
                }
              }
            } else {
              const float mu[4] = {ccs[cc0], ccs[cc0 + 1], ccs[cc0 + 2], ccs[cc0 + 3]};
              const float sc[4] = {ccs[T.NT + cc0], ccs[T.NT + cc0 + 1], ccs[T.NT + cc0 + 2], ccs[T.NT + cc0 + 3]};
              const float bb[4] = {ccs[2 * T.NT + cc0], ccs[2 * T.NT + cc0 + 1], ccs[2 * T.NT + cc0 + 2], ccs[2 * T.NT + cc0 + 3]};
              const float rr_[4] = {ccs[3 * T.NT + cc0], ccs[3 * T.NT + cc0 + 1], ccs[3 * T.NT + cc0 + 2], ccs[3 * T.NT + cc0 + 3]};
              float t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
              for (int i = 0; i < rows_here; ++i) {
                const float4 a4 = lds128(my_ot + (uint32_t)(4 * i * T2_OT_LD) * 4u);
                float dz[4] = {a4.x, a4.y, a4.z, a4.w};
                const int64_t o = o0 + i * ostep;
                if (!plainD) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float raw = (j < nv) ? D.raw[o + j] : 0.f;
                    dz[j] *= act_grad(fmaf(raw - mu[j], sc[j], bb[j]), D.act);
                    t1[j] += dz[j];
                    t2[j] = fmaf(dz[j], (raw - mu[j]) * rr_[j], t2[j]);
                  }
                }
                float* dst = D.dz + o;
                if (atomic_dst) {                     // plain destination split over its fan-in
                  if (vec) t2_red_add_v4(dst, make_float4(dz[0], dz[1], dz[2], dz[3]));
                  else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < nv) atomicAdd(dst + j, dz[j]);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) if (j < nv) dst[j] = accumulate ? dst[j] + dz[j] : dz[j];
                }
              }
              s1 = make_float4(t1[0], t1[1], t1[2], t1[3]); s2 = make_float4(t2[0], t2[1], t2[2], t2[3]);
            }
          }
          prefetch(sp + 1);
        }
        if (tid == 32 * T2_NSTAGER) { T2_STAMP(3, ev); ++ev; }
        const bool want_stats = (MODE == T2_FWD) ? (G.stats_out != nullptr) : (has_norm && D.dstats != nullptr);
        if (want_stats) {           // uniform over the group
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {          // the four row sub-groups of the warp
            s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
            s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
            s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
            s2.z += __shfl_xor_sync(0xffffffffu, s2.z, o); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, o);
          }
          if (rsub == 0 && nv > 0) {
            *reinterpret_cast<float4*>(red + (0 * 4 + q) * 32 + c4) = s1;
            *reinterpret_cast<float4*>(red + (1 * 4 + q) * 32 + c4) = s2;
          }
          named_bar(bar_g, 128);
          if (tid == 32 * T2_NSTAGER) { T2_STAMP(3, ev); ++ev; }
          if (gtid < nvalid && cnt_rows > 0) {      // one thread per column: fp32 sum over the group's warps, widened once
            const float S1 = red[gtid] + red[32 + gtid] + red[64 + gtid] + red[96 + gtid];
            const float S2 = red[128 + gtid] + red[160 + gtid] + red[192 + gtid] + red[224 + gtid];
            const int col = T.n0 + pc0 + gtid;
            if (MODE == T2_FWD) {
              const float bias = ld_opt(G.bias, col, 0.f) + ld_opt(G.bias2, col, 0.f);
              const double y0 = (double)t2_epi_val(ot[gtid], bias, G.e_act, G.e_scale), c = (double)cnt_rows, d1 = (double)S1;
              atomicAdd(G.stats_out + 2 * col, c * y0 + d1);
              atomicAdd(G.stats_out + 2 * col + 1, c * y0 * y0 + 2.0 * y0 * d1 + (double)S2);
            } else {
              atomicAdd(D.dstats + 2 * col, (double)S1);
              atomicAdd(D.dstats + 2 * col + 1, (double)S2);
            }
          }
        }
      }
      if (tid == 32 * T2_NSTAGER) { T2_STAMP(3, ev); ++ev; }
    }
  }
  if (tid == 32 * T2_NSTAGER) T2_STAMP(3, 127);
  if (tid == 0) T2_STAMP(2, 127);
  fence_before_sync();
  __syncthreads();
  if (C > 1) cluster_sync_all();           // no CTA leaves while a peer may still signal its barriers
  if (warp == T2_W_MMA) { __syncwarp(); tmem_dealloc(tmem, T2_TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------------------
// weight / bias gradient: dWeff[n, j] = sum_b dY[b, n] * act(norm(A))[b, j],  db[n] = sum_b dY[b, n]
// accumulator rows = output features n, columns = input features j, contraction = the batch rows of one split.
// TMA brings three raw tiles per k-block (32 batch rows): dz [32][128 n], raw output [32][128 n] (BatchNorm layers),
// input activation [32][NT j], unswizzled.  The stagers form dY^T in registers (lane = output feature) -> TMEM, and
// act(norm(A)) -> hi / lo MN-major swizzled tiles in shared memory (the MMA's column operand).
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
  const int SB = p.sb, SR = p.sr, F = p.flush;
  int t_begin, t_end;
  t2_range(p, p.n_tiles, 1, t_begin, t_end);
  if (tid == 0) T2_STAMP(0, 126);
  if (warp == T2_W_MMA + 1)
    for (int i = lane; i < t_end - t_begin; i += 32) sh.tile[i] = t2_decode_wgrad(p, t_begin + i);
  if (warp == T2_W_TMA && lane < p.n_groups) {
    tma_prefetch_desc(&p.tm0[lane]); tma_prefetch_desc(&p.tm2[lane]);
    if (p.g[lane].Y.norm.mode == SWR_NORM_BATCH) tma_prefetch_desc(&p.tm1[lane]);
  }
  t2_setup(sh, tid, warp, T2_NSTAGER, 1);
  const uint32_t tmem = sh.tmem_base;
  if (tid == 0) T2_STAMP(0, 127);
  if (warp >= T2_NSTAGER + T2_NEPI) {
   reg_dec<T2_REG_ISSUE>();
   if (warp == T2_W_TMA) {
    T2Ring rr; rr.init();
    int ev = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = sh.tile[t - t_begin];
      const FcGroup& G = p.g[T.g];
      const bool two = G.Y.norm.mode == SWR_NORM_BATCH;
      const uint32_t bytes = (two ? 2u : 1u) * T2_RAW_BYTES + 32u * (uint32_t)T.NT * 4u;
      for (int kb = 0; kb < T.nkb; ++kb, rr.next(SR)) {
        mbar_wait(&sh.empty_r[rr.s], rr.ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&sh.full_r[rr.s], bytes);
          const uint32_t rdst = smem_s + (uint32_t)(p.off_r + rr.s * p.stage_r);
          const int b0 = T.b_begin + kb * KBLK;
          tma_load_2d(rdst, &p.tm0[T.g], T.m0, b0, &sh.full_r[rr.s]);
          if (two) tma_load_2d(rdst + T2_RAW_BYTES, &p.tm1[T.g], T.m0, b0, &sh.full_r[rr.s]);
          tma_load_2d(rdst + 2 * T2_RAW_BYTES, &p.tm2[T.g], T.n0, b0, &sh.full_r[rr.s]);
        }
        __syncwarp();
        if (lane == 0) { T2_STAMP(0, ev); } ++ev;
      }
    }
   } else if (warp == T2_W_MMA) {
    T2Ring ra; ra.init();
    uint32_t acc_it = 0;
    int ev = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = sh.tile[t - t_begin];
      const uint32_t b_bytes = (uint32_t)T.NT * 128u;
      const uint32_t idesc = make_idesc_tf32(T2_BM, T.NT, false, true);
      int fpos = 0;
      for (int kb = 0; kb < T.nkb; ++kb, ra.next(SB)) {
        const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        if (fpos == 0) mbar_wait(&sh.acc_empty[buf], aph ^ 1u);
        mbar_wait(&sh.full_a[ra.s], ra.ph);
        fence_after_sync();
        const bool last = (fpos + 1 == F) || (kb == T.nkb - 1);
        if (elect_one()) {
          t2_issue(tmem + buf * T2_ACC_COLS, tmem + T2_A_COL0 + (uint32_t)ra.s * 64u, smem_s + (uint32_t)(ra.s * p.stage_b), b_bytes,
                   true, idesc, fpos == 0);
          mma_commit(&sh.empty_a[ra.s]);
          if (last) mma_commit(&sh.acc_full[buf]);
        }
        __syncwarp();
        if (last) { ++acc_it; fpos = 0; } else ++fpos;
        if (lane == 0) { T2_STAMP(1, ev); } ++ev;
      }
    }
   }
  } else if (warp < T2_NSTAGER) {
    /* 96 at launch: nothing to give back */
    const int q = warp & 3, kh = warp >> 2, stid = tid;
    const int nl = 32 * q + lane;                    // output feature inside the tile = TMEM lane
    const uint32_t ta = tmem + T2_A_COL0 + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * kh);
    float* coef = reinterpret_cast<float*>(smem + p.off_coef);   // [3][NT]: mu, s, b of the input columns of the tile
    T2Ring rr, ra; rr.init(); ra.init();
    int cur_g = -1, cur_n0 = -1, ev = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = sh.tile[t - t_begin];
      if (T.nkb == 0) continue;
      const FcGroup& G = p.g[T.g];
      const int N = G.Y.n, K = G.A.n, NT = T.NT;
      const int lg = (NT == 128) ? 5 : (NT == 64 ? 4 : 3);       // log2(NT / 4): float4 quads per row of the column operand
      const int nit = NT >> 5;                                    // float4 per thread and k-block
      const int qd = stid & ((1 << lg) - 1), c0 = stid >> lg, cstep = (32 * T2_NSTAGER) >> lg;
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
      const int n = T.m0 + nl;
      const bool n_ok = n < N;
      DyCoef dc = {0.f, 0.f, 0.f};
      if (n_ok) dc = dy_coef(G.Y, n, p.inv_count);
      const bool two = (G.Y.norm.mode == SWR_NORM_BATCH);
      const float slope = t2_slope(G.A.act);
      const bool sig = G.A.act == SWR_ACT_SIGMOID;
      const int rows = T.b_end - T.b_begin;
      // this thread's input-column coefficients do not change over the k-blocks of a tile
      const float4 cmu = t2_ld4s(coef + 4 * qd), csc = t2_ld4s(coef + NT + 4 * qd), cbb = t2_ld4s(coef + 2 * NT + 4 * qd);
      const uint32_t b_bytes = (uint32_t)NT * 128u;
      float rowsum = 0.f;
      for (int kb = 0; kb < T.nkb; ++kb, rr.next(SR), ra.next(SB)) {
        mbar_wait(&sh.full_r[rr.s], rr.ph);
        if (tid == 0) { T2_STAMP(2, ev); ++ev; }
        const float* rt = reinterpret_cast<const float*>(smem + p.off_r + rr.s * p.stage_r);
        // ---- dY^T: 16 batch rows of output feature n ----
        float v[16];
        const int left = n_ok ? rows - (kb * KBLK + 16 * kh) : 0;   // c2 != 0: contraction padding must stay exactly zero
        const float* dzs = rt + (16 * kh) * T2_BM + nl;
        if (two) {
          const float* rws = dzs + T2_RAW_BYTES / 4;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (i < left) ? fmaf(dc.c0, dzs[i * T2_BM], fmaf(dc.c1, rws[i * T2_BM], dc.c2)) : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (i < left) ? fmaf(dc.c0, dzs[i * T2_BM], dc.c2) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) rowsum += v[i];
        // ---- act(norm(A)): float4 (batch row c, feature quad qd) ----
        const float* as = rt + 2 * (T2_RAW_BYTES / 4);
        float4 xb[4];
#pragma unroll
        for (int it = 0; it < 4; ++it)
          if (it < nit) xb[it] = t2_ld4s(as + (c0 + it * cstep) * NT + 4 * qd);
        mbar_wait(&sh.empty_a[ra.s], ra.ph ^ 1u);
        fence_after_sync();
        t2_store_split16(ta + (uint32_t)ra.s * 64u, v);
        const uint32_t bh = smem_s + (uint32_t)(ra.s * p.stage_b), bl = bh + b_bytes;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          if (it < nit) {
            float4 x = xb[it];
            if (!plainA) {
              x.x = fmaf(x.x - cmu.x, csc.x, cbb.x); x.y = fmaf(x.y - cmu.y, csc.y, cbb.y);
              x.z = fmaf(x.z - cmu.z, csc.z, cbb.z); x.w = fmaf(x.w - cmu.w, csc.w, cbb.w);
              if (sig) { x.x = 1.f / (1.f + expf(-x.x)); x.y = 1.f / (1.f + expf(-x.y)); x.z = 1.f / (1.f + expf(-x.z)); x.w = 1.f / (1.f + expf(-x.w)); }
              else { x.x = fmaxf(x.x, slope * x.x); x.y = fmaxf(x.y, slope * x.y); x.z = fmaxf(x.z, slope * x.z); x.w = fmaxf(x.w, slope * x.w); }
            }
            store_split(bh, bl, mnmajor_off(qd, c0 + it * cstep), x);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty_r[rr.s]);
        fence_proxy_async();      // shared-memory stores -> visible to the MMA unit
        tmem_st_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.full_a[ra.s]);
        if (tid == 0) { T2_STAMP(2, ev); ++ev; }
      }
      if (T.n0 == 0 && n_ok) {    // bias gradient: the j-tile 0 CTAs carry it (two threads per output feature)
        if (G.dbias) atomicAdd(G.dbias + n, rowsum);
        if (G.dbias2) atomicAdd(G.dbias2 + n, rowsum);
      }
    }
  } else {
    reg_inc<T2W_REG_EPI>();
    // two independent epilogue groups, see fc_tc2_kernel
    const int e = warp - T2_NSTAGER, q = warp & 3, ch = e >> 2;
    const int arow = 32 * q + lane;
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const int bar_g = T2_BAR_EGRP + ch;
    const float* ot = reinterpret_cast<const float*>(smem + p.off_ot) + ch * T2_OT_GROUP;
    const uint32_t ot_s = smem_s + (uint32_t)p.off_ot + (uint32_t)(ch * T2_OT_GROUP) * 4u;
    uint32_t acc_it = 0;
    int ev = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const T2Tile T = sh.tile[t - t_begin];
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
        if (tid == 32 * T2_NSTAGER) { T2_STAMP(3, ev); ++ev; }
        t2_drain_add(tmem + buf * T2_ACC_COLS, q, ch, T.NT, acc);
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.acc_empty[buf]);
      }
      const bool kn = (G.w_layout == SWR_W_KN);
      const int mvalid = min(N - T.m0, T2_BM);
      const int row_base = 32 * q + rsub;
      const int rows_here = min((max(mvalid - row_base, 0) + 3) >> 2, 8);
      for (int sp = 0; sp < 2; ++sp) {
        const int pc0 = 64 * ch + 32 * sp;
        if (pc0 >= T.NT) break;
        const int pc = min(32, T.NT - pc0);
        const int nvalid = min(max(K - (T.n0 + pc0), 0), pc);
        named_bar(bar_g, 128);
        t2_acc_to_smem(ot_s, arow, sp, pc, acc);
        named_bar(bar_g, 128);
        const int j = T.n0 + pc0 + c4, nv = nvalid - c4;
        if (!kn) {                                    // dW[n, j]: 8 lanes cover 32 consecutive input features of one output feature
          const bool vec = !G.W2 && G.dW && (nv >= 4) && (G.ldw % 4 == 0) && is_al16(G.dW) && (j % 4 == 0);
          if (nv > 0) {
            const uint32_t my_ot = ot_s + (uint32_t)(row_base * T2_OT_LD + c4) * 4u;
            const int64_t o0 = (int64_t)(T.m0 + row_base) * G.ldw + j, ostep = 4 * (int64_t)G.ldw;
            if (vec) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < rows_here) t2_red_add_v4(G.dW + o0 + i * ostep, lds128(my_ot + (uint32_t)(4 * i * T2_OT_LD) * 4u));
            } else {
#pragma unroll 1
              for (int i = 0; i < rows_here; ++i) {
                const float4 a4 = lds128(my_ot + (uint32_t)(4 * i * T2_OT_LD) * 4u);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const int64_t o = o0 + i * ostep;
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
        } else {                                      // dW[j, n]: lanes run over output features n
          for (int col = q; col < nvalid; col += 4) {
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
      if (tid == 32 * T2_NSTAGER) { T2_STAMP(3, ev); ++ev; }
    }
  }
  if (tid == 32 * T2_NSTAGER) T2_STAMP(3, 127);
  if (tid == 0) T2_STAMP(2, 127);
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
// weight image [2][rows][pitch] fp32 -> 3-D map {pitch, rows, 2}, box {32, box_rows, 2}, 128-byte swizzle; rows outside
// the image read as zeros
static int t2_make_map(CUtensorMap* tm, const float* img, int rows, int pitch, int box_rows, int box_planes = 2) {
  t2_encode_fn enc = t2_encoder();
  if (!enc) { set_error("fc_tc2: cuTensorMapEncodeTiled is not available"); return SWR_ERR_UNSUPPORTED; }
  const cuuint64_t gdim[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, 2};
  const cuuint64_t gstr[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)rows * pitch * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
  const cuuint32_t est[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("fc_tc2: cuTensorMapEncodeTiled failed (%d) rows=%d pitch=%d box=%d", (int)r, rows, pitch, box_rows); return SWR_ERR_CUDA; }
  return SWR_OK;
}
// activation [rows][ld] fp32 with `cols` valid columns -> 2-D map {cols, rows}, box {box_cols, box_rows}; elements
// outside [cols) x [rows) read as zeros.  swizzled: 128-byte swizzle (box_cols == 32), else plain row-major tiles.
static int t2_make_act_map(CUtensorMap* tm, const float* base, int rows, int cols, int ld, int box_cols, int box_rows, bool swizzled) {
  t2_encode_fn enc = t2_encoder();
  if (!enc) { set_error("fc_tc2: cuTensorMapEncodeTiled is not available"); return SWR_ERR_UNSUPPORTED; }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t est[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzled ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("fc_tc2: cuTensorMapEncodeTiled (activation) failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld); return SWR_ERR_CUDA; }
  return SWR_OK;
}

static int t2_pick_nt(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }
static int t2_flush_for(int nt) {
  static int forced = [] { const char* e = getenv("SWR_TC_FLUSH"); return e ? atoi(e) : 0; }();
  if (forced > 0) return forced;
  return 256 / nt;                        // NT = 128: every 2 k-blocks (192 products per truncated chain), the time one drain takes
}

template <class K>
static int t2_set_smem(K kernel, size_t bytes) {
  constexpr int kMaxDyn = 227 * 1024 - 2048;      // the opt-in maximum covers static (barriers, decoded tiles: ~1.6 KB) + dynamic shared memory
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

// shared-memory plan: ring B [sb][stage_b] | ring R [sr][stage_r] | coef | ccs | ot | red
// two_groups: the forward / data-gradient kernels, whose two stager groups take alternate k-blocks.  Their raw ring must
// have an EVEN number of stages, so that a stage always belongs to the same group: with three stages a stage alternates
// between the groups, each group revisits it only every second pass, and its parity wait is satisfied by the pass in
// between -- a group running two k-blocks ahead of the other then staged the other group's (older) tile.  That was a
// latent race (about one forward in forty at B = 16384 with 10+ tiles per CTA, tools/stress_forward.py).
static int t2_plan_smem(Tc2Params& p, size_t stage_b, size_t stage_r, size_t coef_bytes, size_t ccs_bytes, size_t* smem_bytes,
                        bool two_groups = true) {
  const size_t fixed = ((coef_bytes + 15) & ~(size_t)15) + ((ccs_bytes + 15) & ~(size_t)15) + T2_OT_BYTES + T2_RED_BYTES;
  // 227 KB opt-in maximum minus static shared memory and alignment slack (SWR_TC_SMEM_KB: experiments with the
  // 196 KB carve-out, which leaves 32 KB of L1, showed no gain)
  static const size_t budget = [] { const char* e = getenv("SWR_TC_SMEM_KB"); return (size_t)(e ? atoi(e) : 224) * 1024 - 1024; }();
  static const int pref[][2] = {{4, 4}, {4, 3}, {3, 4}, {3, 3}, {4, 2}, {3, 2}, {2, 4}, {2, 3}, {2, 2}};
  int sb = 0, sr = 0;
  for (auto& c : pref) {
    if (two_groups && (c[1] & 1)) continue;
    if (fixed + c[0] * stage_b + c[1] * stage_r <= budget) { sb = c[0]; sr = c[1]; break; }
  }
  if (!sb) { set_error("fc_tc2: %zu bytes of tables do not fit beside the stage rings", fixed); return SWR_ERR_UNSUPPORTED; }
  p.sb = sb; p.sr = sr; p.stage_b = (int)stage_b; p.stage_r = (int)stage_r;
  size_t off = (size_t)sb * stage_b;
  p.off_r = (int)off; off += (size_t)sr * stage_r;
  p.off_coef = (int)off; off += (coef_bytes + 15) & ~(size_t)15;
  p.off_ccs = (int)off; off += (ccs_bytes + 15) & ~(size_t)15;
  p.off_ot = (int)off; off += T2_OT_BYTES;
  p.off_red = (int)off; off += T2_RED_BYTES;
  *smem_bytes = off + 1024;
  return SWR_OK;
}

// CTAs per cluster for the forward / data-gradient kernels: every weight tile is fetched once per cluster (each CTA
// multicasts 1 / C of it).  Needs 128-wide tiles throughout the launch (slices of whole swizzle atoms) and at least C row
// tiles.  Measured on cfg2 (profiles/r02_fc_tc2_notes.md): no gain at C = 2 or 4 -- the kernels are bound by shared-memory
// bandwidth (TMA writes + three operand reads per k-step + the stagers' reads), not by L2 -> SM traffic -- so the default
// is 1; SWR_TC_CLUSTER=2|4 enables it.
static int t2_pick_cluster(bool all128, int mtiles) {
  static int want = [] { const char* e = getenv("SWR_TC_CLUSTER"); const int v = e ? atoi(e) : 1; return (v == 1 || v == 2 || v == 4) ? v : 1; }();
  int c = all128 ? want : 1;
  while (c > 1 && mtiles < c) c >>= 1;
  return c;
}
// CTAs (clusters) of the persistent grid: one per SM, never more than tiles
template <class K>
static int t2_grid(K kernel, int n_tiles, int cluster, size_t smem) {
  const int sms = t2_num_sms();
  int max_clusters = sms / cluster;
  if (cluster > 1) {
    // clusters that can be resident at once (GPC boundaries strand a few SMs at cluster size 4)
    static thread_local int cached[5] = {0, 0, 0, 0, 0};
    if (!cached[cluster]) {
      cudaLaunchConfig_t cfg{};
      cfg.blockDim = dim3(T2_THREADS, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cfg.gridDim = dim3(sms / cluster * cluster, 1, 1);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = max_clusters; }
      cached[cluster] = n;
    }
    max_clusters = min(max_clusters, cached[cluster]);
  }
  return max(1, min(min(n_tiles, max_clusters), T2_MAX_CTAS));
}
// every CTA must be able to hold its decoded tiles (T2_TILE_CAP): a balanced partition that gives one CTA more (many
// cheap tiles) is dropped for the equal split; launches whose equal split does not fit are refused (FFMA serves them)
static int t2_check_ranges(Tc2Params& p, int n_tiles, int n_ctas) {
  if (p.balanced) {
    int worst = 0;
    for (int b = 0; b < n_ctas; ++b) worst = max(worst, (int)p.cta_tile[b + 1] - (int)p.cta_tile[b]);
    if (worst <= T2_TILE_CAP) return SWR_OK;
    p.balanced = 0;
  }
  return (n_tiles + n_ctas - 1) / n_ctas <= T2_TILE_CAP ? SWR_OK : SWR_ERR_UNSUPPORTED;
}

// Contiguous tile ranges of (almost) equal cost: the smallest bottleneck cost `cap` for which a greedy walk needs at most
// `nb` ranges (binary search), then that walk.  cost[t] > 0.
static void t2_balance(Tc2Params& p, const std::vector<int>& cost, int nb) {
  const int n = (int)cost.size();
  p.balanced = 0;
  if (n > 65535 || nb > T2_MAX_CTAS || nb <= 0) return;
  long long lo = 0, hi = 0;
  for (int c : cost) { lo = std::max<long long>(lo, c); hi += c; }
  auto ranges_needed = [&](long long cap) {
    int r = 1; long long cur = 0;
    for (int c : cost) { if (cur + c > cap) { ++r; cur = 0; } cur += c; }
    return r;
  };
  while (lo < hi) {
    const long long mid = (lo + hi) / 2;
    if (ranges_needed(mid) <= nb) hi = mid; else lo = mid + 1;
  }
  // tiles of mixed cost (gate tiles beside expert tiles): deal them out longest first to the least loaded CTA and keep
  // that assignment when its bottleneck beats the contiguous one by a tenth (contiguous ranges change group less often,
  // and every change rebuilds a coefficient table)
  if (n <= T2_MAX_PERM && nb > 1) {
    std::vector<int> order(n);
    for (int t = 0; t < n; ++t) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b_) { return cost[a] > cost[b_]; });
    std::vector<long long> load(nb, 0);
    std::vector<std::vector<int>> mine(nb);
    for (int t : order) {
      int best = 0;
      for (int c = 1; c < nb; ++c) if (load[c] < load[best]) best = c;
      load[best] += cost[t]; mine[best].push_back(t);
    }
    const long long lpt = *std::max_element(load.begin(), load.end());
    if (lpt * 10 < lo * 9) {
      int pos = 0;
      for (int c = 0; c < nb; ++c) {
        p.cta_tile[c] = (unsigned short)pos;
        for (int t : mine[c]) p.perm[pos++] = (unsigned short)t;
      }
      p.cta_tile[nb] = (unsigned short)pos;
      p.balanced = 2;
      return;
    }
  }
  int b = 0; long long cur = 0;
  p.cta_tile[0] = 0;
  for (int t = 0; t < n; ++t) {
    // start a new range when the cap would be exceeded, or when the tiles left are just enough to give every remaining
    // CTA one (no CTA stays idle while another holds two tiles)
    const bool must = (n - t) <= (nb - 1 - b) && cur > 0;
    if ((cur + cost[t] > lo || must) && b + 1 < nb) { p.cta_tile[++b] = (unsigned short)t; cur = 0; }
    cur += cost[t];
  }
  while (b < nb) p.cta_tile[++b] = (unsigned short)n;
  p.balanced = 1;
}
// modelled cycles of one tile (measured on cfg2, profiles/r02_fc_tc2_notes.md): a k-block costs what the slower of the
// stagers (~600) and the MMAs (12 x NT / 128 x ~105) takes, the final epilogue ~1500 + 12 NT
static int t2_tile_cost(int nkb, int nt) { return nkb * std::max(600, 10 * nt) + 1500 + 12 * nt; }

template <class K>
static int t2_launch(K kernel, int n_ctas, int cluster, size_t smem, const Tc2Params& p, cudaStream_t st, const char* name) {
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(T2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (cluster > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  cfg.gridDim = dim3(n_ctas * cluster, 1, 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
  if (e != cudaSuccess) { set_error("launch of %s failed: %s", name, cudaGetErrorString(e)); return SWR_ERR_CUDA; }
  count_launch();
  return SWR_OK;
}

// SWR_FC_SIMT / SWR_FC_TC / SWR_FC_AUTO (include/swr_b200.h); preset by env SWR_FC_TC, changed by swr_set_fc_mode()
static std::atomic<int> g_tc_mode{-1};
int fc_mode_get() {
  int m = g_tc_mode.load();
  if (m < 0) {
    const char* e = getenv("SWR_FC_TC");
    m = e ? atoi(e) : SWR_FC_AUTO;
    if (m < 0 || m > 2) m = SWR_FC_AUTO;
    g_tc_mode.store(m);
  }
  return m;
}
int fc_mode_set(int mode) {
  const int prev = fc_mode_get();
  if (mode >= 0 && mode <= 2) g_tc_mode.store(mode);
  return prev;
}
static int tc_mode() { return fc_mode_get(); }
static int64_t tc_min_macs() {
  static int64_t v = -1;
  if (v < 0) { const char* e = getenv("SWR_FC_TC_MIN_MACS"); v = e ? atoll(e) : (int64_t)1 << 24; }
  return v;
}

bool fc_tc_wanted(const FcGroup* groups, int n_groups, int64_t B) {
  const int mode = tc_mode();
  if (mode == 0 || B < 64) return false;
  int64_t macs = 0;
  int kmax = 0;
  for (int g = 0; g < n_groups; ++g) {
    macs += (int64_t)groups[g].A.n * groups[g].Y.n;
    kmax = max(kmax, max(groups[g].A.n, groups[g].Y.n));
  }
  if (kmax > 4096) return false;   // coefficient tables live in shared memory
  if (mode == 1) return true;
  return macs * B >= tc_min_macs();
}


static inline bool is_al16_host(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool fc_tc2_usable(const FcGroup* groups, int n_groups, int pass) {
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

// development aid (SWR_TC_DEBUG): clock64 stamps of CTA 0 per warp role, printed after the launch
struct T2Debug {
  bool on;
  static long long* buffer() { static long long* d = nullptr; if (!d) cudaMalloc(&d, 512 * sizeof(long long)); return d; }
  T2Debug(Tc2Params& p, cudaStream_t st) {
    static const bool debug = getenv("SWR_TC_DEBUG") != nullptr;
    on = debug;
    if (on) { cudaMemsetAsync(buffer(), 0, 512 * sizeof(long long), st); p.dbg = buffer(); }
  }
  void report(const char* what, const Tc2Params& p, int tiles, cudaStream_t st) const {
    if (!on) return;
    long long h[512];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, buffer(), sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[126];
    int t_per = tiles / 148 + (tiles % 148 ? 1 : 0);
    fprintf(stderr, "[tc2 %s dbg] tiles=%d (cta0: %d) groups=%d K0=%d N0=%d cluster=%d nt0=%d sb=%d sr=%d flush=%d splits=%d setup=%lld\n", what, tiles, t_per,
            p.n_groups, p.g[0].A.n, p.g[0].Y.n, p.cluster, p.nt[0], p.sb, p.sr, p.flush, p.splits, h[127] - t0);
    const char* names[4] = {"tma", "mma", "stg", "epi"};
    for (int r = 0; r < 4; ++r) {
      fprintf(stderr, "  %s:", names[r]);
      for (int i = 0; i < 126 && h[r * 128 + i]; ++i) fprintf(stderr, " %lld", h[r * 128 + i] - t0);
      if (r == 2 || r == 3) fprintf(stderr, " | end %lld", h[r * 128 + 127] - t0);
      fprintf(stderr, "\n");
    }
  }
};

int launch_fc_tc2_fwd(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  Tc2Params p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  const int mtiles = ceil_div(B, T2_BM);
  bool all128 = true;
  for (int g = 0; g < n_groups; ++g) all128 = all128 && t2_pick_nt(groups[g].Y.n) == 128;
  const int C = t2_pick_cluster(all128, mtiles);
  p.cluster = C;
  int tiles = 0, kmax = 0, nt_max = 0, flush = 1 << 30;
  for (int g = 0; g < n_groups; ++g) {
    p.g[g] = groups[g];
    const int N = groups[g].Y.n, K = groups[g].A.n;
    p.nt[g] = t2_pick_nt(N);
    p.tile_start[g] = tiles;
    tiles += ceil_div(mtiles, C) * ceil_div(N, p.nt[g]);
    kmax = max(kmax, K); nt_max = max(nt_max, p.nt[g]);
    flush = min(flush, t2_flush_for(p.nt[g]));
    int rc = C == 1 ? t2_make_map(&p.tm0[g], groups[g].img_f, N, t2_round_up(K, 32), p.nt[g])
                    : t2_make_map(&p.tm0[g], groups[g].img_f, N, t2_round_up(K, 32), p.nt[g] / C, 1);
    if (rc) return rc;
    rc = t2_make_act_map(&p.tm1[g], groups[g].A.raw, (int)B, K, groups[g].A.ld, 32, T2_BM, true);
    if (rc) return rc;
  }
  p.tile_start[n_groups] = tiles; p.n_tiles = tiles; p.flush = flush;
  size_t smem = 0;
  int rc = t2_plan_smem(p, 2 * (size_t)nt_max * 128, T2_RAW_BYTES, 3 * sizeof(float) * (size_t)t2_round_up(kmax, KBLK), 0, &smem);
  if (rc) return rc;
  rc = t2_set_smem(fc_tc2_kernel<T2_FWD>, smem);
  if (rc) return rc;
  const int nb = t2_grid(fc_tc2_kernel<T2_FWD>, tiles, C, smem);
  {
    std::vector<int> cost;
    cost.reserve(tiles);
    for (int g = 0; g < n_groups; ++g)
      cost.insert(cost.end(), p.tile_start[g + 1] - p.tile_start[g], t2_tile_cost(ceil_div(groups[g].A.n, KBLK), p.nt[g]));
    t2_balance(p, cost, nb);
  }
  if (t2_check_ranges(p, tiles, nb)) return SWR_ERR_UNSUPPORTED;
  T2Debug dbg(p, st);
  rc = t2_launch(fc_tc2_kernel<T2_FWD>, nb, C, smem, p, st, "fc_tc2_kernel<fwd>");
  if (rc) return rc;
  dbg.report("fwd", p, tiles, st);
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
  bool all128 = true;
  for (int d = 0; d < n_dst; ++d) all128 = all128 && t2_pick_nt(groups[dst_group[d]].A.n) == 128;
  const int C = t2_pick_cluster(all128, mtiles);
  p.cluster = C;
  int tiles = 0, nkb_max = 0, nt_max = 0, flush = 1 << 30;
  for (int d = 0; d < n_dst; ++d) {
    const int kd = groups[dst_group[d]].A.n;
    p.nt[d] = t2_pick_nt(kd);
    p.dst_tile[d] = tiles;
    tiles += ceil_div(mtiles, C) * ceil_div(kd, p.nt[d]);
    nkb_max = max(nkb_max, p.tile_start[dst_group[d + 1]] - p.tile_start[dst_group[d]]);
    nt_max = max(nt_max, p.nt[d]);
    flush = min(flush, t2_flush_for(p.nt[d]));
    for (int g = dst_group[d]; g < dst_group[d + 1]; ++g) {
      const FcGroup& G = groups[g];
      // the image holds every input feature of the layer (k_full rows); the destination may use only its first A.n
      int rc = C == 1 ? t2_make_map(&p.tm0[g], G.img_d, G.k_full, t2_round_up(G.Y.n, 32), p.nt[d])
                      : t2_make_map(&p.tm0[g], G.img_d, G.k_full, t2_round_up(G.Y.n, 32), p.nt[d] / C, 1);
      if (rc) return rc;
      rc = t2_make_act_map(&p.tm1[g], G.Y.dz, (int)B, G.Y.n, G.Y.ld, 32, T2_BM, true);
      if (rc) return rc;
      if (G.Y.norm.mode == SWR_NORM_BATCH) {
        rc = t2_make_act_map(&p.tm2[g], G.Y.raw, (int)B, G.Y.n, G.Y.ld, 32, T2_BM, true);
        if (rc) return rc;
      }
    }
  }
  p.dst_tile[n_dst] = tiles; p.n_tiles = tiles; p.flush = flush;
  size_t smem = 0;
  int rc = t2_plan_smem(p, 2 * (size_t)nt_max * 128, 2 * (size_t)T2_RAW_BYTES, 3 * sizeof(float) * (size_t)nkb_max * KBLK,
                        4 * sizeof(float) * (size_t)nt_max, &smem);
  if (rc) return rc;
  rc = t2_set_smem(fc_tc2_kernel<T2_DGRAD>, smem);
  if (rc) return rc;
  const int nb = t2_grid(fc_tc2_kernel<T2_DGRAD>, tiles, C, smem);
  {
    std::vector<int> cost;
    cost.reserve(tiles);
    for (int d = 0; d < n_dst; ++d)
      cost.insert(cost.end(), p.dst_tile[d + 1] - p.dst_tile[d], t2_tile_cost(p.tile_start[dst_group[d + 1]] - p.tile_start[dst_group[d]], p.nt[d]));
    t2_balance(p, cost, nb);
  }
  if (t2_check_ranges(p, tiles, nb)) return SWR_ERR_UNSUPPORTED;
  T2Debug dbg(p, st);
  rc = t2_launch(fc_tc2_kernel<T2_DGRAD>, nb, C, smem, p, st, "fc_tc2_kernel<dgrad>");
  if (rc) return rc;
  dbg.report("dgrad", p, tiles, st);
  return SWR_OK;
}

int launch_fc_tc2_wgrad(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  Tc2Params p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  int base = 0, nt_max = 0, flush = 1 << 30;
  for (int g = 0; g < n_groups; ++g) {
    const FcGroup& G = groups[g];
    p.g[g] = G;
    p.nt[g] = max(32, t2_pick_nt(G.A.n));      // the MN-major column operand is staged in 32-wide groups
    base += ceil_div(G.Y.n, T2_BM) * ceil_div(G.A.n, p.nt[g]);
    nt_max = max(nt_max, p.nt[g]);
    flush = min(flush, t2_flush_for(p.nt[g]));
    int rc = t2_make_act_map(&p.tm0[g], G.Y.dz, (int)B, G.Y.n, G.Y.ld, T2_BM, 32, false);
    if (rc) return rc;
    if (G.Y.norm.mode == SWR_NORM_BATCH) {
      rc = t2_make_act_map(&p.tm1[g], G.Y.raw, (int)B, G.Y.n, G.Y.ld, T2_BM, 32, false);
      if (rc) return rc;
    }
    rc = t2_make_act_map(&p.tm2[g], G.A.raw, (int)B, G.A.n, G.A.ld, p.nt[g], 32, false);
    if (rc) return rc;
  }
  // Split the batch so that the modelled duration is smallest: a CTA walks ceil(tiles / SMs) tiles, a tile costs its
  // k-blocks (the stagers bound the loop: ~2000 cycles each) plus ~8000 cycles of fixed work (coefficient tables,
  // pipeline fill, drain and the weight-gradient atomics); measured on cfg2 (profiles/r02_fc_tc2_notes.md).
  // Keep >= 4 k-blocks (128 rows) per split.
  const int sms = t2_num_sms();
  int splits = 1, rows = t2_round_up((int)B, KBLK);
  {
    long long best = -1;
    const int smax = max(1, min(ceil_div(B, 4 * KBLK), 64));
    for (int s_ = 1; s_ <= smax; ++s_) {
      const int r_ = t2_round_up(ceil_div(B, s_), KBLK), n_ = ceil_div(B, r_);
      const long long per_cta = ceil_div(base * n_, sms);
      const long long t_ = per_cta * ((long long)(r_ / KBLK) * 2000 + 8000);
      if (best < 0 || t_ < best) { best = t_; splits = n_; rows = r_; }
    }
  }
  p.splits = splits; p.rows_per_split = rows;
  int tiles = 0;
  for (int g = 0; g < n_groups; ++g) {
    p.tile_start[g] = tiles;
    tiles += ceil_div(groups[g].Y.n, T2_BM) * ceil_div(groups[g].A.n, p.nt[g]) * splits;
  }
  p.tile_start[n_groups] = tiles; p.n_tiles = tiles; p.flush = flush;
  size_t smem = 0;
  int rc = t2_plan_smem(p, 2 * (size_t)nt_max * 128, 2 * (size_t)T2_RAW_BYTES + 32 * (size_t)nt_max * 4, 3 * sizeof(float) * (size_t)nt_max, 0, &smem,
                        false);      // all eight stager warps take every k-block
  if (rc) return rc;
  rc = t2_set_smem(fc_tc2_wgrad_kernel, smem);
  if (rc) return rc;
  p.cluster = 1;
  const int nb = t2_grid(fc_tc2_wgrad_kernel, tiles, 1, smem);
  if (t2_check_ranges(p, tiles, nb)) return SWR_ERR_UNSUPPORTED;
  T2Debug dbg(p, st);
  rc = t2_launch(fc_tc2_wgrad_kernel, nb, 1, smem, p, st, "fc_tc2_wgrad_kernel");
  if (rc) return rc;
  dbg.report("wgrad", p, tiles, st);
  return SWR_OK;
}

}  // namespace swr
