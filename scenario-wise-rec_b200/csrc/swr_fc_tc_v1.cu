// swr_fc_tc.cu -- grouped fully-connected kernels on the sm_100a tensor cores (tcgen05 + TMEM).
//
// Same contract as the SIMT kernels in swr_fc.cu (fc_fwd / fc_dgrad / fc_wgrad over FcGroup lists,
// reference: basic/layers.py:253-258 Linear -> BatchNorm1d -> act, star.py:103-110, ppnet.py:21-29,
// hamur.py adapters, m3oe.py:45-68), but the contraction runs as tcgen05.mma.kind::tf32 with the
// accumulator tile (128 rows x up to 256 fp32 columns) in tensor memory.
//
// fp32 parity: tcgen05 has no fp32 MMA, and a single TF32 pass misses the 1e-4 bar, so every operand
// element x is split while it is staged into x = hi + lo (both TF32-representable) and each k-step issues
// three MMAs  lo*hi + hi*lo + hi*hi  into the same accumulator ("3xTF32", error ~2^-21 relative).
//
// Operands cannot come through TMA because they are *computed* on the way in (DESIGN.md "Lazy
// activations"): the CTA's 512 threads load raw values with 16-byte read-only loads (register-prefetched
// one k-block ahead), apply the lazy BatchNorm/activation (forward), the BatchNorm-backward affine map
// (dgrad/wgrad) or the STAR weight product, and split.  Where the results go:
//   forward, dgrad (TS mode)  the 128-row operand (activations / dY) goes registers -> TMEM with tcgen05.st
//                             and is the MMA's A operand in tensor memory; the weight tile goes to shared
//                             memory in the 128-byte swizzled layout the MMA unit reads (swr_tc.cuh)
//   wgrad (SS mode)           both operands (dY^T and the activations, batch-strided) go to shared memory
// Lane 0 of warp (k-block mod 16) issues that block's MMAs once every thread has arrived on the stage's
// mbarrier; tcgen05.commit arrives on a second mbarrier to hand the stage back to the stagers.  The epilogue
// pulls the tile out of TMEM with tcgen05.ld (one row per thread), transposes it through shared memory, and
// runs the same fused tails as the SIMT kernels (bias / GateNU activation / fp64 column moments; act' +
// BatchNorm stage-1 sums; weight-gradient reductions) with 16-byte coalesced global accesses.
#include "swr_common.cuh"
#include "swr_launch.h"
#include "swr_tc.cuh"
#include <atomic>
#include <cstdio>
#include <cstdlib>

namespace swr {
using namespace tc;

constexpr int TC_BM = 128;        // accumulator rows (TMEM lanes) per CTA
constexpr int TC_NP = 512;        // threads per CTA: 16 warps stage operands and run the epilogue; lane 0 of warp (kb % 16) also
constexpr int TC_NT = TC_NP;      // issues the MMAs of k-block kb (a dedicated extra warp would cap the CTA's registers per thread)
constexpr int TC_WARPS = TC_NP / 32;
constexpr int TC_RPI = TC_NP / 8; // rows one pass of a K-major slice covers (8 threads per 128-byte row)
constexpr int TC_MAX_STAGES = 4;
constexpr uint32_t TC_A_BYTES = TC_BM * 128;   // one A tile: 128 rows (or 4 M-groups) x 32 fp32

struct TcParams {
  FcGroup g[kMaxGroups];
  int tile_start[kMaxGroups + 1];  // fwd / wgrad: first CTA of group g;  dgrad: first k-block of group g
  int nt[kMaxGroups];              // accumulator columns per CTA (multiple of 16, <= 256): per group (fwd, wgrad) / per destination (dgrad)
  int n_groups;
  int B;
  float inv_count;
  int stages;
  int epi_off;                     // byte offset (from the aligned shared-memory base) of tables that must survive the epilogue's reuse of the stages
  int splits;                      // wgrad: batch splits
  int rows_per_split;
  long long* dbg;                  // development aid: per-phase cycle counts of CTA 0 (null in production)
  int n_dst;                       // dgrad
  int dst_group[kMaxGroups + 1];
  int dst_tile[kMaxGroups + 1];
  unsigned dst_atomic;             // bit d: destination entry d is one of several partial fan-ins of the same activation: add atomically
};

struct TcShared {
  uint64_t bar_full[TC_MAX_STAGES];   // every thread arrives: the stage holds a complete k-block
  uint64_t bar_free[TC_MAX_STAGES];   // tcgen05.commit arrives: the MMAs that read the stage are done
  uint64_t bar_done;                  // the whole accumulation is done
  uint32_t tmem_base;
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
__device__ __forceinline__ uint32_t tmem_cols(int nt) { uint32_t c = 32; while ((int)c < nt) c <<= 1; return c; }

// barriers + TMEM; every thread calls it, ends with a CTA barrier
__device__ __forceinline__ uint32_t tc_setup(TcShared& sh, int stages, uint32_t tmem_ncols, int tid) {
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&sh.bar_full[s], TC_NP); mbar_init(&sh.bar_free[s], 1); }
    mbar_init(&sh.bar_done, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&sh.tmem_base, tmem_ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  return sh.tmem_base;
}

// one pipeline stage worth of MMAs (4 k-steps x 3 split products), issued by one thread.  Descriptors differ between
// k-steps only in the start address (low word): +32 B per step in a K-major tile, +1024 B in an MN-major one.
__device__ __forceinline__ void tc_issue(uint32_t tmem, uint32_t stage_saddr, uint32_t b_bytes, bool a_mn, bool b_mn, uint32_t idesc, bool first) {
  const uint32_t ah = stage_saddr, al = ah + TC_A_BYTES, bh = al + TC_A_BYTES, bl = bh + b_bytes;
  const uint64_t dah0 = a_mn ? mnmajor_desc(ah, 0) : kmajor_desc(ah, 0);
  const uint64_t dbh0 = b_mn ? mnmajor_desc(bh, 0) : kmajor_desc(bh, 0);
  const uint64_t a_lo_d = (uint64_t)(TC_A_BYTES >> 4), b_lo_d = (uint64_t)(b_bytes >> 4);   // hi tile -> lo tile
  const uint64_t a_step = a_mn ? (1024u >> 4) : ((UMMA_K * 4) >> 4), b_step = b_mn ? (1024u >> 4) : ((UMMA_K * 4) >> 4);
#pragma unroll
  for (int ks = 0; ks < KBLK / UMMA_K; ++ks) {
    const uint64_t dah = dah0 + ks * a_step, dbh = dbh0 + ks * b_step;
    mma_tf32(tmem, dah + a_lo_d, dbh, idesc, (first && ks == 0) ? 0u : 1u);
    mma_tf32(tmem, dah, dbh + b_lo_d, idesc, 1u);
    mma_tf32(tmem, dah, dbh, idesc, 1u);
  }
}

// The software pipeline over k-blocks: stage index and round parity advance together, no divisions in the loop.
struct TcPipe {
  int s;            // stage of the current k-block
  uint32_t par;     // parity of its round (kb / S) & 1
  int kb;
  __device__ __forceinline__ void init() { s = 0; par = 0; kb = 0; }
  __device__ __forceinline__ void advance(int S) { ++kb; if (++s == S) { s = 0; par ^= 1u; } }
};
// wait until the MMAs that last read the current stage are done
__device__ __forceinline__ void tc_acquire(TcShared& sh, const TcPipe& pp, int S) {
  if (pp.kb >= S) { mbar_wait(&sh.bar_free[pp.s], pp.par ^ 1u); fence_after_sync(); }
}
// my part of the k-block is in shared memory; lane 0 of warp (kb % 16) then waits for everybody and issues the MMAs
__device__ __forceinline__ void tc_publish_issue(TcShared& sh, const TcPipe& pp, uint32_t tmem, uint32_t stage_saddr, uint32_t b_bytes, int nkb,
                                                 bool a_mn, bool b_mn, int nt, int tid) {
  fence_proxy_async();
  mbar_arrive(&sh.bar_full[pp.s]);
  if (tid == ((pp.kb & (TC_WARPS - 1)) << 5)) {
    mbar_wait(&sh.bar_full[pp.s], pp.par);
    fence_after_sync();
    tc_issue(tmem, stage_saddr, b_bytes, a_mn, b_mn, make_idesc_tf32(TC_BM, nt, a_mn, b_mn), pp.kb == 0);
    mma_commit(&sh.bar_free[pp.s]);
    if (pp.kb == nkb - 1) mma_commit(&sh.bar_done);
  }
}

// first n (0..4) floats at p, the rest zero
__device__ __noinline__ float4 ld_guard(const float* __restrict__ p, int n, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n == 4 && vec) return __ldg(reinterpret_cast<const float4*>(p));
  if (n > 0) v.x = __ldg(p);
  if (n > 1) v.y = __ldg(p + 1);
  if (n > 2) v.z = __ldg(p + 2);
  if (n > 3) v.w = __ldg(p + 3);
  return v;
}
// ---- TS mode (forward, dgrad): the 128-row operand lives in tensor memory --------------------------------------
// The accumulator-row operand (activations / dY) is written by the stagers straight from registers into TMEM
// (tcgen05.st: one operand row per thread, 8 contraction elements per store) and read by the MMA as its A
// operand; only the weight tile goes through shared memory.  That takes the 128-row operand off the shared-memory
// port twice (the stagers' stores and the three operand reads per k-step), which is what bounds these kernels.
// TMEM columns: [0, 256) accumulator, [256 + 64 s, 256 + 64 s + 64) stage s of the A operand (32 hi + 32 lo).
constexpr uint32_t TS_TMEM_COLS = 512;
constexpr uint32_t TS_A_COL0 = 256;
__device__ __forceinline__ void tc_issue_ts(uint32_t tmem, int s, uint32_t b_saddr, uint32_t b_bytes, bool b_mn, uint32_t idesc, bool first) {
  const uint32_t ta = tmem + TS_A_COL0 + (uint32_t)s * 64u;
  const uint64_t dbh0 = b_mn ? mnmajor_desc(b_saddr, 0) : kmajor_desc(b_saddr, 0);
  const uint64_t b_lo_d = (uint64_t)(b_bytes >> 4);
  const uint64_t b_step = b_mn ? (1024u >> 4) : ((UMMA_K * 4) >> 4);
#pragma unroll
  for (int ks = 0; ks < KBLK / UMMA_K; ++ks) {
    const uint64_t dbh = dbh0 + ks * b_step;
    mma_tf32_ts(tmem, ta + 32 + 8 * ks, dbh, idesc, (first && ks == 0) ? 0u : 1u);
    mma_tf32_ts(tmem, ta + 8 * ks, dbh + b_lo_d, idesc, 1u);
    mma_tf32_ts(tmem, ta + 8 * ks, dbh, idesc, 1u);
  }
}
__device__ __forceinline__ void tc_publish_issue_ts(TcShared& sh, const TcPipe& pp, uint32_t tmem, uint32_t b_saddr, uint32_t b_bytes, int nkb,
                                                    bool b_mn, int nt, int tid) {
  tmem_st_wait();          // my tcgen05.st of the A operand have landed
  fence_before_sync();
  fence_proxy_async();     // my shared-memory stores of the B operand are visible to the MMA unit
  mbar_arrive(&sh.bar_full[pp.s]);
  if (tid == ((pp.kb & (TC_WARPS - 1)) << 5)) {
    mbar_wait(&sh.bar_full[pp.s], pp.par);
    fence_after_sync();
    tc_issue_ts(tmem, pp.s, b_saddr, b_bytes, b_mn, make_idesc_tf32(TC_BM, nt, false, b_mn), pp.kb == 0);
    mma_commit(&sh.bar_free[pp.s]);
    if (pp.kb == nkb - 1) mma_commit(&sh.bar_done);
  }
}
// This thread's share of the TMEM-resident operand: row 32*(warp%4) + lane, contraction elements [8*(warp/4), +8).
struct TsRow {
  const float* p;      // source pointer for k-block 0 (row clamped into the matrix)
  uint32_t taddr;      // TMEM address of (its lane quarter, column 8*(warp/4)) inside stage 0's hi block
  int k8;              // 8 * (warp / 4)
  bool row_ok;
};
__device__ __forceinline__ TsRow make_tsrow(const float* base, int ld, int row0, int row_end, uint32_t tmem, int warp, int lane) {
  TsRow r;
  const int row = 32 * (warp & 3) + lane;
  r.k8 = 8 * (warp >> 2);
  r.row_ok = row0 + row < row_end;
  r.p = base + (int64_t)min(row0 + row, row_end - 1) * ld + r.k8;
  r.taddr = tmem + TS_A_COL0 + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)r.k8;
  return r;
}
// the 8 contraction elements of k-block kb (offset = kb*32 relative to r.p); c_len = contraction length
__device__ __forceinline__ void tsrow_load(const TsRow& r, const float* __restrict__ p, int kb, int c_len, bool vec, float4& x0, float4& x1) {
  const int k = kb * KBLK + r.k8;
  if (vec && k + 8 <= c_len) {
    x0 = __ldg(reinterpret_cast<const float4*>(p + kb * KBLK));
    x1 = __ldg(reinterpret_cast<const float4*>(p + kb * KBLK + 4));
  } else {
    int n0 = c_len - k; n0 = n0 < 0 ? 0 : (n0 > 4 ? 4 : n0);
    int n1 = c_len - k - 4; n1 = n1 < 0 ? 0 : (n1 > 4 ? 4 : n1);
    x0 = ld_guard(p + kb * KBLK, n0, vec);
    x1 = ld_guard(p + kb * KBLK + 4, n1, vec);
  }
}
__device__ __forceinline__ void tsrow_store(const TsRow& r, int s, float4 x0, float4 x1) {
  float4 h0, l0, h1, l1;
  split_tf32(x0.x, h0.x, l0.x); split_tf32(x0.y, h0.y, l0.y); split_tf32(x0.z, h0.z, l0.z); split_tf32(x0.w, h0.w, l0.w);
  split_tf32(x1.x, h1.x, l1.x); split_tf32(x1.y, h1.y, l1.y); split_tf32(x1.z, h1.z, l1.z); split_tf32(x1.w, h1.w, l1.w);
  const uint32_t ta = r.taddr + (uint32_t)s * 64u;
  tmem_st8(ta, h0, h1);
  tmem_st8(ta + 32, l0, l1);
}

// accumulator tile TMEM -> shared memory ot[128][ldo] (row = accumulator row); nt multiple of 16
__device__ __forceinline__ void tc_drain(uint32_t tmem, float* ot, int ldo, int nt, int warp, int lane) {
  const int lane_base = 32 * (warp & 3);
  float* orow = ot + (size_t)(lane_base + lane) * ldo;
  for (int cb = (warp >> 2) * 32; cb < nt; cb += 32 * (TC_WARPS / 4)) {
    const uint32_t taddr = tmem + ((uint32_t)lane_base << 16) + (uint32_t)cb;
    if (cb + 32 <= nt) {
      uint32_t r[32];
      tmem_ld32(taddr, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(orow + cb + 4 * i) =
            make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
    } else {
      uint32_t r[16];
      tmem_ld16(taddr, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(orow + cb + 4 * i) =
            make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
    }
  }
}

// per-warp column partials -> one fp64 atomic per column and CTA.  red: [2][TC_WARPS][nt] doubles.
__device__ __forceinline__ void tc_col_atomics(const double* red, double* gstats, int col0, int nvalid, int nt, int tid) {
  for (int i = tid; i < 2 * nvalid; i += TC_NP) {
    const int which = i / nvalid, col = i - which * nvalid;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < TC_WARPS; ++w) t += red[(which * TC_WARPS + w) * nt + col];
    atomicAdd(gstats + 2 * (col0 + col) + which, t);
  }
}

__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {   // one 16-byte reduction instead of four 4-byte ones
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- what one thread moves per k-block ---------------------------------------------------------------------
// An operand tile is 32 contraction elements x R rows (R = 128 for the accumulator-row operand, NT rounded up to 64 for
// the accumulator-column operand) = R * 8 float4; thread tid moves float4 number it * 512 + tid of it, it < R / 64.
//   K-major  (source rows = M/N index, contraction contiguous): row it*64 + (tid >> 3), contraction chunk tid & 7
//   MN-major (source rows = contraction index, M/N contiguous): contraction row c = v / nq, M/N quad q = v % nq, nq = R / 4
// Fast path (interior k-blocks of 16-byte aligned sources): one unconditional 16-byte load per float4 from a pointer
// fixed at kernel start and advanced by a constant per k-block.  Rows / quads outside the matrix are clamped onto
// the last valid one instead of predicated: they only produce accumulator rows / columns that are never stored.
// Slow path (the last, partial k-block of a contraction; unaligned or odd-width sources): guarded loads, zero fill.
template <int IT>
struct Op {
  const float* p[IT];   // fast-path source pointer for k-block 0
  uint32_t so[IT];      // shared-memory byte offset inside the tile
};
template <int IT>
__device__ __forceinline__ void op_init_k(Op<IT>& o, const float* base, int ld, int row0, int row_end, int nit, int tid) {
  const int cj = tid & 7, r0 = tid >> 3;
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int r = it * TC_RPI + r0;
    const int rc = min(row0 + r, row_end - 1);
    o.p[it] = base + (int64_t)rc * ld + 4 * cj;
    o.so[it] = kmajor_off(r, cj);
  }
}
template <int IT>
__device__ __forceinline__ void op_init_m(Op<IT>& o, const float* base /* source + first M/N index of the tile */, int ld, int mn_len, int nq, int nit, int tid) {
  const int qmax = max(mn_len / 4 - 1, 0);
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    if (it < nit) {
      const int v = it * TC_NP + tid;
      const int c = v / nq, q = v - c * nq;
      o.p[it] = base + (int64_t)c * ld + 4 * min(q, qmax);
      o.so[it] = mnmajor_off(q, c);
    } else { o.p[it] = base; o.so[it] = 0; }
  }
}
template <int IT>
__device__ __forceinline__ void op_load_fast(const Op<IT>& o, int64_t adv, int nit, float4 (&r)[IT]) {
#pragma unroll
  for (int it = 0; it < IT; ++it)
    if (it < nit) r[it] = __ldg(reinterpret_cast<const float4*>(o.p[it] + adv));
}
// K-major slow path: rows [row0, row_end) x contraction [c0, c_len) of base[., ld]
template <int IT>
__device__ __forceinline__ void op_load_slow_k(const float* base, int ld, int row0, int row_end, int c0, int c_len, bool vec, int nit, int tid, float4 (&r)[IT]) {
  const int cj = tid & 7, r0 = tid >> 3;
  int dyn = c_len - c0 - 4 * cj;
  dyn = dyn < 0 ? 0 : (dyn > 4 ? 4 : dyn);
#pragma unroll
  for (int it = 0; it < IT; ++it)
    if (it < nit) {
      const int row = row0 + it * TC_RPI + r0;
      r[it] = ld_guard(base + (int64_t)row * ld + c0 + 4 * cj, row < row_end ? dyn : 0, vec);
    }
}
// MN-major slow path: contraction rows [c0, c_len) x M/N [mn0, mn_end) of base[., ld]
template <int IT>
__device__ __forceinline__ void op_load_slow_m(const float* base, int ld, int mn0, int mn_end, int c0, int c_len, int nq, bool vec, int nit, int tid, float4 (&r)[IT]) {
#pragma unroll
  for (int it = 0; it < IT; ++it)
    if (it < nit) {
      const int v = it * TC_NP + tid;
      const int c = v / nq, q = v - c * nq;
      int n = mn_end - (mn0 + 4 * q);
      n = n < 0 ? 0 : (n > 4 ? 4 : n);
      r[it] = ld_guard(base + (int64_t)(c0 + c) * ld + mn0 + 4 * q, (c0 + c < c_len) ? n : 0, vec);
    }
}

// max(z, slope * z): relu (slope 0), identity (1), leaky (0.1); sigmoid handled apart
__device__ __forceinline__ float act_slope(int act) { return act == SWR_ACT_RELU ? 0.f : (act == SWR_ACT_LEAKY ? 0.1f : 1.f); }
__device__ __forceinline__ float4 norm_act4(float4 x, float4 mu, float4 sc, float4 bb, float slope, bool sigmoid) {
  float4 z;
  z.x = fmaf(x.x - mu.x, sc.x, bb.x); z.y = fmaf(x.y - mu.y, sc.y, bb.y);
  z.z = fmaf(x.z - mu.z, sc.z, bb.z); z.w = fmaf(x.w - mu.w, sc.w, bb.w);
  if (sigmoid) {
    z.x = 1.f / (1.f + expf(-z.x)); z.y = 1.f / (1.f + expf(-z.y)); z.z = 1.f / (1.f + expf(-z.z)); z.w = 1.f / (1.f + expf(-z.w));
    return z;
  }
  z.x = fmaxf(z.x, slope * z.x); z.y = fmaxf(z.y, slope * z.y); z.z = fmaxf(z.z, slope * z.z); z.w = fmaxf(z.w, slope * z.w);
  return z;
}
__device__ __forceinline__ float4 affine4(float4 c0, float4 dz, float4 c1, float4 raw, float4 c2) {
  return make_float4(fmaf(c0.x, dz.x, fmaf(c1.x, raw.x, c2.x)), fmaf(c0.y, dz.y, fmaf(c1.y, raw.y, c2.y)),
                     fmaf(c0.z, dz.z, fmaf(c1.z, raw.z, c2.z)), fmaf(c0.w, dz.w, fmaf(c1.w, raw.w, c2.w)));
}

// ---------------------------------------------------------------------------------------
// forward:  Y[m, n] = sum_k act(norm(A))[m, k] * Weff[n, k] + beff[n]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_NT, 1) fc_tc_fwd_kernel(const __grid_constant__ TcParams p) {
  const long long t_kernel = clock64();
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  uint8_t* smem = align1024(smem_raw);
  const uint32_t smem_s = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int g = 0;
  while (g + 1 < p.n_groups && p.tile_start[g + 1] <= (int)blockIdx.x) ++g;
  const FcGroup& G = p.g[g];
  const int M = p.B, N = G.Y.n, K = G.A.n;
  const int NT = p.nt[g], NTr = (NT + 63) & ~63, nq = NTr >> 2, nitb = NTr >> 6;
  const int nt_n = (N + NT - 1) / NT;
  const int local = blockIdx.x - p.tile_start[g];
  const int m0 = (local / nt_n) * TC_BM, n0 = (local % nt_n) * NT;
  const int nkb = (K + KBLK - 1) / KBLK, Kpad = nkb * KBLK;
  const uint32_t b_bytes = (uint32_t)NTr * 128u, stage_bytes = 2 * b_bytes;   // TS mode: only the weight tile is in shared memory
  const int S = p.stages;
  float* kc = reinterpret_cast<float*>(smem + (size_t)S * stage_bytes);   // [3][Kpad]: mu, s, b of the input columns

  const bool plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
  if (!plainA) {
    for (int k = tid; k < Kpad; k += TC_NT) {
      ColCoef c = {0.f, 0.f, 0.f, 0.f};
      if (k < K) c = col_coef(G.A.norm, k, p.inv_count);
      kc[k] = c.mu; kc[Kpad + k] = c.s; kc[2 * Kpad + k] = c.b;
    }
  }
  const uint32_t tmem = tc_setup(sh, S, TS_TMEM_COLS, tid);
  const bool kn = (G.w_layout == SWR_W_KN);

  const int Nend = min(N, n0 + NT);
  const bool vecA = is_al16(G.A.raw) && (G.A.ld % 4 == 0);
  const bool vecW = is_al16(G.W) && (G.ldw % 4 == 0) && (!G.W2 || is_al16(G.W2));
  const bool fastB = vecW && (!kn || ((Nend - n0) % 4 == 0));
  const bool hasW2 = G.W2 != nullptr;
  const float slope = act_slope(G.A.act);
  const bool sigA = G.A.act == SWR_ACT_SIGMOID;
  const int64_t w2diff = hasW2 ? (G.W2 - G.W) : 0;
  Op<4> ob;
  const TsRow ta = make_tsrow(G.A.raw, G.A.ld, m0, M, tmem, warp, lane);
  if (!kn) op_init_k<4>(ob, G.W, G.ldw, n0, Nend, nitb, tid);             // W [N, K] -> K-major tile
  else op_init_m<4>(ob, G.W + n0, G.ldw, Nend - n0, nq, nitb, tid);       // W [K, N] -> MN-major tile
  const int64_t bstep = kn ? (int64_t)KBLK * G.ldw : KBLK;

  // register prefetch one k-block ahead; W2 (STAR / M3oE only) is read at store time
  struct Regs { float4 a0, a1, b[4]; };
  auto load = [&](int kb, Regs& r) {
    const bool interior = (kb + 1) * KBLK <= K;
    tsrow_load(ta, ta.p, kb, K, vecA, r.a0, r.a1);
    if (fastB && interior) op_load_fast<4>(ob, kb * bstep, nitb, r.b);
    else if (!kn) op_load_slow_k<4>(G.W, G.ldw, n0, Nend, kb * KBLK, K, vecW, nitb, tid, r.b);
    else op_load_slow_m<4>(G.W, G.ldw, n0, Nend, kb * KBLK, K, nq, vecW, nitb, tid, r.b);
  };
  auto store = [&](int kb, uint32_t stage, int stage_idx, Regs& r) {
    const uint32_t bh = stage, bl = bh + b_bytes;
    if (plainA) {
      tsrow_store(ta, stage_idx, r.a0, r.a1);
    } else {   // padded k: coefficients are 0 -> act(0) stays finite and meets a zero weight
      const int k = kb * KBLK + ta.k8;
      const float4 x0 = norm_act4(r.a0, ld4s(kc + k), ld4s(kc + Kpad + k), ld4s(kc + 2 * Kpad + k), slope, sigA);
      const float4 x1 = norm_act4(r.a1, ld4s(kc + k + 4), ld4s(kc + Kpad + k + 4), ld4s(kc + 2 * Kpad + k + 4), slope, sigA);
      tsrow_store(ta, stage_idx, x0, x1);
    }
    if (hasW2) {
      float4 w2[4];
      const bool interior = (kb + 1) * KBLK <= K;
      if (fastB && interior) op_load_fast<4>(ob, kb * bstep + w2diff, nitb, w2);
      else if (!kn) op_load_slow_k<4>(G.W2, G.ldw, n0, Nend, kb * KBLK, K, vecW, nitb, tid, w2);
      else op_load_slow_m<4>(G.W2, G.ldw, n0, Nend, kb * KBLK, K, nq, vecW, nitb, tid, w2);
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (it < nitb) r.b[it] = mul4(r.b[it], w2[it]);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it)
      if (it < nitb) store_split(bh, bl, ob.so[it], r.b[it]);
  };

  TcPipe pp; pp.init();
  Regs r0;
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && (tid == 0 || tid == 33 || tid == 511);
  long long t_acq = 0, t_store = 0, t_load = 0, t_pub = 0, t0 = 0, t_begin = dbg ? clock64() : 0;
  load(0, r0);
  for (; pp.kb < nkb; pp.advance(S)) {
    if (dbg) t0 = clock64();
    tc_acquire(sh, pp, S);
    if (dbg) { const long long t = clock64(); t_acq += t - t0; t0 = t; }
    const uint32_t stage = smem_s + (uint32_t)pp.s * stage_bytes;
    store(pp.kb, stage, pp.s, r0);
    if (dbg) { const long long t = clock64(); t_store += t - t0; t0 = t; }
    if (pp.kb + 1 < nkb) load(pp.kb + 1, r0);
    if (dbg) { const long long t = clock64(); t_load += t - t0; t0 = t; }
    tc_publish_issue_ts(sh, pp, tmem, stage, b_bytes, nkb, kn, NT, tid);
    if (dbg) { const long long t = clock64(); t_pub += t - t0; t0 = t; }
  }
  const long long t_loop_end = dbg ? clock64() : 0;
  mbar_wait(&sh.bar_done, 0);
  fence_after_sync();
  if (dbg) {
    long long* o = p.dbg + (tid == 0 ? 0 : (tid == 33 ? 8 : 16));
    o[0] = t_acq; o[1] = t_store; o[2] = t_load; o[3] = t_pub; o[4] = t_loop_end - t_begin; o[5] = clock64() - t_loop_end; o[6] = nkb; o[7] = t_begin - t_kernel;
  }

  // ---- epilogue: bias, optional activation, store raw Y, fp64 column moments ----
  const int ldo = NT + 4;
  float* ot = reinterpret_cast<float*>(smem);
  double* red = reinterpret_cast<double*>(ot + (size_t)TC_BM * ldo);
  tc_drain(tmem, ot, ldo, NT, warp, lane);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TS_TMEM_COLS);
  float* Y = const_cast<float*>(G.Y.raw);
  const int nvalid = Nend - n0;
  constexpr int RPW = TC_BM / TC_WARPS;   // accumulator rows per warp
  // Moments of a warp's 8 rows: fp32 sums of the values centred on the first one (no cancellation in the second
  // moment), widened to fp64 once per column -- the fp64 pipe is too slow to take every element.
  const bool vec_out = (nvalid % 4 == 0) && (G.Y.ld % 4 == 0) && is_al16(Y);
  if (vec_out) {   // 4 columns per thread: 16-byte shared-memory reads and global stores
    for (int cbase = 0; cbase < nvalid; cbase += 128) {
      const int col = cbase + 4 * lane, n = n0 + col;
      if (col < nvalid) {
        float4 bias;
        bias.x = ld_opt(G.bias, n, 0.f) + ld_opt(G.bias2, n, 0.f); bias.y = ld_opt(G.bias, n + 1, 0.f) + ld_opt(G.bias2, n + 1, 0.f);
        bias.z = ld_opt(G.bias, n + 2, 0.f) + ld_opt(G.bias2, n + 2, 0.f); bias.w = ld_opt(G.bias, n + 3, 0.f) + ld_opt(G.bias2, n + 3, 0.f);
        float4 t1 = zero4(), t2 = zero4(), y0 = zero4();
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const int row = warp * RPW + i, m = m0 + row;
          if (m < M) {
            float4 y = ld4s(ot + (size_t)row * ldo + col);
            y.x += bias.x; y.y += bias.y; y.z += bias.z; y.w += bias.w;
            if (G.e_act != SWR_ACT_NONE) {
              y.x = act_fwd(y.x, G.e_act) * G.e_scale; y.y = act_fwd(y.y, G.e_act) * G.e_scale;
              y.z = act_fwd(y.z, G.e_act) * G.e_scale; y.w = act_fwd(y.w, G.e_act) * G.e_scale;
            }
            *reinterpret_cast<float4*>(Y + (int64_t)m * G.Y.ld + n) = y;
            if (cnt == 0) y0 = y;
            const float4 d = make_float4(y.x - y0.x, y.y - y0.y, y.z - y0.z, y.w - y0.w);
            t1.x += d.x; t1.y += d.y; t1.z += d.z; t1.w += d.w;
            t2.x = fmaf(d.x, d.x, t2.x); t2.y = fmaf(d.y, d.y, t2.y); t2.z = fmaf(d.z, d.z, t2.z); t2.w = fmaf(d.w, d.w, t2.w);
            ++cnt;
          }
        }
        // sum y = cnt*y0 + t1 ; sum y^2 = cnt*y0^2 + 2*y0*t1 + t2
        const float y0a[4] = {y0.x, y0.y, y0.z, y0.w}, t1a[4] = {t1.x, t1.y, t1.z, t1.w}, t2a[4] = {t2.x, t2.y, t2.z, t2.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double dy0 = (double)y0a[j], dt1 = (double)t1a[j];
          red[(0 * TC_WARPS + warp) * NT + col + j] = (double)cnt * dy0 + dt1;
          red[(1 * TC_WARPS + warp) * NT + col + j] = (double)cnt * dy0 * dy0 + 2.0 * dy0 * dt1 + (double)t2a[j];
        }
      }
    }
  } else {
    for (int cbase = 0; cbase < nvalid; cbase += 32) {
      const int col = cbase + lane, n = n0 + col;
      if (col < nvalid) {
        const float bias = ld_opt(G.bias, n, 0.f) + ld_opt(G.bias2, n, 0.f);
        float t1 = 0.f, t2 = 0.f, y0 = 0.f;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const int row = warp * RPW + i, m = m0 + row;
          if (m < M) {
            float y = ot[(size_t)row * ldo + col] + bias;
            if (G.e_act != SWR_ACT_NONE) y = act_fwd(y, G.e_act) * G.e_scale;
            Y[(int64_t)m * G.Y.ld + n] = y;
            if (cnt == 0) y0 = y;
            const float d = y - y0;
            t1 += d; t2 = fmaf(d, d, t2);
            ++cnt;
          }
        }
        const double dy0 = (double)y0, dt1 = (double)t1;
        red[(0 * TC_WARPS + warp) * NT + col] = (double)cnt * dy0 + dt1;
        red[(1 * TC_WARPS + warp) * NT + col] = (double)cnt * dy0 * dy0 + 2.0 * dy0 * dt1 + (double)t2;
      }
    }
  }
  const long long t_epi1 = dbg ? clock64() : 0;
  if (G.stats_out) {
    __syncthreads();
    tc_col_atomics(red, G.stats_out, n0, nvalid, NT, tid);
  }
  if (dbg) {
    long long* o = p.dbg + 24 + (tid == 0 ? 0 : (tid == 33 ? 4 : 8));
    o[0] = t_epi1 - t_kernel; o[1] = clock64() - t_kernel;
  }
}

// ---------------------------------------------------------------------------------------
// data gradient: dA[m, j] = sum_g sum_n dY_g[m, n] * Weff_g[n, j]   (fan-in over the groups of one destination)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_NT, 1) fc_tc_dgrad_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  uint8_t* smem = align1024(smem_raw);
  const uint32_t smem_s = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int d = 0;
  while (d + 1 < p.n_dst && p.dst_tile[d + 1] <= (int)blockIdx.x) ++d;
  const int gs = p.dst_group[d], ge = p.dst_group[d + 1];
  const ActDev& D = p.g[gs].A;
  const int M = p.B, Kd = D.n;
  const int NT = p.nt[d], NTr = (NT + 63) & ~63, nq = NTr >> 2, nitb = NTr >> 6;
  const int nt_n = (Kd + NT - 1) / NT;
  const int local = blockIdx.x - p.dst_tile[d];
  const int m0 = (local / nt_n) * TC_BM, j0 = (local % nt_n) * NT;
  const int kb0 = p.tile_start[gs];
  const int nkb = p.tile_start[ge] - kb0, Kc = nkb * KBLK;
  const uint32_t b_bytes = (uint32_t)NTr * 128u, stage_bytes = 2 * b_bytes;   // TS mode: only the weight tile is in shared memory
  const int S = p.stages;
  float* dc = reinterpret_cast<float*>(smem + (size_t)S * stage_bytes);   // [3][Kc]: c0, c1, c2 over the concatenated group columns

  for (int g = gs; g < ge; ++g) {
    const FcGroup& G = p.g[g];
    const int base = (p.tile_start[g] - kb0) * KBLK, span = (p.tile_start[g + 1] - p.tile_start[g]) * KBLK;
    for (int n = tid; n < span; n += TC_NT) {
      DyCoef c = {0.f, 0.f, 0.f};
      if (n < G.Y.n) c = dy_coef(G.Y, n, p.inv_count);
      dc[base + n] = c.c0; dc[Kc + base + n] = c.c1; dc[2 * Kc + base + n] = c.c2;
    }
  }
  // coefficients of the destination's own norm / activation for the epilogue: one column per thread, once per CTA
  const int Jend = min(Kd, j0 + NT);
  const bool has_norm = D.norm.mode != SWR_NORM_NONE;
  const bool plainD = !has_norm && D.act == SWR_ACT_NONE;
  float* ccs = reinterpret_cast<float*>(smem + p.epi_off);   // [4][NT]: mu, s, b, r
  for (int c = tid; c < NT; c += TC_NT) {
    ColCoef cc = {0.f, 1.f, 0.f, 1.f};
    if (!plainD && j0 + c < Jend) cc = col_coef(D.norm, j0 + c, p.inv_count);
    ccs[c] = cc.mu; ccs[NT + c] = cc.s; ccs[2 * NT + c] = cc.b; ccs[3 * NT + c] = cc.r;
  }
  const uint32_t tmem = tc_setup(sh, S, TS_TMEM_COLS, tid);

  // per-group staging state (rebuilt when the k-block walk enters the next group of the fan-in)
  int cur_g = gs - 1, g_kb0 = 0, g_next = 0, g_N = 0;
  bool kn = false, hasW2 = false, need_raw = false, vecY = false, vecW = false, fastB = false;
  TsRow ta{}; Op<4> ob;
  int64_t w2diff = 0, rawdiff = 0, bstep = 0;
  // shared-memory offsets do not depend on the group, only on the layout of its weight
  uint32_t so_bk[4], so_bm[4];
  {
    const int cj = tid & 7, r0_ = tid >> 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      so_bk[it] = kmajor_off(it * TC_RPI + r0_, cj);
      const int v = it * TC_NP + tid, c = v / nq, q = v - c * nq;
      so_bm[it] = (it < nitb) ? mnmajor_off(q, c) : 0u;
    }
  }
  auto enter_group = [&](int g) {
    const FcGroup& G = p.g[g];
    cur_g = g; g_kb0 = p.tile_start[g] - kb0; g_next = p.tile_start[g + 1] - kb0; g_N = G.Y.n;
    kn = (G.w_layout == SWR_W_KN); hasW2 = G.W2 != nullptr;
    need_raw = (G.Y.norm.mode == SWR_NORM_BATCH);
    vecY = is_al16(G.Y.dz) && is_al16(G.Y.raw) && (G.Y.ld % 4 == 0);
    vecW = is_al16(G.W) && (G.ldw % 4 == 0) && (!G.W2 || is_al16(G.W2));
    fastB = vecW && (kn || ((Jend - j0) % 4 == 0));
    rawdiff = G.Y.raw - G.Y.dz;
    w2diff = hasW2 ? (G.W2 - G.W) : 0;
    ta = make_tsrow(G.Y.dz, G.Y.ld, m0, M, tmem, warp, lane);
    if (!kn) { op_init_m<4>(ob, G.W + j0, G.ldw, Jend - j0, nq, nitb, tid); bstep = (int64_t)KBLK * G.ldw; }   // W[n, j] -> MN-major
    else { op_init_k<4>(ob, G.W, G.ldw, j0, Jend, nitb, tid); bstep = KBLK; }                                     // W[j, n] -> K-major
  };
  // register prefetch one k-block ahead; the set remembers the layout of the block it holds.  The W2 product
  // (STAR / M3oE) is applied at load time: those loads then wait in place, which only those models pay for.
  struct Regs { float4 a0, a1, r0, r1, b[4]; bool b_mn; };
  auto load = [&](int kb, Regs& R) {
    while (cur_g < gs || kb >= g_next) enter_group(cur_g + 1);
    const FcGroup& G = p.g[cur_g];
    const int lkb = kb - g_kb0;   // k-block inside the group
    const bool interior = (lkb + 1) * KBLK <= g_N;
    R.b_mn = !kn;
    tsrow_load(ta, ta.p, lkb, g_N, vecY, R.a0, R.a1);
    if (need_raw) tsrow_load(ta, ta.p + rawdiff, lkb, g_N, vecY, R.r0, R.r1);
    else { R.r0 = zero4(); R.r1 = zero4(); }
    float4 w2[4];
    if (fastB && interior) {
      op_load_fast<4>(ob, lkb * bstep, nitb, R.b);
      if (hasW2) op_load_fast<4>(ob, lkb * bstep + w2diff, nitb, w2);
    } else if (!kn) {
      op_load_slow_m<4>(G.W, G.ldw, j0, Jend, lkb * KBLK, g_N, nq, vecW, nitb, tid, R.b);
      if (hasW2) op_load_slow_m<4>(G.W2, G.ldw, j0, Jend, lkb * KBLK, g_N, nq, vecW, nitb, tid, w2);
    } else {
      op_load_slow_k<4>(G.W, G.ldw, j0, Jend, lkb * KBLK, g_N, vecW, nitb, tid, R.b);
      if (hasW2) op_load_slow_k<4>(G.W2, G.ldw, j0, Jend, lkb * KBLK, g_N, vecW, nitb, tid, w2);
    }
    if (hasW2) {
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (it < nitb) R.b[it] = mul4(R.b[it], w2[it]);
    }
  };
  auto store = [&](int kb, uint32_t stage, int stage_idx, Regs& R) {
    const uint32_t bh = stage, bl = bh + b_bytes;
    const int k = kb * KBLK + ta.k8;
    // rows outside the batch only feed accumulator rows that are never stored
    tsrow_store(ta, stage_idx,
                affine4(ld4s(dc + k), R.a0, ld4s(dc + Kc + k), R.r0, ld4s(dc + 2 * Kc + k)),
                affine4(ld4s(dc + k + 4), R.a1, ld4s(dc + Kc + k + 4), R.r1, ld4s(dc + 2 * Kc + k + 4)));
#pragma unroll
    for (int it = 0; it < 4; ++it)
      if (it < nitb) store_split(bh, bl, R.b_mn ? so_bm[it] : so_bk[it], R.b[it]);
  };

  TcPipe pp; pp.init();
  Regs r0;
  load(0, r0);
  for (; pp.kb < nkb; pp.advance(S)) {
    tc_acquire(sh, pp, S);
    const uint32_t stage = smem_s + (uint32_t)pp.s * stage_bytes;
    const bool b_mn = r0.b_mn;   // layout of the block being stored (load(kb + 1) overwrites it)
    store(pp.kb, stage, pp.s, r0);
    if (pp.kb + 1 < nkb) load(pp.kb + 1, r0);
    tc_publish_issue_ts(sh, pp, tmem, stage, b_bytes, nkb, b_mn, NT, tid);
  }
  mbar_wait(&sh.bar_done, 0);
  fence_after_sync();

  // ---- epilogue: backward through the destination's activation / norm (stage 1), store dz ----
  const int ldo = NT + 4;
  float* ot = reinterpret_cast<float*>(smem);
  double* red = reinterpret_cast<double*>(ot + (size_t)TC_BM * ldo);
  tc_drain(tmem, ot, ldo, NT, warp, lane);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TS_TMEM_COLS);
  const bool accumulate = (p.g[gs].flags & FC_A_ACCUMULATE) != 0;
  const bool atomic_dst = (p.dst_atomic >> d) & 1u;
  const int nvalid = Jend - j0;
  constexpr int RPW = TC_BM / TC_WARPS;
  const bool vec_out = (nvalid % 4 == 0) && (D.ld % 4 == 0) && is_al16(D.dz) && is_al16(D.raw);
  if (vec_out) {   // 4 columns per thread: 16-byte shared-memory reads, global loads / stores / reductions
    for (int cbase = 0; cbase < nvalid; cbase += 128) {
      const int col = cbase + 4 * lane, j = j0 + col;
      if (col < nvalid) {
        const float4 mu = ld4s(ccs + col), sc = ld4s(ccs + NT + col), bb = ld4s(ccs + 2 * NT + col), rr = ld4s(ccs + 3 * NT + col);
        float4 s1 = zero4(), s2 = zero4();
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const int row = warp * RPW + i, m = m0 + row;
          if (m < M) {
            float* dst = D.dz + (int64_t)m * D.ld + j;
            float4 dz = ld4s(ot + (size_t)row * ldo + col);
            if (!plainD) {
              const float4 raw = *reinterpret_cast<const float4*>(D.raw + (int64_t)m * D.ld + j);
              dz.x *= act_grad(fmaf(raw.x - mu.x, sc.x, bb.x), D.act); dz.y *= act_grad(fmaf(raw.y - mu.y, sc.y, bb.y), D.act);
              dz.z *= act_grad(fmaf(raw.z - mu.z, sc.z, bb.z), D.act); dz.w *= act_grad(fmaf(raw.w - mu.w, sc.w, bb.w), D.act);
              s1.x += dz.x; s1.y += dz.y; s1.z += dz.z; s1.w += dz.w;
              s2.x = fmaf(dz.x, (raw.x - mu.x) * rr.x, s2.x); s2.y = fmaf(dz.y, (raw.y - mu.y) * rr.y, s2.y);
              s2.z = fmaf(dz.z, (raw.z - mu.z) * rr.z, s2.z); s2.w = fmaf(dz.w, (raw.w - mu.w) * rr.w, s2.w);
            }
            if (atomic_dst) { red_add_v4(dst, dz); continue; }   // plain destination split over its fan-in
            if (accumulate) { const float4 o = *reinterpret_cast<const float4*>(dst); dz.x += o.x; dz.y += o.y; dz.z += o.z; dz.w += o.w; }
            *reinterpret_cast<float4*>(dst) = dz;
          }
        }
        const float s1a[4] = {s1.x, s1.y, s1.z, s1.w}, s2a[4] = {s2.x, s2.y, s2.z, s2.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          red[(0 * TC_WARPS + warp) * NT + col + jj] = (double)s1a[jj];
          red[(1 * TC_WARPS + warp) * NT + col + jj] = (double)s2a[jj];
        }
      }
    }
  } else {
    for (int cbase = 0; cbase < nvalid; cbase += 32) {
      const int col = cbase + lane, j = j0 + col;
      if (col < nvalid) {
        const ColCoef cc = {ccs[col], ccs[NT + col], ccs[2 * NT + col], ccs[3 * NT + col]};
        float s1 = 0.f, s2 = 0.f;   // 8-row partials in fp32, widened to fp64 once per column
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const int row = warp * RPW + i, m = m0 + row;
          if (m < M) {
            const int64_t o = (int64_t)m * D.ld + j;
            float dz = ot[(size_t)row * ldo + col];
            if (!plainD) {
              const float raw = D.raw[o];
              dz *= act_grad(fmaf(raw - cc.mu, cc.s, cc.b), D.act);
              s1 += dz; s2 = fmaf(dz, (raw - cc.mu) * cc.r, s2);
            }
            if (atomic_dst) { atomicAdd(D.dz + o, dz); continue; }   // plain destination split over its fan-in
            if (accumulate) dz += D.dz[o];
            D.dz[o] = dz;
          }
        }
        red[(0 * TC_WARPS + warp) * NT + col] = (double)s1;
        red[(1 * TC_WARPS + warp) * NT + col] = (double)s2;
      }
    }
  }
  if (has_norm && D.dstats) {
    __syncthreads();
    tc_col_atomics(red, D.dstats, j0, nvalid, NT, tid);
  }
}

// ---------------------------------------------------------------------------------------
// weight / bias gradient: dWeff[n, j] = sum_b dY[b, n] * act(norm(A))[b, j],  db[n] = sum_b dY[b, n]
// accumulator rows = output features n, columns = input features j, contraction = batch rows of this split
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_NT, 1) fc_tc_wgrad_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  __shared__ float bsum[TC_BM];
  uint8_t* smem = align1024(smem_raw);
  const uint32_t smem_s = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int g = 0;
  while (g + 1 < p.n_groups && p.tile_start[g + 1] <= (int)blockIdx.x) ++g;
  const FcGroup& G = p.g[g];
  const int N = G.Y.n, K = G.A.n;
  const int NT = p.nt[g], NTr = (NT + 63) & ~63, nq = NTr >> 2, nitb = NTr >> 6;
  const int nt_m = (N + TC_BM - 1) / TC_BM, nt_n = (K + NT - 1) / NT;
  int local = blockIdx.x - p.tile_start[g];
  const int split = local / (nt_m * nt_n);
  local -= split * nt_m * nt_n;
  const int m0 = (local / nt_n) * TC_BM, j0 = (local % nt_n) * NT;
  const int b_begin = split * p.rows_per_split;
  const int b_end = min(p.B, b_begin + p.rows_per_split);
  if (b_begin >= b_end) return;   // whole CTA, before any barrier
  const int rows = b_end - b_begin;
  const int nkb = (rows + KBLK - 1) / KBLK;
  const uint32_t b_bytes = (uint32_t)NTr * 128u, stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
  const int S = p.stages;
  float* nc = reinterpret_cast<float*>(smem + (size_t)S * stage_bytes);   // [3][NTr]: mu, s, b of the input columns j

  const bool plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
  for (int i = tid; i < NTr; i += TC_NT) {
    ColCoef c = {0.f, 0.f, 0.f, 0.f};
    if (j0 + i < K) c = plainA ? ColCoef{0.f, 1.f, 0.f, 1.f} : col_coef(G.A.norm, j0 + i, p.inv_count);
    nc[i] = c.mu; nc[NTr + i] = c.s; nc[2 * NTr + i] = c.b;
  }
  float* mcs = nc + 3 * NTr;   // [3][128]: c0, c1, c2 of the dY rows (output features) of this tile
  if (tid < TC_BM) {
    bsum[tid] = 0.f;
    DyCoef c = {0.f, 0.f, 0.f};
    if (m0 + tid < N) c = dy_coef(G.Y, m0 + tid, p.inv_count);
    mcs[tid] = c.c0; mcs[TC_BM + tid] = c.c1; mcs[2 * TC_BM + tid] = c.c2;
  }
  const uint32_t tmem = tc_setup(sh, S, tmem_cols(NT), tid);

  // A(n, b) = dY[b, n]: n-contiguous -> MN-major tile; this thread always stages the same 4 output features
  // (quad aq of contraction rows it*16 + tid/32), so their dY coefficients live in registers
  const int aq = tid & 31;
  const int Nend = min(N, m0 + TC_BM), Jend = min(K, j0 + NT);
  const float4 c0 = ld4s(mcs + 4 * aq), c1 = ld4s(mcs + TC_BM + 4 * aq), c2 = ld4s(mcs + 2 * TC_BM + 4 * aq);
  const bool vecY = is_al16(G.Y.dz) && is_al16(G.Y.raw) && (G.Y.ld % 4 == 0);
  const bool vecA = is_al16(G.A.raw) && (G.A.ld % 4 == 0);
  const bool fastA = vecY && ((Nend - m0) % 4 == 0), fastB = vecA && ((Jend - j0) % 4 == 0);
  const bool need_raw = (G.Y.norm.mode == SWR_NORM_BATCH);
  const float slope = act_slope(G.A.act);
  const bool sigA = G.A.act == SWR_ACT_SIGMOID;
  const float* dzsrc = G.Y.dz + (int64_t)b_begin * G.Y.ld;    // contraction row 0 of this split
  const float* rawsrc = G.Y.raw + (int64_t)b_begin * G.Y.ld;
  const float* asrc = G.A.raw + (int64_t)b_begin * G.A.ld;
  Op<2> oa; Op<4> ob;
  op_init_m<2>(oa, dzsrc + m0, G.Y.ld, Nend - m0, 32, 2, tid);
  op_init_m<4>(ob, asrc + j0, G.A.ld, Jend - j0, nq, nitb, tid);
  const int64_t rawdiff = G.Y.raw - G.Y.dz;
  const int64_t astep = (int64_t)KBLK * G.Y.ld, bstep = (int64_t)KBLK * G.A.ld;
  // input-column quads this thread stages (for the lazy activation coefficients)
  uint32_t qpack = 0;
#pragma unroll
  for (int it = 0; it < 4; ++it) qpack |= (uint32_t)((it * TC_NP + tid) % nq) << (8 * it);

  struct Regs { float4 a[2], r[2], b[4]; };
  float4 rowsum = zero4();
  auto load = [&](int kb, Regs& R) {
    const bool interior = (kb + 1) * KBLK <= rows;
    if (fastA && interior) {
      op_load_fast<2>(oa, kb * astep, 2, R.a);
      if (need_raw) op_load_fast<2>(oa, kb * astep + rawdiff, 2, R.r);
    } else {
      op_load_slow_m<2>(dzsrc, G.Y.ld, m0, Nend, kb * KBLK, rows, 32, vecY, 2, tid, R.a);
      if (need_raw) op_load_slow_m<2>(rawsrc, G.Y.ld, m0, Nend, kb * KBLK, rows, 32, vecY, 2, tid, R.r);
    }
    if (!need_raw) { R.r[0] = zero4(); R.r[1] = zero4(); }
    if (fastB && interior) op_load_fast<4>(ob, kb * bstep, nitb, R.b);
    else op_load_slow_m<4>(asrc, G.A.ld, j0, Jend, kb * KBLK, rows, nq, vecA, nitb, tid, R.b);
  };
  auto store = [&](int kb, uint32_t stage, Regs& R) {
    const uint32_t ah = stage, al = ah + TC_A_BYTES, bh = al + TC_A_BYTES, bl = bh + b_bytes;
    const int left = rows - kb * KBLK;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c = it * (TC_NP / 32) + (tid >> 5);
      float4 x = zero4();
      if (c < left) x = affine4(c0, R.a[it], c1, R.r[it], c2);   // c2 != 0: contraction padding must stay exactly zero
      rowsum.x += x.x; rowsum.y += x.y; rowsum.z += x.z; rowsum.w += x.w;
      store_split(ah, al, oa.so[it], x);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (it < nitb) {
        float4 x = R.b[it];
        if (!plainA) {   // rows past the split meet an exactly-zero dY column; values only need to be finite
          const int q4 = 4 * (int)((qpack >> (8 * it)) & 255u);
          x = norm_act4(x, ld4s(nc + q4), ld4s(nc + NTr + q4), ld4s(nc + 2 * NTr + q4), slope, sigA);
        }
        store_split(bh, bl, ob.so[it], x);
      }
    }
  };

  TcPipe pp; pp.init();
  Regs r0;
  load(0, r0);
  for (; pp.kb < nkb; pp.advance(S)) {
    tc_acquire(sh, pp, S);
    const uint32_t stage = smem_s + (uint32_t)pp.s * stage_bytes;
    store(pp.kb, stage, r0);
    if (pp.kb + 1 < nkb) load(pp.kb + 1, r0);
    tc_publish_issue(sh, pp, tmem, stage, b_bytes, nkb, true, true, NT, tid);
  }
  const bool do_bias = (j0 == 0) && (G.dbias || G.dbias2);
  if (do_bias) {
    atomicAdd(&bsum[4 * aq + 0], rowsum.x); atomicAdd(&bsum[4 * aq + 1], rowsum.y);
    atomicAdd(&bsum[4 * aq + 2], rowsum.z); atomicAdd(&bsum[4 * aq + 3], rowsum.w);
  }
  mbar_wait(&sh.bar_done, 0);
  fence_after_sync();

  // ---- epilogue: atomics into dW (/ dW2), dbias ----
  const int ldo = NT + 4;
  float* ot = reinterpret_cast<float*>(smem);
  tc_drain(tmem, ot, ldo, NT, warp, lane);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols(NT));
  const bool kn = (G.w_layout == SWR_W_KN);
  const int nvalid = Jend - j0, mvalid = Nend - m0;
  auto emit = [&](int row, int col) {
    const int n = m0 + row, j = j0 + col;
    const int64_t o = kn ? ((int64_t)j * G.ldw + n) : ((int64_t)n * G.ldw + j);
    const float v = ot[(size_t)row * ldo + col];
    if (G.W2) {
      if (G.dW) atomicAdd(G.dW + o, v * __ldg(G.W2 + o));
      if (G.dW2) atomicAdd(G.dW2 + o, v * __ldg(G.W + o));
    } else if (G.dW) {
      atomicAdd(G.dW + o, v);
    }
  };
  constexpr int RPW = TC_BM / TC_WARPS;
  const bool vec_out = !kn && !G.W2 && G.dW && (nvalid % 4 == 0) && (G.ldw % 4 == 0) && is_al16(G.dW);
  if (vec_out) {   // dW[n, j], plain weight: one 16-byte reduction per 4 input features
    for (int cbase = 0; cbase < nvalid; cbase += 128) {
      const int col = cbase + 4 * lane;
      if (col < nvalid)
        for (int i = 0; i < RPW; ++i) {
          const int row = warp * RPW + i;
          if (row < mvalid) red_add_v4(G.dW + (int64_t)(m0 + row) * G.ldw + j0 + col, ld4s(ot + (size_t)row * ldo + col));
        }
    }
  } else if (!kn) {   // dW[n, j]: lanes run over j
    for (int cbase = 0; cbase < nvalid; cbase += 32) {
      const int col = cbase + lane;
      if (col < nvalid)
        for (int i = 0; i < RPW; ++i) {
          const int row = warp * RPW + i;
          if (row < mvalid) emit(row, col);
        }
    }
  } else {     // dW[j, n]: lanes run over n
    for (int col = warp; col < nvalid; col += TC_WARPS)
      for (int rb_ = 0; rb_ < TC_BM; rb_ += 32) {
        const int row = rb_ + lane;
        if (row < mvalid) emit(row, col);
      }
  }
  if (do_bias && tid < mvalid) {
    const float v = bsum[tid];
    if (G.dbias) atomicAdd(G.dbias + m0 + tid, v);
    if (G.dbias2) atomicAdd(G.dbias2 + m0 + tid, v);
  }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// accumulator columns per CTA for a width-n output: even split, multiple of 16, <= 256, and narrow enough that
// `mtiles` row tiles times the column tiles cover most of the 148 SMs
static int pick_nt(int n, int mtiles, int other_tiles) {
  int parts = ceil_div(n, 256);
  const int want = ceil_div(128 - other_tiles, mtiles > 0 ? mtiles : 1);
  if (want > parts) parts = want;
  const int max_parts = ceil_div(n, 64);   // keep tiles >= 64 wide (staging the 128-row operand dominates below that)
  if (parts > max_parts) parts = max_parts;
  if (parts < 1) parts = 1;
  int nt = round_up(ceil_div(n, parts), 16);
  if (nt > 256) nt = 256;
  return nt;
}

static constexpr size_t kTcSmemBudget = 200 * 1024;

// extra_bytes: tables behind the stages (dead once the main loop ends); keep_bytes: tables the epilogue still reads,
// placed at *epi_off behind both the stages (+ extra) and the epilogue's output tile + reduction scratch
static int pick_stages(int nt_max, size_t extra_bytes, int nkb_max, size_t* smem_bytes, size_t keep_bytes = 0, int* epi_off = nullptr,
                       bool a_in_smem = true) {
  const size_t stage = (a_in_smem ? 2 * (size_t)TC_A_BYTES : 0) + 2 * (size_t)round_up(nt_max, 64) * 128;
  int s = (int)((kTcSmemBudget - extra_bytes - keep_bytes) / stage);
  if (s > TC_MAX_STAGES) s = TC_MAX_STAGES;
  if (s > nkb_max) s = nkb_max;
  if (s < 2) s = 2;
  size_t need = (size_t)s * stage + extra_bytes;
  const size_t epi = (size_t)TC_BM * (nt_max + 4) * sizeof(float) + 2 * (size_t)TC_WARPS * nt_max * sizeof(double);   // ot + red
  if (epi > need) need = epi;
  need = (need + 15) & ~(size_t)15;
  if (epi_off) *epi_off = (int)need;
  *smem_bytes = 1024 + need + keep_bytes;
  return s;
}

// The opt-in limit is raised once per kernel to the device maximum and never lowered: a CUDA graph node keeps the
// size it was captured with, and tools that re-launch graph nodes (ncu) check it against the *current* attribute.
template <class K>
static int tc_set_smem(K kernel, size_t bytes) {
  constexpr int kMaxDyn = 227 * 1024 - 1024;   // the opt-in maximum covers static + dynamic shared memory
  if (bytes > (size_t)kMaxDyn) { set_error("fc_tc: %zu bytes of shared memory needed", bytes); return SWR_ERR_UNSUPPORTED; }
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

int launch_fc_tc_fwd(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  TcParams p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  const int mtiles = ceil_div(B, TC_BM);
  int tiles = 0, kmax = 0, nt_max = 0;
  for (int g = 0; g < n_groups; ++g) {
    p.g[g] = groups[g];
    p.nt[g] = pick_nt(groups[g].Y.n, mtiles * n_groups, 0);
    p.tile_start[g] = tiles;
    tiles += mtiles * ceil_div(groups[g].Y.n, p.nt[g]);
    kmax = max(kmax, groups[g].A.n);
    nt_max = max(nt_max, p.nt[g]);
  }
  p.tile_start[n_groups] = tiles;
  size_t smem = 0;
  p.stages = pick_stages(nt_max, 3 * sizeof(float) * (size_t)round_up(kmax, KBLK), ceil_div(kmax, KBLK), &smem, 0, nullptr, false);
  int rc = tc_set_smem(fc_tc_fwd_kernel, smem);
  if (rc) return rc;
  static const bool debug = getenv("SWR_TC_DEBUG") != nullptr;
  static long long* dbg_dev = nullptr;
  if (debug) {
    if (!dbg_dev) cudaMalloc(&dbg_dev, 40 * sizeof(long long));
    cudaMemsetAsync(dbg_dev, 0, 40 * sizeof(long long), st);
    p.dbg = dbg_dev;
  }
  fc_tc_fwd_kernel<<<tiles, TC_NT, smem, st>>>(p);
  SWR_LAUNCH_OK("fc_tc_fwd_kernel");
  if (debug) {
    long long h[40];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    for (int t = 0; t < 3; ++t)
      fprintf(stderr, "[tc_fwd dbg thr%d] nkb=%lld nt0=%d stages=%d | prologue %lld | acquire %lld store %lld load %lld publish+issue %lld | loop %lld tail-wait %lld | drain+store-out done at %lld, stats atomics done at %lld (cycles since kernel start)\n",
              t, h[8 * t + 6], p.nt[0], p.stages, h[8 * t + 7], h[8 * t + 0], h[8 * t + 1], h[8 * t + 2], h[8 * t + 3], h[8 * t + 4], h[8 * t + 5], h[24 + 4 * t], h[24 + 4 * t + 1]);
  }
  return SWR_OK;
}

// dst_group[0..n_dst]: group ranges of the destinations (prepared by launch_fc_dgrad)
int launch_fc_tc_dgrad(const FcGroup* groups, const int* dst_group_in, int n_dst, int n_groups, int64_t B, cudaStream_t st) {
  TcParams p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  int kb = 0;
  for (int g = 0; g < n_groups; ++g) { p.g[g] = groups[g]; p.tile_start[g] = kb; kb += ceil_div(groups[g].Y.n, KBLK); }
  p.tile_start[n_groups] = kb;
  const int mtiles = ceil_div(B, TC_BM);
  int dst_group[kMaxGroups + 1];
  for (int d = 0; d <= n_dst; ++d) dst_group[d] = dst_group_in[d];
  // A single plain destination (the embedding output: no norm, no activation, so its epilogue is a bare store) with a
  // long fan-in is split into a few partial fan-ins that add atomically: each CTA then stages a wider column tile
  // for a shorter contraction, and the operand rows restaged per column tile drop by about a third.
  static const bool split_fanin = !(getenv("SWR_TC_DGRAD_SPLIT") && atoi(getenv("SWR_TC_DGRAD_SPLIT")) == 0);
  if (split_fanin && n_dst == 1 && n_groups >= 2 && kb >= 16) {
    const ActDev& D = groups[0].A;
    const bool plain = D.norm.mode == SWR_NORM_NONE && D.act == SWR_ACT_NONE;
    const int ntiles_min = ceil_div(D.n, 256);
    int chunks = min(n_groups, 148 / max(1, mtiles * ntiles_min));
    if (plain && chunks >= 2) {
      if (!(groups[0].flags & FC_A_ACCUMULATE))
        SWR_CUDA_OK(cudaMemsetAsync(D.dz, 0, sizeof(float) * (size_t)B * D.ld, st));
      int d = 0, acc = 0;
      dst_group[0] = 0;
      for (int g = 0; g < n_groups; ++g) {
        acc += p.tile_start[g + 1] - p.tile_start[g];
        const int left_groups = n_groups - 1 - g, left_chunks = chunks - 1 - d;
        const int next = g + 1 < n_groups ? p.tile_start[g + 2] - p.tile_start[g + 1] : 0;
        // cut where the running length is closest to an even share
        if (left_chunks > 0 && left_groups >= left_chunks && 2 * acc * chunks + next * chunks >= 2 * kb * (d + 1)) dst_group[++d] = g + 1;
      }
      n_dst = d + 1;
      dst_group[n_dst] = n_groups;
      p.dst_atomic = (1u << n_dst) - 1u;
    }
  }
  p.n_dst = n_dst;
  for (int d = 0; d <= n_dst; ++d) p.dst_group[d] = dst_group[d];
  int tiles = 0, nkb_max = 0, nt_max = 0;
  for (int d = 0; d < n_dst; ++d) {
    const int kd = groups[dst_group[d]].A.n;
    p.nt[d] = pick_nt(kd, mtiles * n_dst, 0);
    p.dst_tile[d] = tiles;
    tiles += mtiles * ceil_div(kd, p.nt[d]);
    nkb_max = max(nkb_max, p.tile_start[dst_group[d + 1]] - p.tile_start[dst_group[d]]);
    nt_max = max(nt_max, p.nt[d]);
  }
  p.dst_tile[n_dst] = tiles;
  size_t smem = 0;
  p.stages = pick_stages(nt_max, 3 * sizeof(float) * (size_t)nkb_max * KBLK, nkb_max, &smem, 4 * sizeof(float) * (size_t)nt_max, &p.epi_off, false);
  int rc = tc_set_smem(fc_tc_dgrad_kernel, smem);
  if (rc) return rc;
  fc_tc_dgrad_kernel<<<tiles, TC_NT, smem, st>>>(p);
  SWR_LAUNCH_OK("fc_tc_dgrad_kernel");
  return SWR_OK;
}

int launch_fc_tc_wgrad(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st) {
  TcParams p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B;
  int base = 0, nt_max = 0;
  for (int g = 0; g < n_groups; ++g) {
    p.g[g] = groups[g];
    p.nt[g] = pick_nt(groups[g].A.n, 0, 128);   // even split of the input width only: the batch split supplies the CTAs
    base += ceil_div(groups[g].Y.n, TC_BM) * ceil_div(groups[g].A.n, p.nt[g]);
    nt_max = max(nt_max, p.nt[g]);
  }
  // split the batch until the grid fills the machine *without spilling into a second wave* (one CTA per SM: a
  // 149th CTA doubles the kernel's duration); keep >= 4 k-blocks (128 rows) per split
  int splits = max(1, min(base <= 148 ? 148 / base : 1, ceil_div(B, 4 * KBLK)));
  int rows = round_up(ceil_div(B, splits), KBLK);
  splits = ceil_div(B, rows);
  p.splits = splits; p.rows_per_split = rows;
  int tiles = 0;
  for (int g = 0; g < n_groups; ++g) {
    p.tile_start[g] = tiles;
    tiles += ceil_div(groups[g].Y.n, TC_BM) * ceil_div(groups[g].A.n, p.nt[g]) * splits;
  }
  p.tile_start[n_groups] = tiles;
  size_t smem = 0;
  p.stages = pick_stages(nt_max, 3 * sizeof(float) * ((size_t)round_up(nt_max, 64) + TC_BM), ceil_div(rows, KBLK), &smem);
  int rc = tc_set_smem(fc_tc_wgrad_kernel, smem);
  if (rc) return rc;
  fc_tc_wgrad_kernel<<<tiles, TC_NT, smem, st>>>(p);
  SWR_LAUNCH_OK("fc_tc_wgrad_kernel");
  return SWR_OK;
}

}  // namespace swr
