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
// activations"): the CTA's 256 threads load raw values with 16-byte read-only loads (register-prefetched
// one k-block ahead), apply the lazy BatchNorm/activation (forward), the BatchNorm-backward affine map
// (dgrad/wgrad) or the STAR weight product, split, and store straight into the 128-byte swizzled tile
// layout the MMA unit reads (swr_tc.cuh).  One elected thread issues the MMAs; tcgen05.commit arrives on
// an mbarrier per pipeline stage to hand the stage back to the stagers.  The epilogue pulls the tile out
// of TMEM with tcgen05.ld (one row per thread), transposes it through shared memory, and runs the same
// fused tails as the SIMT kernels (bias / GateNU activation / fp64 column moments; act' + BatchNorm stage-1
// sums; weight-gradient atomics) with coalesced global accesses.
#include "swr_common.cuh"
#include "swr_launch.h"
#include "swr_tc.cuh"
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
  int splits;                      // wgrad: batch splits
  int rows_per_split;
  int n_dst;                       // dgrad
  int dst_group[kMaxGroups + 1];
  int dst_tile[kMaxGroups + 1];
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
__device__ __forceinline__ uint32_t tc_setup(TcShared& sh, int stages, int nt, int tid) {
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&sh.bar_full[s], TC_NP); mbar_init(&sh.bar_free[s], 1); }
    mbar_init(&sh.bar_done, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&sh.tmem_base, tmem_cols(nt));
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  return sh.tmem_base;
}

// one pipeline stage worth of MMAs (4 k-steps x 3 split products), issued by one thread
__device__ __forceinline__ void tc_issue(uint32_t tmem, uint32_t stage_saddr, uint32_t b_bytes, bool a_mn, bool b_mn, uint32_t idesc, bool first) {
  const uint32_t ah = stage_saddr, al = ah + TC_A_BYTES, bh = al + TC_A_BYTES, bl = bh + b_bytes;
#pragma unroll
  for (int ks = 0; ks < KBLK / UMMA_K; ++ks) {
    const uint64_t dah = a_mn ? mnmajor_desc(ah, ks) : kmajor_desc(ah, ks);
    const uint64_t dal = a_mn ? mnmajor_desc(al, ks) : kmajor_desc(al, ks);
    const uint64_t dbh = b_mn ? mnmajor_desc(bh, ks) : kmajor_desc(bh, ks);
    const uint64_t dbl = b_mn ? mnmajor_desc(bl, ks) : kmajor_desc(bl, ks);
    mma_tf32(tmem, dal, dbh, idesc, (first && ks == 0) ? 0u : 1u);
    mma_tf32(tmem, dah, dbl, idesc, 1u);
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

// first n (1..3, or 4 unaligned) floats at p, the rest zero: tile edges only, kept out of line
__device__ __noinline__ float4 ld_partial(const float* __restrict__ p, int n) {
  float4 v = zero4();
  v.x = __ldg(p);
  if (n > 1) v.y = __ldg(p + 1);
  if (n > 2) v.z = __ldg(p + 2);
  if (n > 3) v.w = __ldg(p + 3);
  return v;
}
// first n (0..4) floats at p, the rest zero; one 16-byte read-only load when everything is there and aligned
__device__ __forceinline__ float4 ldn(const float* __restrict__ p, int n, bool vec) {
  float4 v = zero4();
  if (n == 4 && vec) v = __ldg(reinterpret_cast<const float4*>(p));
  else if (n > 0) v = ld_partial(p, n);
  return v;
}

// ---- what one thread moves per k-block ---------------------------------------------------------------------
// K-major operand (source rows = M/N index, contraction contiguous): chunk cj = tid & 7 of rows it*64 + (tid >> 3).
struct KSlice {
  const float* p0;      // element (row0 + r0, 4 cj) of the source
  int64_t rstride;      // 64 * ld
  uint32_t so0;         // shared-memory offset of (r0, cj); row it*64 + r0 is 8192 B further per it
  uint32_t rowmask;     // bit it: row it*64 + r0 exists (inside the tile and inside the matrix)
  int cj4;              // 4 * cj
};
__device__ __forceinline__ KSlice make_kslice(const float* base, int ld, int row0, int row_end, int tile_rows, int tid, int its) {
  KSlice s;
  const int cj = tid & 7, r0 = tid >> 3;
  s.p0 = base + (int64_t)(row0 + r0) * ld + 4 * cj;
  s.rstride = (int64_t)TC_RPI * ld;
  s.so0 = kmajor_off(r0, cj);
  s.cj4 = 4 * cj;
  s.rowmask = 0;
  for (int it = 0; it < its; ++it) {
    const int r = it * TC_RPI + r0;
    if (r < tile_rows && row0 + r < row_end) s.rowmask |= 1u << it;
  }
  return s;
}
// contraction elements [kb*32 + 4cj, +4) of every row of the slice; c_len = contraction length
template <int IT>
__device__ __forceinline__ void kslice_load(const KSlice& s, const float* __restrict__ p0, float4 (&r)[IT], int kb, int c_len, bool vec) {
  int dyn = c_len - kb * KBLK - s.cj4;
  dyn = dyn < 0 ? 0 : (dyn > 4 ? 4 : dyn);
  const float* pk = p0 + kb * KBLK;
#pragma unroll
  for (int it = 0; it < IT; ++it) r[it] = ldn(pk + it * s.rstride, ((s.rowmask >> it) & 1u) ? dyn : 0, vec);
}

// MN-major operand (source rows = contraction index, M/N contiguous): float4 v = it*512 + tid -> contraction row
// c = v / nq, M/N quad q = v % nq  (nq = quads per contraction row of the tile).
template <int IT>
struct MSlice {
  int32_t goff[IT];     // c * ld + 4 q
  uint32_t so[IT];      // shared-memory offset
  uint32_t nv;          // 4 bits per it: valid floats along M/N (0..4), 15 = outside the tile (nothing to store)
  uint32_t cpack;       // 5 bits per it: c
  uint32_t qpack;       // 8 bits per it: q
  int64_t kstride;      // 32 * ld
};
template <int IT>
__device__ __forceinline__ MSlice<IT> make_mslice(int ld, int mn0, int mn_end, int nq, int tid) {
  static_assert(IT <= 4, "packed fields hold 4 entries");
  MSlice<IT> s;
  s.nv = 0; s.cpack = 0; s.qpack = 0; s.kstride = (int64_t)32 * ld;
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int v = it * TC_NP + tid;
    const int c = v / nq, q = v - c * nq;
    uint32_t n = 15;
    if (c < KBLK) {
      int k = mn_end - (mn0 + 4 * q);
      n = (uint32_t)(k < 0 ? 0 : (k > 4 ? 4 : k));
      s.goff[it] = c * ld + 4 * q;
      s.so[it] = mnmajor_off(q, c);
      s.cpack |= (uint32_t)c << (5 * it);
      s.qpack |= (uint32_t)q << (8 * it);
    } else {
      s.goff[it] = 0; s.so[it] = 0;
    }
    s.nv |= n << (4 * it);
  }
  return s;
}
// base = source + mn0 (first M/N index of the tile) + c0 * ld (first contraction row of k-block 0); c_rows = rows left from c0
template <int IT>
__device__ __forceinline__ void mslice_load(const MSlice<IT>& s, const float* __restrict__ base, float4 (&r)[IT], int kb, int c_rows, bool vec) {
  const int left = c_rows - kb * KBLK;   // contraction rows still inside the matrix
  const float* pk = base + kb * s.kstride;
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const uint32_t nv = (s.nv >> (4 * it)) & 15u;
    if (nv != 15u) {
      const int c = (int)((s.cpack >> (5 * it)) & 31u);
      r[it] = ldn(pk + s.goff[it], c < left ? (int)nv : 0, vec);
    }
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
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  uint8_t* smem = align1024(smem_raw);
  const uint32_t smem_s = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int g = 0;
  while (g + 1 < p.n_groups && p.tile_start[g + 1] <= (int)blockIdx.x) ++g;
  const FcGroup& G = p.g[g];
  const int M = p.B, N = G.Y.n, K = G.A.n;
  const int NT = p.nt[g], NTp = (NT + 31) & ~31, nq = NTp >> 2;
  const int nt_n = (N + NT - 1) / NT;
  const int local = blockIdx.x - p.tile_start[g];
  const int m0 = (local / nt_n) * TC_BM, n0 = (local % nt_n) * NT;
  const int nkb = (K + KBLK - 1) / KBLK, Kpad = nkb * KBLK;
  const uint32_t b_bytes = (uint32_t)NTp * 128u, stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
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
  const uint32_t tmem = tc_setup(sh, S, NT, tid);
  const bool kn = (G.w_layout == SWR_W_KN);

  const int Nend = min(N, n0 + NT);
  const bool vecA = is_al16(G.A.raw) && (G.A.ld % 4 == 0);
  const bool vecW = is_al16(G.W) && (G.ldw % 4 == 0) && (!G.W2 || is_al16(G.W2));
  const bool hasW2 = G.W2 != nullptr;
  const float slope = act_slope(G.A.act);
  const bool sigA = G.A.act == SWR_ACT_SIGMOID;
  const KSlice sa = make_kslice(G.A.raw, G.A.ld, m0, M, TC_BM, tid, 2);
  KSlice sbk{};          // W [N, K] -> K-major tile
  MSlice<4> sbm{};       // W [K, N] -> MN-major tile
  const float* wbase = G.W;
  if (!kn) sbk = make_kslice(G.W, G.ldw, n0, Nend, NT, tid, 4);
  else { sbm = make_mslice<4>(G.ldw, n0, Nend, nq, tid); wbase += n0; }
  const int64_t w2diff = hasW2 ? (G.W2 - G.W) : 0;

  float4 ra[2], rb[4], rb2[4];
  auto load = [&](int kb) {
    kslice_load<2>(sa, sa.p0, ra, kb, K, vecA);
    if (!kn) {
      kslice_load<4>(sbk, sbk.p0, rb, kb, K, vecW);
      if (hasW2) kslice_load<4>(sbk, sbk.p0 + w2diff, rb2, kb, K, vecW);
    } else {
      mslice_load<4>(sbm, wbase, rb, kb, K, vecW);
      if (hasW2) mslice_load<4>(sbm, wbase + w2diff, rb2, kb, K, vecW);
    }
  };
  auto store = [&](int kb, uint32_t stage) {
    const uint32_t ah = stage, al = ah + TC_A_BYTES, bh = al + TC_A_BYTES, bl = bh + b_bytes;
    if (plainA) {
#pragma unroll
      for (int it = 0; it < 2; ++it) store_split(ah, al, sa.so0 + it * (TC_RPI * 128), ra[it]);
    } else {   // padded k: coefficients are 0 -> act(0) stays finite and meets a zero weight
      const int k = kb * KBLK + sa.cj4;
      const float4 mu = ld4s(kc + k), sc = ld4s(kc + Kpad + k), bb = ld4s(kc + 2 * Kpad + k);
#pragma unroll
      for (int it = 0; it < 2; ++it) store_split(ah, al, sa.so0 + it * (TC_RPI * 128), norm_act4(ra[it], mu, sc, bb, slope, sigA));
    }
    if (!kn) {
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (it * TC_RPI + (tid >> 3) < NT) store_split(bh, bl, sbk.so0 + it * (TC_RPI * 128), hasW2 ? mul4(rb[it], rb2[it]) : rb[it]);
    } else {
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (((sbm.nv >> (4 * it)) & 15u) != 15u) store_split(bh, bl, sbm.so[it], hasW2 ? mul4(rb[it], rb2[it]) : rb[it]);
    }
  };

  TcPipe pp; pp.init();
  load(0);
  for (; pp.kb < nkb; pp.advance(S)) {
    tc_acquire(sh, pp, S);
    const uint32_t stage = smem_s + (uint32_t)pp.s * stage_bytes;
    store(pp.kb, stage);
    if (pp.kb + 1 < nkb) load(pp.kb + 1);
    tc_publish_issue(sh, pp, tmem, stage, b_bytes, nkb, false, kn, NT, tid);
  }
  mbar_wait(&sh.bar_done, 0);
  fence_after_sync();

  // ---- epilogue: bias, optional activation, store raw Y, fp64 column moments ----
  const int ldo = NT + 4;
  float* ot = reinterpret_cast<float*>(smem);
  double* red = reinterpret_cast<double*>(ot + (size_t)TC_BM * ldo);
  tc_drain(tmem, ot, ldo, NT, warp, lane);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols(NT));
  float* Y = const_cast<float*>(G.Y.raw);
  const int nvalid = Nend - n0;
  constexpr int RPW = TC_BM / TC_WARPS;   // accumulator rows per warp
  for (int cbase = 0; cbase < nvalid; cbase += 32) {
    const int col = cbase + lane, n = n0 + col;
    if (col < nvalid) {
      const float bias = ld_opt(G.bias, n, 0.f) + ld_opt(G.bias2, n, 0.f);
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int row = warp * RPW + i, m = m0 + row;
        if (m < M) {
          float y = ot[(size_t)row * ldo + col] + bias;
          if (G.e_act != SWR_ACT_NONE) y = act_fwd(y, G.e_act) * G.e_scale;
          Y[(int64_t)m * G.Y.ld + n] = y;
          s1 += (double)y; s2 += (double)y * (double)y;
        }
      }
      red[(0 * TC_WARPS + warp) * NT + col] = s1;
      red[(1 * TC_WARPS + warp) * NT + col] = s2;
    }
  }
  if (G.stats_out) {
    __syncthreads();
    tc_col_atomics(red, G.stats_out, n0, nvalid, NT, tid);
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
  const int NT = p.nt[d], NTp = (NT + 31) & ~31, nq = NTp >> 2;
  const int nt_n = (Kd + NT - 1) / NT;
  const int local = blockIdx.x - p.dst_tile[d];
  const int m0 = (local / nt_n) * TC_BM, j0 = (local % nt_n) * NT;
  const int kb0 = p.tile_start[gs];
  const int nkb = p.tile_start[ge] - kb0, Kc = nkb * KBLK;
  const uint32_t b_bytes = (uint32_t)NTp * 128u, stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
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
  const uint32_t tmem = tc_setup(sh, S, NT, tid);

  const int Jend = min(Kd, j0 + NT);
  float4 ra[2], rr[2], rb[4], rb2[4];
  // per-group staging state (rebuilt when the k-block walk enters the next group of the fan-in)
  int cur_g = gs - 1, g_kb0 = 0, g_N = 0;
  bool kn = false, hasW2 = false, need_raw = false, vecY = false, vecW = false;
  KSlice sa{}, sbk{};
  MSlice<4> sbm{};
  const float* wbase = nullptr;
  int64_t w2diff = 0, rawdiff = 0;
  auto enter_group = [&](int g) {
    const FcGroup& G = p.g[g];
    cur_g = g; g_kb0 = p.tile_start[g] - kb0; g_N = G.Y.n;
    kn = (G.w_layout == SWR_W_KN); hasW2 = G.W2 != nullptr;
    need_raw = (G.Y.norm.mode == SWR_NORM_BATCH);
    vecY = is_al16(G.Y.dz) && is_al16(G.Y.raw) && (G.Y.ld % 4 == 0);
    vecW = is_al16(G.W) && (G.ldw % 4 == 0) && (!G.W2 || is_al16(G.W2));
    sa = make_kslice(G.Y.dz, G.Y.ld, m0, M, TC_BM, tid, 2);
    rawdiff = G.Y.raw - G.Y.dz;
    w2diff = hasW2 ? (G.W2 - G.W) : 0;
    if (!kn) { sbm = make_mslice<4>(G.ldw, j0, Jend, nq, tid); wbase = G.W + j0; }   // W[n, j] -> MN-major
    else sbk = make_kslice(G.W, G.ldw, j0, Jend, NT, tid, 4);                          // W[j, n] -> K-major
  };
  auto load = [&](int kb) {
    while (cur_g < gs || (cur_g + 1 < ge && p.tile_start[cur_g + 1] - kb0 <= kb)) enter_group(cur_g + 1);
    const int lkb = kb - g_kb0;   // k-block inside the group
    kslice_load<2>(sa, sa.p0, ra, lkb, g_N, vecY);
    if (need_raw) kslice_load<2>(sa, sa.p0 + rawdiff, rr, lkb, g_N, vecY);
    else { rr[0] = zero4(); rr[1] = zero4(); }
    if (!kn) {
      mslice_load<4>(sbm, wbase, rb, lkb, g_N, vecW);
      if (hasW2) mslice_load<4>(sbm, wbase + w2diff, rb2, lkb, g_N, vecW);
    } else {
      kslice_load<4>(sbk, sbk.p0, rb, lkb, g_N, vecW);
      if (hasW2) kslice_load<4>(sbk, sbk.p0 + w2diff, rb2, lkb, g_N, vecW);
    }
  };
  // called before load(kb + 1): the slices and layout flags still describe the group of block kb
  auto store = [&](int kb, uint32_t stage) {
    const uint32_t ah = stage, al = ah + TC_A_BYTES, bh = al + TC_A_BYTES, bl = bh + b_bytes;
    const int k = kb * KBLK + sa.cj4;
    const float4 c0 = ld4s(dc + k), c1 = ld4s(dc + Kc + k), c2 = ld4s(dc + 2 * Kc + k);
#pragma unroll
    for (int it = 0; it < 2; ++it)   // rows outside the batch only feed accumulator rows that are never stored
      store_split(ah, al, sa.so0 + it * (TC_RPI * 128), affine4(c0, ra[it], c1, rr[it], c2));
    if (!kn) {
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (((sbm.nv >> (4 * it)) & 15u) != 15u) store_split(bh, bl, sbm.so[it], hasW2 ? mul4(rb[it], rb2[it]) : rb[it]);
    } else {
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (it * TC_RPI + (tid >> 3) < NT) store_split(bh, bl, sbk.so0 + it * (TC_RPI * 128), hasW2 ? mul4(rb[it], rb2[it]) : rb[it]);
    }
  };

  TcPipe pp; pp.init();
  load(0);
  for (; pp.kb < nkb; pp.advance(S)) {
    tc_acquire(sh, pp, S);
    const uint32_t stage = smem_s + (uint32_t)pp.s * stage_bytes;
    const bool b_mn = !kn;   // W[n, j] tiles are MN-major
    store(pp.kb, stage);
    if (pp.kb + 1 < nkb) load(pp.kb + 1);
    tc_publish_issue(sh, pp, tmem, stage, b_bytes, nkb, false, b_mn, NT, tid);
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
  if (warp == 0) tmem_dealloc(tmem, tmem_cols(NT));
  const bool accumulate = (p.g[gs].flags & FC_A_ACCUMULATE) != 0;
  const bool has_norm = D.norm.mode != SWR_NORM_NONE;
  const bool plainD = !has_norm && D.act == SWR_ACT_NONE;
  const int nvalid = Jend - j0;
  constexpr int RPW = TC_BM / TC_WARPS;
  for (int cbase = 0; cbase < nvalid; cbase += 32) {
    const int col = cbase + lane, j = j0 + col;
    if (col < nvalid) {
      ColCoef cc = {0.f, 1.f, 0.f, 1.f};
      if (!plainD) cc = col_coef(D.norm, j, p.inv_count);
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int row = warp * RPW + i, m = m0 + row;
        if (m < M) {
          const int64_t o = (int64_t)m * D.ld + j;
          float dz = ot[(size_t)row * ldo + col];
          if (!plainD) {
            const float raw = D.raw[o];
            dz *= act_grad(fmaf(raw - cc.mu, cc.s, cc.b), D.act);
            s1 += (double)dz; s2 += (double)dz * (double)((raw - cc.mu) * cc.r);
          }
          if (accumulate) dz += D.dz[o];
          D.dz[o] = dz;
        }
      }
      red[(0 * TC_WARPS + warp) * NT + col] = s1;
      red[(1 * TC_WARPS + warp) * NT + col] = s2;
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
  const int NT = p.nt[g], NTp = (NT + 31) & ~31, nq = NTp >> 2;
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
  const uint32_t b_bytes = (uint32_t)NTp * 128u, stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
  const int S = p.stages;
  float* nc = reinterpret_cast<float*>(smem + (size_t)S * stage_bytes);   // [3][NTp]: mu, s, b of the input columns j

  const bool plainA = (G.A.norm.mode == SWR_NORM_NONE && G.A.act == SWR_ACT_NONE);
  for (int i = tid; i < NTp; i += TC_NT) {
    ColCoef c = {0.f, 0.f, 0.f, 0.f};
    if (j0 + i < K) c = plainA ? ColCoef{0.f, 1.f, 0.f, 1.f} : col_coef(G.A.norm, j0 + i, p.inv_count);
    nc[i] = c.mu; nc[NTp + i] = c.s; nc[2 * NTp + i] = c.b;
  }
  if (tid < TC_BM) bsum[tid] = 0.f;
  const uint32_t tmem = tc_setup(sh, S, NT, tid);

  // A(n, b) = dY[b, n]: n-contiguous -> MN-major tile; this thread always stages the same 4 output features
  // (quad aq of contraction rows it*16 + tid/32), so their dY coefficients live in registers
  const int aq = tid & 31;
  const int Nend = min(N, m0 + TC_BM), Jend = min(K, j0 + NT);
  float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (m0 + 4 * aq + i < N) { const DyCoef c = dy_coef(G.Y, m0 + 4 * aq + i, p.inv_count); f0[i] = c.c0; f1[i] = c.c1; f2[i] = c.c2; }
  const float4 c0 = make_float4(f0[0], f0[1], f0[2], f0[3]), c1 = make_float4(f1[0], f1[1], f1[2], f1[3]), c2 = make_float4(f2[0], f2[1], f2[2], f2[3]);
  const bool vecY = is_al16(G.Y.dz) && is_al16(G.Y.raw) && (G.Y.ld % 4 == 0);
  const bool vecA = is_al16(G.A.raw) && (G.A.ld % 4 == 0);
  const bool need_raw = (G.Y.norm.mode == SWR_NORM_BATCH);
  const float slope = act_slope(G.A.act);
  const bool sigA = G.A.act == SWR_ACT_SIGMOID;
  const MSlice<2> sa = make_mslice<2>(G.Y.ld, m0, Nend, 32, tid);
  const MSlice<4> sb = make_mslice<4>(G.A.ld, j0, Jend, nq, tid);
  const float* dzbase = G.Y.dz + (int64_t)b_begin * G.Y.ld + m0;
  const float* rawbase = G.Y.raw + (int64_t)b_begin * G.Y.ld + m0;
  const float* abase = G.A.raw + (int64_t)b_begin * G.A.ld + j0;

  float4 ra[2], rr[2], rb[4];
  float4 rowsum = zero4();
  auto load = [&](int kb) {
    mslice_load<2>(sa, dzbase, ra, kb, rows, vecY);
    if (need_raw) mslice_load<2>(sa, rawbase, rr, kb, rows, vecY);
    else { rr[0] = zero4(); rr[1] = zero4(); }
    mslice_load<4>(sb, abase, rb, kb, rows, vecA);
  };
  auto store = [&](int kb, uint32_t stage) {
    const uint32_t ah = stage, al = ah + TC_A_BYTES, bh = al + TC_A_BYTES, bl = bh + b_bytes;
    const int left = rows - kb * KBLK;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c = it * (TC_NP / 32) + (tid >> 5);
      float4 x = zero4();
      if (c < left) x = affine4(c0, ra[it], c1, rr[it], c2);   // c2 != 0: contraction padding must stay exactly zero
      rowsum.x += x.x; rowsum.y += x.y; rowsum.z += x.z; rowsum.w += x.w;
      store_split(ah, al, sa.so[it], x);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (((sb.nv >> (4 * it)) & 15u) != 15u) {
        float4 x = rb[it];
        if (!plainA) {   // rows past the split meet an exactly-zero dY column; values only need to be finite
          const int q4 = 4 * (int)((sb.qpack >> (8 * it)) & 255u);
          x = norm_act4(x, ld4s(nc + q4), ld4s(nc + NTp + q4), ld4s(nc + 2 * NTp + q4), slope, sigA);
        }
        store_split(bh, bl, sb.so[it], x);
      }
    }
  };

  TcPipe pp; pp.init();
  load(0);
  for (; pp.kb < nkb; pp.advance(S)) {
    tc_acquire(sh, pp, S);
    const uint32_t stage = smem_s + (uint32_t)pp.s * stage_bytes;
    store(pp.kb, stage);
    if (pp.kb + 1 < nkb) load(pp.kb + 1);
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
  if (!kn) {   // dW[n, j]: lanes run over j
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

static int pick_stages(int nt_max, size_t extra_bytes, int nkb_max, size_t* smem_bytes) {
  const size_t stage = 2 * (size_t)TC_A_BYTES + 2 * (size_t)round_up(nt_max, 32) * 128;
  int s = (int)((kTcSmemBudget - extra_bytes) / stage);
  if (s > TC_MAX_STAGES) s = TC_MAX_STAGES;
  if (s > nkb_max) s = nkb_max;
  if (s < 2) s = 2;
  size_t need = (size_t)s * stage + extra_bytes;
  const size_t epi = (size_t)TC_BM * (nt_max + 4) * sizeof(float) + 2 * (size_t)TC_WARPS * nt_max * sizeof(double);   // ot + red
  if (epi > need) need = epi;
  *smem_bytes = 1024 + need;
  return s;
}

template <class K>
static int tc_set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) { set_error("fc_tc: %zu bytes of shared memory needed", bytes); return SWR_ERR_UNSUPPORTED; }
  SWR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SWR_OK;
}

// 0: never, 1: whenever the shapes allow, 2 (default): when the launch is big enough to pay for the pipeline
static int tc_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SWR_FC_TC");
    mode = e ? atoi(e) : 2;
  }
  return mode;
}
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
  p.stages = pick_stages(nt_max, 3 * sizeof(float) * (size_t)round_up(kmax, KBLK), ceil_div(kmax, KBLK), &smem);
  int rc = tc_set_smem(fc_tc_fwd_kernel, smem);
  if (rc) return rc;
  fc_tc_fwd_kernel<<<tiles, TC_NT, smem, st>>>(p);
  SWR_LAUNCH_OK("fc_tc_fwd_kernel");
  return SWR_OK;
}

// p.n_dst / p.dst_group prepared by the caller (launch_fc_dgrad)
int launch_fc_tc_dgrad(const FcGroup* groups, const int* dst_group, int n_dst, int n_groups, int64_t B, cudaStream_t st) {
  TcParams p{};
  p.n_groups = n_groups; p.B = (int)B; p.inv_count = 1.0f / (float)B; p.n_dst = n_dst;
  int kb = 0;
  for (int g = 0; g < n_groups; ++g) { p.g[g] = groups[g]; p.tile_start[g] = kb; kb += ceil_div(groups[g].Y.n, KBLK); }
  p.tile_start[n_groups] = kb;
  for (int d = 0; d <= n_dst; ++d) p.dst_group[d] = dst_group[d];
  const int mtiles = ceil_div(B, TC_BM);
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
  p.stages = pick_stages(nt_max, 3 * sizeof(float) * (size_t)nkb_max * KBLK, nkb_max, &smem);
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
  // split the batch until the grid covers the machine once; keep >= 4 k-blocks (128 rows) per split
  int splits = max(1, min(ceil_div(148, base), ceil_div(B, 4 * KBLK)));
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
  p.stages = pick_stages(nt_max, 3 * sizeof(float) * (size_t)round_up(nt_max, 32), ceil_div(rows, KBLK), &smem);
  int rc = tc_set_smem(fc_tc_wgrad_kernel, smem);
  if (rc) return rc;
  fc_tc_wgrad_kernel<<<tiles, TC_NT, smem, st>>>(p);
  SWR_LAUNCH_OK("fc_tc_wgrad_kernel");
  return SWR_OK;
}

}  // namespace swr
