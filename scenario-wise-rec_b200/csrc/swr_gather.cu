// swr_gather.cu -- K1 fused multi-field embedding gather, K2 embedding-gradient scatter,
// and the column-moment kernel used by STAR's partitioned norm.
//
// Both K1 and K2 are HBM-bound byte movers (DESIGN.md "K1/K2"): the index columns of a
// row tile are staged through shared memory once (coalesced, any int width), turned into
// bounds-checked row offsets, and every thread then moves 16-byte pieces: 4 lanes cover
// one 64 B embedding row, a warp writes 512 contiguous bytes of the dense output.
#include "swr_common.cuh"
#include "swr_launch.h"

namespace swr {

constexpr int kMaxFields = 64;

struct SparseField {
  const float* table;   // K1: [vocab, E] weights          K2: unused
  float* gtable;        // K2: [vocab, E] dense gradient   K1: unused
  const void* idx;      // [B] any integer dtype
  int64_t vocab;
  int dtype;
  int E;
  int col;              // first output column of this field
  int world;            // > 0: the table is row-sharded over `world` ranks: row r lives on rank r % world at local row r / world,
  float* const* peers;  //      peers[rank] = base of that rank's shard (K1: weights, K2: gradient); reached over NVLink
};
struct DenseField { const void* ptr; int dtype; int col; };

struct GatherParams {
  SparseField sp[kMaxFields];
  DenseField de[kMaxFields];
  float* out;
  int64_t ld;
  int64_t B;
  int n_sparse, n_dense;
  int S4;               // float4 columns of the sparse part (vector path)
  int S;                // float columns of the sparse part
  int* oob;
};

struct ScatterParams {
  SparseField sp[kMaxFields];
  const float* g;
  int64_t ld;
  int64_t B;
  int n_sparse;
};

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  // sm_90+: one 16-byte reduction instead of four 4-byte ones
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// stage the row offsets (idx * E, or -1 when out of range) of TB rows x n_sparse fields
template <int TB, class P>
__device__ __forceinline__ void stage_offsets(const P& p, int64_t b0, int nb, int64_t* rowoff, int* oob) {
  for (int i = threadIdx.x; i < p.n_sparse * TB; i += blockDim.x) {
    const int f = i / TB, bl = i - f * TB;
    int64_t off = -1;
    if (bl < nb) {
      const int64_t ix = load_index(p.sp[f].idx, p.sp[f].dtype, b0 + bl);
      if (ix >= 0 && ix < p.sp[f].vocab) {
        if (p.sp[f].world > 0) {      // absolute address (in floats) of the row in its owner's shard
          const int w = p.sp[f].world;
          off = (int64_t)(reinterpret_cast<uintptr_t>(p.sp[f].peers[ix % w]) >> 2) + (ix / w) * p.sp[f].E;
        } else {
          off = ix * p.sp[f].E;
        }
      } else if (oob) { oob[0] = 1; oob[1] = f; }
    }
    rowoff[i] = off;
  }
}

// ---------------------------------------------------------------------------------------
// K1, vector path: every E_f and ld are multiples of 4 floats, bases 16-byte aligned.
// ---------------------------------------------------------------------------------------
template <int TB>
__global__ void __launch_bounds__(256) gather_vec_kernel(const __grid_constant__ GatherParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  int64_t* rowoff = reinterpret_cast<int64_t*>(smem);                       // [n_sparse][TB]
  unsigned short* cmap = reinterpret_cast<unsigned short*>(rowoff + p.n_sparse * TB);  // [S4] field<<8 | q
  const int64_t b0 = (int64_t)blockIdx.x * TB;
  const int nb = (int)min((int64_t)TB, p.B - b0);

  stage_offsets<TB>(p, b0, nb, rowoff, p.oob);
  for (int c4 = threadIdx.x; c4 < p.S4; c4 += blockDim.x) {
    int f = 0;
    while (f + 1 < p.n_sparse && (p.sp[f + 1].col >> 2) <= c4) ++f;
    cmap[c4] = (unsigned short)((f << 8) | (c4 - (p.sp[f].col >> 2)));
  }
  __syncthreads();

  const int total = nb * p.S4;
#pragma unroll 4
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int bl = i / p.S4, c4 = i - bl * p.S4;
    const int f = cmap[c4] >> 8, q = cmap[c4] & 0xff;
    const int64_t off = rowoff[f * TB + bl];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (off >= 0) v = ldg_stream(reinterpret_cast<const float4*>((p.sp[f].world > 0 ? static_cast<const float*>(nullptr) : p.sp[f].table) + off) + q);
    *reinterpret_cast<float4*>(p.out + (b0 + bl) * p.ld + 4 * c4) = v;
  }
  // dense scalars: thread -> (j, bl) with bl fastest so the column reads coalesce
  for (int i = threadIdx.x; i < p.n_dense * nb; i += blockDim.x) {
    const int j = i / nb, bl = i - j * nb;
    p.out[(b0 + bl) * p.ld + p.de[j].col] = load_scalar(p.de[j].ptr, p.de[j].dtype, b0 + bl);
  }
}

// K1, scalar path: arbitrary E_f / alignment (e.g. embed_dim not a multiple of 4).
template <int TB>
__global__ void __launch_bounds__(256) gather_scalar_kernel(const __grid_constant__ GatherParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  int64_t* rowoff = reinterpret_cast<int64_t*>(smem);
  unsigned short* cfield = reinterpret_cast<unsigned short*>(rowoff + p.n_sparse * TB);  // [S]
  const int64_t b0 = (int64_t)blockIdx.x * TB;
  const int nb = (int)min((int64_t)TB, p.B - b0);
  stage_offsets<TB>(p, b0, nb, rowoff, p.oob);
  for (int c = threadIdx.x; c < p.S; c += blockDim.x) {
    int f = 0;
    while (f + 1 < p.n_sparse && p.sp[f + 1].col <= c) ++f;
    cfield[c] = (unsigned short)f;
  }
  __syncthreads();
  const int total = nb * p.S;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int bl = i / p.S, c = i - bl * p.S;
    const int f = cfield[c];
    const int64_t off = rowoff[f * TB + bl];
    p.out[(b0 + bl) * p.ld + c] = off >= 0 ? __ldg((p.sp[f].world > 0 ? static_cast<const float*>(nullptr) : p.sp[f].table) + off + (c - p.sp[f].col)) : 0.f;
  }
  for (int i = threadIdx.x; i < p.n_dense * nb; i += blockDim.x) {
    const int j = i / nb, bl = i - j * nb;
    p.out[(b0 + bl) * p.ld + p.de[j].col] = load_scalar(p.de[j].ptr, p.de[j].dtype, b0 + bl);
  }
}

// ---------------------------------------------------------------------------------------
// K2: scatter-add of the [B, ld] input gradient into the dense per-table gradients.
//  * tiny vocabularies (vocab*E <= kPrivFloats): the CTA accumulates a private copy of
//    the whole table in shared memory and flushes it once (histogram privatisation);
//  * everything else: 16-byte red.global.add.v4.f32, duplicates inside a warp combined
//    first (match.any on the destination address, leader sums its peers from shared memory).
// ---------------------------------------------------------------------------------------
constexpr int kPrivFloats = 512;   // vocab <= 32 at E = 16
constexpr int kScatterTB = 32;

__global__ void __launch_bounds__(256) scatter_kernel(const __grid_constant__ ScatterParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int TB = kScatterTB;
  int64_t* rowoff = reinterpret_cast<int64_t*>(smem);                   // [n_sparse][TB]
  float4* stage = reinterpret_cast<float4*>(rowoff + p.n_sparse * TB);  // [256] warp staging
  float* priv = reinterpret_cast<float*>(stage + 256);                  // [kPrivFloats]
  const int64_t b0 = (int64_t)blockIdx.x * TB;
  const int nb = (int)min((int64_t)TB, p.B - b0);
  const int lane = threadIdx.x & 31;
  stage_offsets<TB>(p, b0, nb, rowoff, nullptr);
  __syncthreads();

  for (int f = 0; f < p.n_sparse; ++f) {
    const SparseField& sf = p.sp[f];
    const int E = sf.E;
    const int64_t tbl = sf.vocab * E;
    float* const gbase = sf.world > 0 ? static_cast<float*>(nullptr) : sf.gtable;   // sharded: rowoff holds absolute addresses
    if (tbl <= kPrivFloats && sf.world == 0) {
      for (int i = threadIdx.x; i < (int)tbl; i += blockDim.x) priv[i] = 0.f;
      __syncthreads();
      for (int i = threadIdx.x; i < nb * E; i += blockDim.x) {
        const int bl = i / E, e = i - bl * E;
        const int64_t off = rowoff[f * TB + bl];
        if (off >= 0) atomicAdd(&priv[off + e], p.g[(b0 + bl) * p.ld + sf.col + e]);
      }
      __syncthreads();
      for (int i = threadIdx.x; i < (int)tbl; i += blockDim.x) {
        const float v = priv[i];
        if (v != 0.f) atomicAdd(sf.gtable + i, v);
      }
      __syncthreads();
    } else if ((E & 3) == 0 && (sf.col & 3) == 0 && (p.ld & 3) == 0) {
      const int E4 = E >> 2;
      const int total = nb * E4;
      // warp-uniform trip count so the match/shuffle below stay convergent
      for (int base = (threadIdx.x & ~31); base < total; base += blockDim.x) {
        const int i = base + lane;
        const bool valid = i < total;
        int64_t dst = -1 - lane;            // unique key for idle lanes
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          const int bl = i / E4, q = i - bl * E4;
          const int64_t off = rowoff[f * TB + bl];
          if (off >= 0) {
            dst = off + 4 * q;
            v = *reinterpret_cast<const float4*>(p.g + (b0 + bl) * p.ld + sf.col + 4 * q);
          }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, dst);
        const bool dup = __any_sync(0xffffffffu, peers != (1u << lane));
        if (dup) {
          stage[threadIdx.x] = v;
          __syncwarp();
          if (dst >= 0 && (peers & ((1u << lane) - 1)) == 0) {       // lowest lane of its group leads
            unsigned rest = peers & ~(1u << lane);
            while (rest) {
              const int src = __ffs(rest) - 1;
              rest &= rest - 1;
              const float4 o = stage[(threadIdx.x & ~31) + src];
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            red_add_v4(gbase + dst, v);
          }
          __syncwarp();
        } else if (dst >= 0) {
          red_add_v4(gbase + dst, v);
        }
      }
    } else {
      for (int i = threadIdx.x; i < nb * E; i += blockDim.x) {
        const int bl = i / E, e = i - bl * E;
        const int64_t off = rowoff[f * TB + bl];
        if (off >= 0) atomicAdd(gbase + off + e, p.g[(b0 + bl) * p.ld + sf.col + e]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// column moments of a plain [B, n] buffer: stats[c] = (sum_b x, sum_b x^2) in fp64
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ x, int64_t B, int n, int64_t ld,
                                                      double* __restrict__ stats, int rows_per_cta) {
  __shared__ double s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(B, r0 + rows_per_cta);
  double a = 0.0, b = 0.0;
  if (c < n)
    for (int64_t r = r0 + ry; r < r1; r += 8) { const double v = (double)x[r * ld + c]; a += v; b += v * v; }
  s1[ry][cx] = a; s2[ry][cx] = b;
  __syncthreads();
  if (ry == 0 && c < n) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a += s1[k][cx]; b += s2[k][cx]; }
    atomicAdd(stats + 2 * c, a);
    atomicAdd(stats + 2 * c + 1, b);
  }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_gather(const GatherLaunch& g, cudaStream_t st) {
  if (g.B <= 0) return SWR_OK;
  if (g.n_sparse < 0 || g.n_dense < 0 || g.n_sparse + g.n_dense == 0) { set_error("gather: empty feature list"); return SWR_ERR_INVALID; }
  if (g.n_sparse > kMaxFields || g.n_dense > kMaxFields) { set_error("gather: more than %d fields per launch", kMaxFields); return SWR_ERR_UNSUPPORTED; }
  GatherParams p{};
  bool vec = (g.ld % 4 == 0) && aligned16(g.out);
  int col = 0;
  for (int f = 0; f < g.n_sparse; ++f) {
    p.sp[f].table = g.tables[f]; p.sp[f].gtable = nullptr; p.sp[f].idx = g.idx[f]; p.sp[f].vocab = g.vocab[f];
    p.sp[f].dtype = g.idx_dtype[f]; p.sp[f].E = g.E[f]; p.sp[f].col = col;
    p.sp[f].world = g.world ? g.world[f] : 0; p.sp[f].peers = g.peers ? const_cast<float* const*>(g.peers[f]) : nullptr;
    if (p.sp[f].world > 0 && !p.sp[f].peers) { set_error("gather: sharded field %d without peer table", f); return SWR_ERR_INVALID; }
    if (g.E[f] <= 0 || g.E[f] > 1020) { set_error("gather: embed_dim %d unsupported", g.E[f]); return SWR_ERR_UNSUPPORTED; }
    vec = vec && (g.E[f] % 4 == 0) && (p.sp[f].world > 0 || aligned16(g.tables[f]));
    col += g.E[f];
  }
  p.S = col; p.S4 = col / 4;
  for (int j = 0; j < g.n_dense; ++j) { p.de[j].ptr = g.dense[j]; p.de[j].dtype = g.dense_dtype[j]; p.de[j].col = col + j; }
  if (col + g.n_dense > g.ld) { set_error("gather: ld_out %lld < %d columns", (long long)g.ld, col + g.n_dense); return SWR_ERR_INVALID; }
  p.out = g.out; p.ld = g.ld; p.B = g.B; p.n_sparse = g.n_sparse; p.n_dense = g.n_dense; p.oob = g.oob;
  // row tile: 16 rows while the batch is small enough that 32-row tiles leave SMs idle
  const bool small = g.B < 32 * 148 * 2;
  const int TB = small ? 16 : 32;
  const int grid = ceil_div(g.B, TB);
  if (vec) {
    const size_t sm = (size_t)g.n_sparse * TB * 8 + (size_t)p.S4 * 2 + 16;
    if (small) gather_vec_kernel<16><<<grid, 256, sm, st>>>(p); else gather_vec_kernel<32><<<grid, 256, sm, st>>>(p);
  } else {
    const size_t sm = (size_t)g.n_sparse * TB * 8 + (size_t)p.S * 2 + 16;
    if (small) gather_scalar_kernel<16><<<grid, 256, sm, st>>>(p); else gather_scalar_kernel<32><<<grid, 256, sm, st>>>(p);
  }
  SWR_LAUNCH_OK("gather_kernel");
  return SWR_OK;
}

int launch_scatter(const ScatterLaunch& s, cudaStream_t st) {
  if (s.B <= 0 || s.n_sparse == 0) return SWR_OK;
  if (s.n_sparse > kMaxFields) { set_error("scatter: more than %d fields per launch", kMaxFields); return SWR_ERR_UNSUPPORTED; }
  ScatterParams p{};
  int col = 0;
  for (int f = 0; f < s.n_sparse; ++f) {
    p.sp[f].table = nullptr; p.sp[f].gtable = s.gtables[f]; p.sp[f].idx = s.idx[f]; p.sp[f].vocab = s.vocab[f];
    p.sp[f].dtype = s.idx_dtype[f]; p.sp[f].E = s.E[f]; p.sp[f].col = s.col ? s.col[f] : col;
    p.sp[f].world = s.world ? s.world[f] : 0; p.sp[f].peers = s.peers ? s.peers[f] : nullptr;
    if (p.sp[f].world > 0 && !p.sp[f].peers) { set_error("scatter: sharded field %d without peer table", f); return SWR_ERR_INVALID; }
    col += s.E[f];
  }
  p.g = s.g; p.ld = s.ld; p.B = s.B; p.n_sparse = s.n_sparse;
  const size_t sm = (size_t)s.n_sparse * kScatterTB * 8 + 256 * 16 + kPrivFloats * 4;
  scatter_kernel<<<ceil_div(s.B, kScatterTB), 256, sm, st>>>(p);
  SWR_LAUNCH_OK("scatter_kernel");
  return SWR_OK;
}

int launch_colstats(const float* x, int64_t B, int n, int64_t ld, double* stats, cudaStream_t st) {
  if (B <= 0 || n <= 0) return SWR_OK;
  const int rows = 256;
  dim3 grid(ceil_div(n, 32), ceil_div(B, rows));
  colstats_kernel<<<grid, 256, 0, st>>>(x, B, n, ld, stats, rows);
  SWR_LAUNCH_OK("colstats_kernel");
  return SWR_OK;
}

}  // namespace swr
