// swr_tc.cuh -- sm_100a tensor-core plumbing for the grouped FC kernels (swr_fc_tc.cu):
// mbarrier, TMEM allocation, tcgen05.mma (kind::tf32) issue, tcgen05.ld, and the shared-memory
// operand layouts + matrix descriptors the MMA unit reads.
//
// Weight tiles come through TMA (cp.async.bulk.tensor, SWIZZLE_128B) from images the presplit kernel wrote
// (effective weight W (.) W2, hi/lo TF32 split, both orientations).  Activation-side operands are *computed*
// while staging (lazy BatchNorm + activation, BatchNorm-backward affine map, hi/lo split) by stager warps and go
// registers -> TMEM (tcgen05.st) or, for the weight-gradient's column operand, to shared memory in the same
// canonical 128-byte swizzled layouts TMA produces:
//   K-major  (contraction contiguous): row r (an M or N index) owns 128 B = 32 fp32 of contraction;
//            16-byte chunk j of row r sits at r*128 + ((j ^ (r & 7)) << 4); 8 rows = one 1024 B atom.
//            (SWIZZLE_128B, 16-byte base)
//   MN-major (M/N index contiguous):   32-bit operands use the 128-byte swizzle with a 32-byte base
//            (SWIZZLE_128B_BASE32B): a 128 B row holds 32 consecutive M/N indices of one contraction
//            index c, rows of a 32-wide M/N group are contiguous over c ([M/N group][32 c][128 B]), and the
//            32-byte unit u of row c sits at unit (u ^ (c & 3)).  A swizzle atom is 4 rows = 512 B.
// One tcgen05.mma of kind::tf32 contracts 8 elements: 32 B of a K-major row, or 8 rows (1024 B)
// of an MN-major group.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swr {
namespace tc {

constexpr int KBLK = 32;            // contraction elements per pipeline stage (= one 128 B swizzle row)
constexpr int UMMA_K = 8;           // contraction elements per tcgen05.mma.kind::tf32

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  // the suspend-time hint lets the hardware park the warp until the phase completes instead of polling
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {   // one arrival + `bytes` of pending transactions
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a descriptor or protocol bug must surface as a trapped kernel (an error the host
// reports), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// exactly one lane of a converged warp returns true; the compiler knows the guarded region has a single active thread,
// so instructions with warp-uniform operands (tcgen05.mma, TMA) are issued without a per-value waterfall loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// per-warpgroup register budget (all four warps of a warpgroup execute it): the epilogue warps take registers the
// producer / issuer warps give back
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// non-blocking phase test
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// named barrier over a subset of the CTA's warps (id 1..15; `nthreads` a multiple of 32)
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- TMA ----------------------------------------------------------------------------------------
// 3-D tiled load {c0 (innermost), c1, c2} of the box the tensor map was encoded with into shared memory; completion
// is signalled on `bar` as transaction bytes.  `tmap` points at a CUtensorMap in kernel-parameter space
// (__grid_constant__).
__device__ __forceinline__ void tma_load_3d(uint32_t dst_saddr, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_saddr, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
// The same 3-D box delivered to the same shared-memory offset of every CTA in `cta_mask` of the cluster; each destination
// CTA's mbarrier (same offset) receives the transaction bytes.
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst_saddr, const void* tmap, int c0, int c1, int c2, uint64_t* bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4}], [%5], %6;"
               ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---- TMEM / tcgen05 ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the MMA unit)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// all previously issued MMAs of this thread arrive on `bar` when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// the same, arriving on the barrier at the same shared-memory offset in every CTA of `cta_mask`
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, one thread issues
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (lane_base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 8 consecutive 32-bit columns: thread t of the warp writes lane (lane_base + t)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, float4 a, float4 b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
                 "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T: A = 128 lanes x 8 columns (one TF32 element per 32-bit column) at a_tmem
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory matrix descriptor (SWIZZLE_128B): start address, leading / stride byte offsets in
// 16-byte units, descriptor version 1 (sm_100), layout type 2 at bits [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// K-major tile, k-step ks (8 contraction elements = 32 B inside the 128 B swizzle row)
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t tile_saddr, int ks) {
  return make_smem_desc(tile_saddr + ks * (UMMA_K * 4), 16, 1024);
}
// MN-major tile, k-step ks = contraction rows [8 ks, 8 ks + 8): leading offset = stride between 32-wide M/N
// groups (KBLK rows of 128 B), stride offset = stride between 4-row swizzle atoms along the contraction.
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t tile_saddr, int ks, int variant = 0) {
  const uint32_t grp = KBLK * 128u;
  return variant == 0 ? make_smem_desc(tile_saddr + ks * 1024u, grp, 512, 1) : make_smem_desc(tile_saddr + ks * 1024u, 512, grp, 1);
}
// Instruction descriptor: fp32 accumulate, TF32 x TF32, dense, M x N, operand majors.
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- tile addressing ------------------------------------------------------------------------------
// byte offset of 16-byte chunk `chunk` (4 contraction elements) of row `row` in a K-major tile
__device__ __forceinline__ uint32_t kmajor_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }
// byte offset of the 16-byte chunk holding M/N indices [4q, 4q+4) of contraction row c (0..31) in an MN-major tile
__device__ __forceinline__ uint32_t mnmajor_off(int q, int c) {
  return (uint32_t)((q >> 3) * (KBLK * 128) + c * 128 + (((((q & 7) >> 1) ^ (c & 3)) << 5) | ((q & 1) << 4)));
}

// ---- TF32 split -------------------------------------------------------------------------------------
// x = hi + lo with hi = x rounded to TF32 (nearest, ties away: add half an ulp of the 10-bit mantissa and clear the
// low 13 bits -- the same result as cvt.rna.tf32.f32, which sm_100a expands into a longer integer sequence) and
// lo = x - hi (exact in fp32; the MMA unit reads the TF32 part of it).  A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi keeps
// ~2^-21 relative accuracy per product (the dropped lo*lo term is 2^-22).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// hi_tile / lo_tile / off: shared-space byte addresses
__device__ __forceinline__ void store_split(uint32_t hi_tile, uint32_t lo_tile, uint32_t off, float4 v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
  sts128(hi_tile + off, h);
  sts128(lo_tile + off, l);
}
__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float4 v) {
  store_split(smem_u32(hi_tile), smem_u32(lo_tile), off, v);
}

}  // namespace tc
}  // namespace swr
