// tc_probe.cu -- stand-alone check of the tcgen05 plumbing in csrc/swr_tc.cuh on a real B200.
//   D[128, N] = sum_c A(m, c) * B(n, c)   one CTA, 2-stage pipeline, operands staged by the CTA's threads.
// usage: tc_probe <aMN 0|1> <bMN 0|1> <variant 0|1> <N> <C> <split3 0|1>
//   aMN/bMN: operand is staged MN-major (source stored [C][M]) instead of K-major (source [M][C]).
// Prints max |err| against a double-precision CPU product.  Build: see tools/tc_probe.sh
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../scenario-wise-rec_b200/csrc/swr_tc.cuh"

using namespace swr::tc;

struct ProbeParams {
  const float* A; const float* B; float* D;
  int N, C, aMN, bMN, variant, split3, positive, aTMEM;
};

__global__ void __launch_bounds__(256, 1) probe_kernel(ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_free[2];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, C = p.C;
  const int Npad = (N + 31) / 32 * 32;       // rows of the B tile in shared memory
  const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)Npad * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  auto a_hi = [&](int s) { return smem + s * stage_bytes; };
  auto a_lo = [&](int s) { return smem + s * stage_bytes + a_bytes; };
  auto b_hi = [&](int s) { return smem + s * stage_bytes + 2 * a_bytes; };
  auto b_lo = [&](int s) { return smem + s * stage_bytes + 2 * a_bytes + b_bytes; };

  if (tid == 0) { mbar_init(&bar_free[0], 1); mbar_init(&bar_free[1], 1); mbar_init(&bar_done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_tf32(128, (N + 15) / 16 * 16, p.aMN != 0, p.bMN != 0);

  const int nkb = C / KBLK;
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb & 1;
    if (kb >= 2) { mbar_wait(&bar_free[s], ((kb >> 1) - 1) & 1); fence_after_sync(); }
    const int c0 = kb * KBLK;
    // ---- stage A ----
    if (p.aTMEM) {   // registers -> TMEM: warp w writes lanes 32*(w%4).., columns 8*(w/4).. of the hi block, +32 for the lo block
      const int row = 32 * (warp & 3) + lane, cg = warp >> 2;   // 8 warps: cg = 0, 1 -> each warp does two column groups
      for (int half = 0; half < 2; ++half) {
        const int c8 = 8 * (cg * 2 + half);
        const float4 x0 = *reinterpret_cast<const float4*>(p.A + (size_t)row * C + c0 + c8);
        const float4 x1 = *reinterpret_cast<const float4*>(p.A + (size_t)row * C + c0 + c8 + 4);
        float4 h0, l0, h1, l1;
        split_tf32(x0.x, h0.x, l0.x); split_tf32(x0.y, h0.y, l0.y); split_tf32(x0.z, h0.z, l0.z); split_tf32(x0.w, h0.w, l0.w);
        split_tf32(x1.x, h1.x, l1.x); split_tf32(x1.y, h1.y, l1.y); split_tf32(x1.z, h1.z, l1.z); split_tf32(x1.w, h1.w, l1.w);
        const uint32_t ta = tmem + 256 + (uint32_t)s * 64 + ((uint32_t)(32 * (warp & 3)) << 16);
        tmem_st8(ta + c8, h0, h1);
        tmem_st8(ta + 32 + c8, l0, l1);
      }
      tmem_st_wait();
      fence_before_sync();
    } else
    for (int v = tid; v < 128 * 8; v += 256) {
      float4 x; uint32_t off;
      if (!p.aMN) { const int r = v >> 3, j = v & 7; x = *reinterpret_cast<const float4*>(p.A + (size_t)r * C + c0 + 4 * j); off = kmajor_off(r, j); }
      else { const int q = v & 31, c = v >> 5; x = *reinterpret_cast<const float4*>(p.A + (size_t)(c0 + c) * 128 + 4 * q); off = mnmajor_off(q, c); }
      store_split(a_hi(s), a_lo(s), off, x);
    }
    // ---- stage B ----
    for (int v = tid; v < Npad * 8; v += 256) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f); uint32_t off;
      if (!p.bMN) { const int r = v >> 3, j = v & 7; if (r < N) x = *reinterpret_cast<const float4*>(p.B + (size_t)r * C + c0 + 4 * j); off = kmajor_off(r, j); }
      else { const int nq = Npad / 4; const int q = v % nq, c = v / nq; if (4 * q < N) x = *reinterpret_cast<const float4*>(p.B + (size_t)(c0 + c) * N + 4 * q); off = mnmajor_off(q, c); }
      store_split(b_hi(s), b_lo(s), off, x);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t ah = smem_u32(a_hi(s)), al = smem_u32(a_lo(s)), bh = smem_u32(b_hi(s)), bl = smem_u32(b_lo(s));
      for (int ks = 0; ks < KBLK / UMMA_K; ++ks) {
        const uint64_t dah = p.aMN ? mnmajor_desc(ah, ks, p.variant) : kmajor_desc(ah, ks);
        const uint64_t dal = p.aMN ? mnmajor_desc(al, ks, p.variant) : kmajor_desc(al, ks);
        const uint64_t dbh = p.bMN ? mnmajor_desc(bh, ks, p.variant) : kmajor_desc(bh, ks);
        const uint64_t dbl = p.bMN ? mnmajor_desc(bl, ks, p.variant) : kmajor_desc(bl, ks);
        const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
        if (p.aTMEM) {
          const uint32_t ta = tmem + 256 + (uint32_t)s * 64 + 8 * ks;
          mma_tf32_ts(tmem, ta + 32, dbh, idesc, first);
          mma_tf32_ts(tmem, ta, dbl, idesc, 1u);
          mma_tf32_ts(tmem, ta, dbh, idesc, 1u);
        } else if (p.split3 == 2) {   // corrections in their own accumulator (columns 256..)
          mma_tf32(tmem + 256, dal, dbh, idesc, first);
          mma_tf32(tmem + 256, dah, dbl, idesc, 1u);
          mma_tf32(tmem, dah, dbh, idesc, first);
        } else if (p.split3) {
          mma_tf32(tmem, dal, dbh, idesc, first);
          mma_tf32(tmem, dah, dbl, idesc, 1u);
          mma_tf32(tmem, dah, dbh, idesc, 1u);
        } else {
          mma_tf32(tmem, dah, dbh, idesc, first);
        }
      }
      mma_commit(&bar_free[s]);
      if (kb == nkb - 1) mma_commit(&bar_done);
    }
  }
  mbar_wait(&bar_done, 0);
  fence_after_sync();
  // epilogue: warp w reads lanes 32*(w%4).., column half (w/4)
  const int lane_base = 32 * (warp & 3);
  const int Nmma = (N + 15) / 16 * 16;
  for (int cb = (warp >> 2) * 32; cb < Nmma; cb += 64) {
    uint32_t r[32];
    if (cb + 32 <= Nmma) {
      tmem_ld32(tmem + ((uint32_t)lane_base << 16) + cb, r);
    } else {
      uint32_t r16[16];
      tmem_ld16(tmem + ((uint32_t)lane_base << 16) + cb, r16);
      for (int i = 0; i < 16; ++i) r[i] = r16[i];
      for (int i = 16; i < 32; ++i) r[i] = 0;
    }
    tmem_ld_wait();
    if (p.split3 == 2) {
      uint32_t r2[32];
      if (cb + 32 <= Nmma) tmem_ld32(tmem + 256 + ((uint32_t)lane_base << 16) + cb, r2);
      else { uint32_t r16[16]; tmem_ld16(tmem + 256 + ((uint32_t)lane_base << 16) + cb, r16); for (int i = 0; i < 16; ++i) r2[i] = r16[i]; for (int i = 16; i < 32; ++i) r2[i] = 0; }
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
    }
    const int m = lane_base + lane;
    for (int i = 0; i < 32; ++i)
      if (cb + i < N) p.D[(size_t)m * N + cb + i] = __uint_as_float(r[i]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  if (argc < 7) { printf("usage: tc_probe aMN bMN variant N C split3\n"); return 2; }
  ProbeParams p{};
  p.aMN = atoi(argv[1]); p.bMN = atoi(argv[2]); p.variant = atoi(argv[3]); p.N = atoi(argv[4]); p.C = atoi(argv[5]); p.split3 = atoi(argv[6]); p.positive = argc > 7 ? atoi(argv[7]) : 0; p.aTMEM = argc > 8 ? atoi(argv[8]) : 0;
  const int M = 128, N = p.N, C = p.C;
  std::vector<float> A((size_t)M * C), B((size_t)N * C), D((size_t)M * N, 0.f);
  srand(1234);
  const int pos = p.positive;
  auto rnd = [pos]() { float v = (float)((rand() % 20001) - 10000) / 10000.f; return pos ? fabsf(v) + 0.001f : v; };
  std::vector<double> Am((size_t)M * C), Bm((size_t)N * C);
  for (int m = 0; m < M; ++m) for (int c = 0; c < C; ++c) { float v = rnd(); Am[(size_t)m * C + c] = v; if (p.aMN) A[(size_t)c * M + m] = v; else A[(size_t)m * C + c] = v; }
  for (int n = 0; n < N; ++n) for (int c = 0; c < C; ++c) { float v = rnd(); Bm[(size_t)n * C + c] = v; if (p.bMN) B[(size_t)c * N + n] = v; else B[(size_t)n * C + c] = v; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  p.A = dA; p.B = dB; p.D = dD;
  const int Npad = (N + 31) / 32 * 32;
  const size_t smem = 2 * (2 * 128 * 128 + 2 * (size_t)Npad * 128) + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 256, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("aMN=%d bMN=%d var=%d N=%d C=%d split3=%d : CUDA ERROR %s\n", p.aMN, p.bMN, p.variant, N, C, p.split3, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0.0, maxref = 0.0, serr = 0.0, sref = 0.0, serr32 = 0.0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0.0;
      for (int c = 0; c < C; ++c) acc += Am[(size_t)m * C + c] * Bm[(size_t)n * C + c];
      float acc32 = 0.f;
      for (int c = 0; c < C; ++c) acc32 = fmaf((float)Am[(size_t)m * C + c], (float)Bm[(size_t)n * C + c], acc32);
      serr += (double)D[(size_t)m * N + n] - acc; sref += fabs(acc); serr32 += (double)acc32 - acc;
      maxerr = fmax(maxerr, fabs(acc - (double)D[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(acc));
    }
  printf("aTMEM=%d aMN=%d bMN=%d var=%d N=%d C=%d split3=%d pos=%d : max|err| = %.3e (max|ref| = %.2f) mean signed err / mean|ref| = %.3e (cpu fp32 fma chain: %.3e) %s\n",
         p.aTMEM, p.aMN, p.bMN, p.variant, N, C, p.split3, p.positive, maxerr, maxref, serr / sref, serr32 / sref, maxerr < (p.split3 ? 1e-6 * maxref * 4 : 5e-2) ? "OK" : "MISMATCH");
  return 0;
}
