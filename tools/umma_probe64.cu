// Bring-up probe for a 64-point-tile variant of the fused step kernel:
//  (1) where do the rows of an M=64 accumulator (cta_group::1) land in TMEM (32x32b dump of all 128 lanes),
//  (2) may D start at lane 16 (two M=64 tiles packed into one column block),
//  (3) register mapping of tcgen05.ld 16x256b.x4,
//  (4) cycles per MMA for M=64 vs M=128 at N=32 (K-major operands, no swizzle).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include "../vmap_b200/csrc/umma_ptx.cuh"

__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

// A: [64 rows][K] K-major no-swizzle image (core matrices 8x8), value A[m][k] = m + 1 for k == 0 else 0
// B: [32 rows][K] K-major, B[n][k] = (n + 1) * 0.001 for k == 0 else 0   ->  D[m][n] = (m+1) * (n+1) * 0.001
__global__ void __launch_bounds__(128) probe64(float* dump32, float* dump16, long long* cyc, int lane_off) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  __half* sA = reinterpret_cast<__half*>(smem);            // 128 rows x 16 k  (4 KB) -- rows 64..127 used by the M=128 timing only
  __half* sB = reinterpret_cast<__half*>(smem + 8192);     // 32 rows x 16 k
  for (int i = tid; i < 4096; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  // K = 16: two 8-wide K chunks; K-major core matrix: byte = (r/8)*128 + (k/8)*(R/8)*128 + (r%8)*16 + (k%8)*2
  // D[m][n] = (m + 1) + 256 (n + 1):  k = 0: A = m+1, B = 1;  k = 1: A = 1, B = 256 (n+1)
  if (tid < 128) { sA[((tid / 8) * 128 + (tid % 8) * 16) / 2] = __float2half((float)(tid + 1)); sA[((tid / 8) * 128 + (tid % 8) * 16) / 2 + 1] = __float2half(1.f); }
  if (tid < 32) { sB[((tid / 8) * 128 + (tid % 8) * 16) / 2] = __float2half(1.f); sB[((tid / 8) * 128 + (tid % 8) * 16) / 2 + 1] = __float2half(256.f * (tid + 1)); }
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init_fence(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 128); ptx::tmem_relinquish(); }
  ptx::fence_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_base;
  // zero all 128 lanes x 64 columns first
  for (int c = 0; c < 64; c += 16) ptx::tmem_st_zero16(tb + ((uint32_t)(warp * 32) << 16) + c);
  ptx::tmem_st_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t par = 0;
  if (tid == 0) {
    // A image for M=64: rows 0..63 -> SBO (8-row group stride) 128, LBO (K chunk stride) = (64/8)*128 = 1024
    // (the image above was laid out for R = 128 rows: K chunk stride 2048; we only use k chunk 0 values, chunk 1 is zero
    //  wherever it is read, so LBO 2048 is fine for both)
    const uint64_t ad = ptx::smem_desc(ptx::smem_u32(sA), 2048, 128);
    const uint64_t bd = ptx::smem_desc(ptx::smem_u32(sB), 512, 128);
    ptx::umma_f16(tb + ((uint32_t)lane_off << 16), ad, bd, ptx::idesc_f16(64, 32, 0, 0), 0u);
    ptx::umma_commit(&bar);
  }
  bool ok = ptx::mbar_wait_bounded(&bar, par, 4000000u); par ^= 1;
  ptx::tc_fence_after();
  if (ok) {
    float v[32];
    ptx::tmem_ld32(tb + ((uint32_t)(warp * 32) << 16), v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) dump32[tid * 32 + j] = v[j];
    float w[16];
    tmem_ld_16x256b_x4(tb + ((uint32_t)(warp * 32 + lane_off) << 16), w);     // 16 lanes x (8 x 4) columns of this quadrant
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) dump16[tid * 16 + j] = w[j];
  } else if (tid == 0) cyc[7] = -1;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // timing: 256 back-to-back MMAs, M=64 then M=128
  for (int pass = 0; pass < 2; ++pass) {
    if (warp == 0) {
      long long t0 = 0, t1 = 0, t2 = 0;
      if (ptx::elect_one()) {
        const uint64_t ad = ptx::smem_desc(ptx::smem_u32(sA), 2048, 128), bd = ptx::smem_desc(ptx::smem_u32(sB), 512, 128);
        const uint32_t idesc = ptx::idesc_f16(pass ? 128 : 64, 32, 0, 0);
        t0 = clock64();
        for (int i = 0; i < 256; ++i) ptx::umma_f16(tb + 64, ad, bd, idesc, 1u);
        t1 = clock64();
        ptx::umma_commit(&bar);
      }
      __syncwarp();
      ptx::mbar_wait(&bar, par);
      if (ptx::elect_one()) { t2 = clock64(); cyc[pass * 2] = t1 - t0; cyc[pass * 2 + 1] = t2 - t0; }
    } else ptx::mbar_wait(&bar, par);
    par ^= 1;
    __syncthreads();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tb, 128);
}

int main() {
  float *d32, *d16; long long* dc;
  cudaMalloc(&d32, 128 * 32 * 4); cudaMalloc(&d16, 128 * 16 * 4); cudaMalloc(&dc, 64);
  cudaFuncSetAttribute(probe64, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int lane_off : {0, 16}) {
    cudaMemset(d32, 0, 128 * 32 * 4); cudaMemset(d16, 0, 128 * 16 * 4); cudaMemset(dc, 0, 64);
    probe64<<<1, 128, 16384>>>(d32, d16, dc, lane_off);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lane_off=%d CUDA ERROR %s\n", lane_off, cudaGetErrorString(e)); return 2; }
    std::vector<float> h32(128 * 32), h16(128 * 16); long long c[8];
    cudaMemcpy(h32.data(), d32, h32.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h16.data(), d16, h16.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(c, dc, 64, cudaMemcpyDeviceToHost);
    printf("== D lane offset %d (status %lld)\n", lane_off, c[7]);
    // D[m][n] = (m+1)(n+1)/16: row m = v[0]*16 - ... use column 0: D[m][0] = (m+1)/16
    printf("TMEM lane -> accumulator row (from column 0; '.' = empty):\n");
    for (int l = 0; l < 128; ++l) {
      const int x = (int)lroundf(h32[l * 32]);
      if (x == 0) printf("  . "); else printf("%3d ", x % 256 - 1);
      if (l % 32 == 31) printf("\n");
    }
    // check columns: D[m][n]/D[m][0] = n+1
    int bad = 0;
    for (int l = 0; l < 128; ++l) if (h32[l * 32] != 0.f) for (int n = 0; n < 32; ++n) if ((int)lroundf(h32[l * 32 + n]) / 256 - 1 != n) ++bad;
    printf("column order check: %d mismatches\n", bad);
    printf("16x256b.x4 mapping, warp 0 (thread: reg -> (row, col)):\n");
    for (int t = 0; t < 32; ++t) {
      printf(" t%02d:", t);
      for (int j = 0; j < 16; ++j) {
        const int x = (int)lroundf(h16[t * 16 + j]);
        if (x == 0) printf("   (.,.)"); else printf(" (%2d,%2d)", x % 256 - 1, x / 256 - 1);
      }
      printf("\n");
    }
    printf("cycles per MMA (N=32, K=16): M=64 issue %.1f complete %.1f | M=128 issue %.1f complete %.1f\n",
           c[0] / 256.0, c[1] / 256.0, c[2] / 256.0, c[3] / 256.0);
  }
  return 0;
}
