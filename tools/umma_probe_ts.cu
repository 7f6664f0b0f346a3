// Bring-up probe (round 2): tcgen05.mma with the A operand in TENSOR MEMORY ("TS" form).  Threads write the fp16 A
// tile with tcgen05.st.32x32b (lane = row m, one 32-bit column = two consecutive K elements) and the result is checked
// against a CPU GEMM.  pack = 0: column c of a K=16 step holds (A[m][2c], A[m][2c+1]); pack = 1: (A[m][c], A[m][c+8]).
// Usage: umma_probe_ts <b_mn_major> <pack>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include "../vmap_b200/csrc/umma_ptx.cuh"

struct Cfg { int M, N, K, b_mn, pack; };

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128) probe(const __half* A, const __half* imgB, int bytesB, Cfg c, float* D, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < bytesB / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(imgB)[i];
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init_fence(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 256); ptx::tmem_relinquish(); }
  ptx::fence_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_base;
  const uint32_t ta = tb + 128;                  // A tile: K/2 columns starting at column 128
  {   // row m = tid: write K/2 packed columns, 8 at a time
    const uint32_t lane_addr = ta + ((uint32_t)(warp * 32) << 16);
    for (int kk = 0; kk < c.K / 16; ++kk) {
      uint32_t r[8];
      for (int cc = 0; cc < 8; ++cc) {
        const int k0 = c.pack ? kk * 16 + cc : kk * 16 + 2 * cc, k1 = c.pack ? kk * 16 + cc + 8 : kk * 16 + 2 * cc + 1;
        const __half lo = A[(size_t)tid * c.K + k0], hi = A[(size_t)tid * c.K + k1];
        r[cc] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   ::"r"(lane_addr + kk * 8), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (tid == 0) {
    const uint32_t SBR_B = 128, SBK_B = (c.N / 8) * 128;
    const uint32_t idesc = ptx::idesc_f16(c.M, c.N, 0, c.b_mn);
    for (int kk = 0; kk < c.K / 16; ++kk) {
      const uint32_t baddr = ptx::smem_u32(smem) + kk * 2 * SBK_B;
      const uint64_t bd = c.b_mn ? ptx::smem_desc(baddr, SBR_B, SBK_B) : ptx::smem_desc(baddr, SBK_B, SBR_B);
      umma_f16_ts(tb, ta + kk * 8, bd, idesc, kk > 0 ? 1u : 0u);
    }
    ptx::umma_commit(&bar);
  }
  const bool ok = ptx::mbar_wait_bounded(&bar, 0, 4000000u);
  ptx::tc_fence_after();
  if (!ok) { if (tid == 0) status[0] = 1; }
  else {
    for (int n0 = 0; n0 < c.N; n0 += 16) {
      float v[16];
      ptx::tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + n0, v);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[(size_t)tid * c.N + n0 + j] = v[j];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tb, 256);
}

static void build_image(std::vector<__half>& img, const std::vector<float>& X, int R, int K, int mn_major) {
  img.assign((size_t)R * K, __float2half(0.f));
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) {
      size_t byte = (size_t)(r / 8) * 128 + (size_t)(k / 8) * (R / 8) * 128 +
                    (mn_major ? (k % 8) * 16 + (r % 8) * 2 : (r % 8) * 16 + (k % 8) * 2);
      img[byte / 2] = __float2half(X[(size_t)r * K + k]);
    }
}

int main(int argc, char** argv) {
  const int b_mn = argc > 1 ? atoi(argv[1]) : 0, pack = argc > 2 ? atoi(argv[2]) : 0;
  const int Ns[] = {32, 16, 96, 48}, Ks[] = {16, 32, 96, 128};
  int n_pass = 0, n_tot = 0;
  for (int N : Ns) for (int K : Ks) {
    Cfg c{128, N, K, b_mn, pack};
    std::vector<float> A((size_t)c.M * K), B((size_t)N * K);
    srand(4321 + N * 7 + K);
    for (auto& x : A) x = (float)((rand() % 9) - 4) * 0.25f;
    for (auto& x : B) x = (float)((rand() % 9) - 4) * 0.25f;
    std::vector<__half> hA(A.size()), iB;
    for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2half(A[i]);
    build_image(iB, B, N, K, b_mn);
    __half *dA, *dB; float* dD; int* dS;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, iB.size() * 2); cudaMalloc(&dD, sizeof(float) * c.M * N); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, iB.data(), iB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof(float) * c.M * N); cudaMemset(dS, 0, 4);
    const int bytesB = (int)iB.size() * 2, smem = bytesB + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, bytesB, c, dD, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("b_mn=%d pack=%d N=%d K=%d CUDA ERROR %s\n", b_mn, pack, N, K, cudaGetErrorString(e)); return 2; }
    std::vector<float> D((size_t)c.M * N); int st = 0;
    cudaMemcpy(D.data(), dD, sizeof(float) * c.M * N, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < c.M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      double d = fabs(ref - D[(size_t)m * N + n]); if (!(d <= maxerr)) maxerr = d;
    }
    const bool pass = st == 0 && maxerr < 1e-3;
    printf("TS b_mn=%d pack=%d N=%3d K=%3d status=%d maxerr=%g %s\n", b_mn, pack, N, K, st, maxerr, pass ? "PASS" : "FAIL");
    n_pass += pass; ++n_tot;
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  }
  printf("SUMMARY TS b_mn=%d pack=%d : %d/%d\n", b_mn, pack, n_pass, n_tot);
  return 0;
}
