// Bring-up probe: checks the no-swizzle UMMA shared-memory descriptor conventions used by
// k_step_fused.cuh (K-major and MN-major operands, LBO/SBO roles) and the TMEM load mapping
// against a CPU GEMM.  Usage: umma_probe <a_mn_major> <b_mn_major> <swap_lbo_sbo>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include "../vmap_b200/csrc/umma_ptx.cuh"

struct Cfg { int M, N, K, a_mn, b_mn, swap; };

__global__ void __launch_bounds__(128) probe(const __half* imgA, const __half* imgB, int bytesA, int bytesB, Cfg c,
                                             float* D, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  unsigned char* sA = smem;
  unsigned char* sB = smem + ((bytesA + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < bytesA / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(imgA)[i];
  for (int i = tid; i < bytesB / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(imgB)[i];
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init_fence(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 128); ptx::tmem_relinquish(); }
  ptx::fence_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t SBR_A = 128, SBK_A = (c.M / 8) * 128, SBR_B = 128, SBK_B = (c.N / 8) * 128;
    const uint32_t idesc = ptx::idesc_f16(c.M, c.N, c.a_mn, c.b_mn);
    for (int kk = 0; kk < c.K / 16; ++kk) {
      const uint32_t aaddr = ptx::smem_u32(sA) + kk * 2 * SBK_A;
      const uint32_t baddr = ptx::smem_u32(sB) + kk * 2 * SBK_B;
      const uint64_t ad = c.swap ? ptx::smem_desc(aaddr, SBR_A, SBK_A) : ptx::smem_desc(aaddr, SBK_A, SBR_A);
      const uint64_t bd = c.swap ? ptx::smem_desc(baddr, SBR_B, SBK_B) : ptx::smem_desc(baddr, SBK_B, SBR_B);
      ptx::umma_f16(tb, ad, bd, idesc, kk > 0 ? 1u : 0u);
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
  if (warp == 0) ptx::tmem_dealloc(tb, 128);
}

static void build_image(std::vector<__half>& img, const std::vector<float>& X, int R, int K, int mn_major) {
  // element (r,k): K-major core matrix [r%8][k%8], MN-major core matrix [k%8][r%8]; groups: r/8 -> 128 B, k/8 -> (R/8)*128 B
  img.assign((size_t)R * K, __float2half(0.f));
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) {
      size_t byte = (size_t)(r / 8) * 128 + (size_t)(k / 8) * (R / 8) * 128 +
                    (mn_major ? (k % 8) * 16 + (r % 8) * 2 : (r % 8) * 16 + (k % 8) * 2);
      img[byte / 2] = __float2half(X[(size_t)r * K + k]);
    }
}

int main(int argc, char** argv) {
  const int a_mn = argc > 1 ? atoi(argv[1]) : 0, b_mn = argc > 2 ? atoi(argv[2]) : 0, swap = argc > 3 ? atoi(argv[3]) : 0;
  const int Ns[] = {32, 16, 96, 48}, Ks[] = {16, 32, 96, 128};
  int n_pass = 0, n_tot = 0;
  for (int N : Ns) for (int K : Ks) {
    Cfg c{128, N, K, a_mn, b_mn, swap};
    std::vector<float> A((size_t)c.M * K), B((size_t)N * K);
    srand(1234 + N * 7 + K);
    for (auto& x : A) x = (float)((rand() % 9) - 4) * 0.25f;
    for (auto& x : B) x = (float)((rand() % 9) - 4) * 0.25f;
    std::vector<__half> iA, iB;
    build_image(iA, A, c.M, K, a_mn);
    build_image(iB, B, N, K, b_mn);
    __half *dA, *dB; float* dD; int* dS;
    cudaMalloc(&dA, iA.size() * 2); cudaMalloc(&dB, iB.size() * 2); cudaMalloc(&dD, sizeof(float) * c.M * N); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, iA.data(), iA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, iB.data(), iB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof(float) * c.M * N); cudaMemset(dS, 0, 4);
    const int bytesA = (int)iA.size() * 2, bytesB = (int)iB.size() * 2;
    const int smem = ((bytesA + 1023) / 1024) * 1024 + bytesB + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, bytesA, bytesB, c, dD, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("a_mn=%d b_mn=%d swap=%d N=%d K=%d CUDA ERROR %s\n", a_mn, b_mn, swap, N, K, cudaGetErrorString(e)); return 2; }
    std::vector<float> D((size_t)c.M * N); int st = 0;
    cudaMemcpy(D.data(), dD, sizeof(float) * c.M * N, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < c.M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      double d = fabs(ref - D[(size_t)m * N + n]); if (!(d <= maxerr)) maxerr = d;
    }
    const bool pass = st == 0 && maxerr < 1e-3;
    printf("a_mn=%d b_mn=%d swap=%d N=%3d K=%3d status=%d maxerr=%g %s\n", a_mn, b_mn, swap, N, K, st, maxerr, pass ? "PASS" : "FAIL");
    n_pass += pass; ++n_tot;
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  }
  printf("SUMMARY a_mn=%d b_mn=%d swap=%d : %d/%d\n", a_mn, b_mn, swap, n_pass, n_tot);
  return 0;
}
