// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, fp16) for the operand layouts the step
// kernel could use.  Data is garbage (timing only).  One CTA per SM by default.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../vmap_b200/csrc/umma_ptx.cuh"

struct Var { const char* name; int N; int a_mn, b_mn; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int swz; uint32_t a_step, b_step; int M; };

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, int swz) {
  uint64_t d = ptx::smem_desc(addr, lbo, sbo);
  if (swz) d |= (uint64_t)2 << 61;      // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128) bench(Var v, int n_mma, int ksteps, int nacc, int fresh, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init_fence(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 512); ptx::tmem_relinquish(); }
  ptx::fence_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 0) {
    const uint32_t sa = ptx::smem_u32(smem), sb = ptx::smem_u32(smem + 96 * 1024);
    const uint32_t idesc = ptx::idesc_f16(v.M, v.N, v.a_mn, v.b_mn);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (ptx::elect_one()) {
      uint64_t ad[4], bd[4];
      for (int ks = 0; ks < 4; ++ks) {
        ad[ks] = mkdesc(sa + ks * v.a_step, v.a_lbo, v.a_sbo, v.swz);
        bd[ks] = mkdesc(sb + ks * v.b_step, v.b_lbo, v.b_sbo, v.swz);
      }
      const uint32_t acc = fresh ? 0u : 1u;
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u) ptx::umma_f16(tb + ((u % 16) % nacc) * 32, ad[u & 3], bd[u & 3], idesc, acc);
      }
      t1 = clock64();
      ptx::umma_commit(&bar);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    if (ptx::elect_one()) {
      t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else {
    ptx::mbar_wait(&bar, 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  std::vector<Var> vars = {
      // name                          N   amn bmn  a_lbo a_sbo b_lbo b_sbo swz a_step b_step M
      {"K/K noswz N32 (forward)",      32, 0, 0, 2048, 128, 512, 128, 0, 4096, 1024, 128},
      {"K/K noswz N16",                16, 0, 0, 2048, 128, 256, 128, 0, 4096, 512, 128},
      {"K/K noswz N64",                64, 0, 0, 2048, 128, 1024, 128, 0, 4096, 2048, 128},
      {"K/K noswz N128",              128, 0, 0, 2048, 128, 2048, 128, 0, 4096, 4096, 128},
      {"K/K noswz N256",              256, 0, 0, 2048, 128, 4096, 128, 0, 4096, 8192, 128},
      {"K/MN noswz N32 (dgrad)",       32, 0, 1, 2048, 128, 128, 512, 0, 4096, 256, 128},
      {"K/MN noswz N96 (dgrad emb)",   96, 0, 1, 2048, 128, 128, 512, 0, 4096, 256, 128},
      {"MN/MN noswz N32 (wgrad)",      32, 1, 1, 128, 2048, 128, 2048, 0, 256, 256, 128},
      {"MN/MN noswz N16 (wgrad hd)",   16, 1, 1, 128, 2048, 128, 2048, 0, 256, 256, 128},
      {"MN/MN noswz N32 sbo=128",      32, 1, 1, 2048, 128, 2048, 128, 0, 4096, 4096, 128},
      {"K/K sw128 N32",                32, 0, 0, 16, 1024, 16, 1024, 1, 32, 32, 128},
      {"K/K sw128 N64",                64, 0, 0, 16, 1024, 16, 1024, 1, 32, 32, 128},
      {"K/K sw128 N128",              128, 0, 0, 16, 1024, 16, 1024, 1, 32, 32, 128},
      {"K/K sw128 N256",              256, 0, 0, 16, 1024, 16, 1024, 1, 32, 32, 128},
      {"MN/MN sw128 N32",              32, 1, 1, 2048, 1024, 2048, 1024, 1, 2048, 2048, 128},
      {"K/K noswz N32 M64",            32, 0, 0, 2048, 128, 512, 128, 0, 4096, 1024, 64},
      {"K/K sw128 N32 M64",            32, 0, 0, 16, 1024, 16, 1024, 1, 32, 32, 64},
  };
  long long* d_out; cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int accs[] = {1, 2, 4, 8, 16};
  for (int vi : {0, 7, 3}) {
    Var v = vars[vi];
    for (int fresh = 0; fresh < 2; ++fresh)
      for (int nacc : accs) {
        if (v.N * nacc > 512) continue;
        const int n = 2048, ks = 4;
        for (int rep = 0; rep < 2; ++rep) { bench<<<grid, 128, 200 * 1024>>>(v, n, ks, nacc, fresh, d_out); cudaDeviceSynchronize(); }
        long long h[2]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("%-28s nacc=%2d %s issue %7.1f complete %7.1f cyc/mma\n", v.name, nacc, fresh ? "overwrite " : "accumulate", (double)h[0] / n, (double)h[1] / n);
      }
  }
  return 0;
}
