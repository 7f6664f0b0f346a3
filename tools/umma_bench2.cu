// Microbenchmark 2 (round 2): what bounds a dependent tcgen05 stage of the fused step kernel?
//   (a) throughput of tcgen05.mma with the A operand in TMEM (".ts" form) vs shared memory, M=128, K=16;
//   (b) round-trip latency of one stage: issue n MMAs -> tcgen05.commit -> mbarrier wait, n = 1..8;
//   (c) epilogue pieces: tcgen05.ld x16 + wait, tcgen05.st x8 + wait, fence.proxy.async after 2 st.shared.v4,
//       named-barrier round trip of 256 threads.
// Data is garbage (timing only).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench2 tools/umma_bench2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../vmap_b200/csrc/umma_ptx.cuh"

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Cfg { int N; int ts; int b_mn; int n_mma; int nacc; };

template <int TS, int NACC>
__global__ void __launch_bounds__(256) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init_fence(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 512); ptx::tmem_relinquish(); }
  ptx::fence_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp < 4) {      // fill the TMEM A-operand columns [384, 448) with 1.0h pairs
    const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + 384;
    for (int cc = 0; cc < 64; cc += 16) {
      const uint32_t one = 0x3c003c00u;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta + cc), "r"(one) : "memory");
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) {
    const uint32_t sa = ptx::smem_u32(smem), sb = ptx::smem_u32(smem + 96 * 1024);
    const uint32_t idesc = ptx::idesc_f16(128, c.N, 0, c.b_mn);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (ptx::elect_one()) {
      uint64_t ad[4], bd[4];
      for (int ks = 0; ks < 4; ++ks) {
        ad[ks] = ptx::smem_desc(sa + ks * 4096, 2048, 128);
        bd[ks] = c.b_mn ? ptx::smem_desc(sb + ks * 256, 128, 512) : ptx::smem_desc(sb + ks * (c.N * 32), c.N * 16, 128);
      }
      const uint64_t a0 = ad[0], a1 = ad[1], a2 = ad[2], a3 = ad[3], b0 = bd[0], b1 = bd[1], b2 = bd[2], b3 = bd[3];
      const int n4 = c.n_mma >> 2, rem = c.n_mma & 3;
      t0 = clock64();
      for (int i = 0; i < n4; ++i) {          // 4 MMAs per iteration, descriptors in registers
        if (TS) {
          umma_f16_ts(tb, tb + 384, b0, idesc, 1u);
          umma_f16_ts(tb + (NACC > 1 ? 96 : 0), tb + 392, b1, idesc, 1u);
          umma_f16_ts(tb + (NACC > 2 ? 192 : 0), tb + 400, b2, idesc, 1u);
          umma_f16_ts(tb + (NACC > 3 ? 288 : 0), tb + 408, b3, idesc, 1u);
        } else {
          ptx::umma_f16(tb, a0, b0, idesc, 1u);
          ptx::umma_f16(tb + (NACC > 1 ? 96 : 0), a1, b1, idesc, 1u);
          ptx::umma_f16(tb + (NACC > 2 ? 192 : 0), a2, b2, idesc, 1u);
          ptx::umma_f16(tb + (NACC > 3 ? 288 : 0), a3, b3, idesc, 1u);
        }
      }
      for (int i = 0; i < rem; ++i) {
        if (TS) umma_f16_ts(tb, tb + 384, b0, idesc, 1u); else ptx::umma_f16(tb, a0, b0, idesc, 1u);
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
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tb, 512);
}

// (c) epilogue pieces, 256 threads (8 warps), every thread times its own copy; warp 0 lane 0 reports
__global__ void __launch_bounds__(256) pieces(long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init_fence(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 16;
  long long t[8];
  float v[16];
  uint32_t acc = 0;
  const int REP = 64;
  // tcgen05.ld x16 + wait
  __syncthreads();
  t[0] = clock64();
  for (int r = 0; r < REP; ++r) { ptx::tmem_ld16(tl, v); ptx::tmem_ld_wait(); acc += __float_as_uint(v[r & 15]); }
  t[1] = clock64();
  // tcgen05.st x8 + wait
  for (int r = 0; r < REP; ++r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tl), "r"(acc + r) : "memory");
    ptx::tmem_st_wait();
  }
  t[2] = clock64();
  // 2 x st.shared.v4 + fence.proxy.async
  uint4* dst = reinterpret_cast<uint4*>(smem + tid * 16);
  for (int r = 0; r < REP; ++r) { dst[0] = make_uint4(acc, r, 0, 0); dst[256] = make_uint4(r, acc, 0, 0); ptx::fence_async_smem(); }
  t[3] = clock64();
  // named barrier of 256 threads
  for (int r = 0; r < REP; ++r) asm volatile("bar.sync 1, 256;" ::: "memory");
  t[4] = clock64();
  // st.shared + fence + tcgen05 fence + barrier (the full operand hand-off)
  for (int r = 0; r < REP; ++r) {
    dst[0] = make_uint4(acc, r, 0, 0); dst[256] = make_uint4(r, acc, 0, 0);
    ptx::fence_async_smem(); ptx::tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    ptx::tc_fence_after();
  }
  t[5] = clock64();
  // mbarrier try_wait on an already-completed phase (arrive by thread 0, everyone waits)
  uint32_t par = 0;
  for (int r = 0; r < REP; ++r) {
    if (tid == 0) ptx::mbar_arrive(&bar);
    while (!ptx::mbar_try_wait(&bar, par)) {}
    par ^= 1;
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  t[6] = clock64();
  if (tid == 0 && blockIdx.x == 0) {
    for (int i = 0; i < 6; ++i) out[i] = (t[i + 1] - t[i]) / REP;
    out[7] = acc;
  }
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  long long* d_out; cudaMalloc(&d_out, 64);
  cudaFuncSetAttribute(bench<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(pieces, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  auto run = [&](Cfg c, const char* name) {
    for (int rep = 0; rep < 2; ++rep) {
      if (c.ts && c.nacc > 1) bench<1, 4><<<grid, 256, 200 * 1024>>>(c, d_out);
      else if (c.ts) bench<1, 1><<<grid, 256, 200 * 1024>>>(c, d_out);
      else if (c.nacc > 1) bench<0, 4><<<grid, 256, 200 * 1024>>>(c, d_out);
      else bench<0, 1><<<grid, 256, 200 * 1024>>>(c, d_out);
      cudaDeviceSynchronize();
    }
    cudaError_t e = cudaGetLastError();
    long long h[2]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    printf("%-34s N=%3d n_mma=%4d issue %8.1f total %8.1f  (%.1f cyc/mma)  %s\n", name, c.N, c.n_mma, (double)h[0], (double)h[1],
           (double)h[1] / c.n_mma, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  puts("== (a) throughput, 1024 back-to-back MMAs (M=128, K=16)");
  for (int N : {16, 32, 48, 96}) {
    run({N, 0, 0, 1024, 4}, "A smem (SS), B K-major");
    run({N, 1, 0, 1024, 4}, "A tmem (TS), B K-major");
    run({N, 0, 1, 1024, 4}, "A smem (SS), B MN-major");
    run({N, 1, 1, 1024, 4}, "A tmem (TS), B MN-major");
  }
  puts("== (b) stage round trip: issue n MMAs + commit + mbarrier wait (N=32)");
  for (int n : {1, 2, 3, 4, 6, 8, 16}) {
    run({32, 0, 0, n, 1}, "SS chain (1 accumulator)");
    run({32, 1, 0, n, 1}, "TS chain (1 accumulator)");
  }
  puts("== (c) epilogue pieces, cycles per iteration (256 threads, thread 0's view)");
  for (int rep = 0; rep < 2; ++rep) { pieces<<<grid, 256, 64 * 1024>>>(d_out); cudaDeviceSynchronize(); }
  long long h[8]; cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
  printf("tcgen05.ld x16 + wait::ld           %lld\n", h[0]);
  printf("tcgen05.st x8 + wait::st            %lld\n", h[1]);
  printf("2 x st.shared.v4 + fence.proxy.async %lld\n", h[2]);
  printf("bar.sync 256                        %lld\n", h[3]);
  printf("st.shared + fences + bar.sync 256   %lld\n", h[4]);
  printf("mbarrier arrive -> 256 x try_wait + bar.sync  %lld\n", h[5]);
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
