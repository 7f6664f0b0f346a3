// Generic tcgen05 GEMM used by the layer-wise path for wide models (hidden 128 / 256: background
// model, iMAP whole-scene model), where one object's weights and a tile's activations no longer fit
// in shared memory together and the layers run as separate launches over all points.
//
//   D[M][N] (+)= A[M][K] * B[N][K]^T        fp16 operands, fp32 accumulate in TMEM
//
// * 128 x 128 output tile per CTA, K in chunks of 64, 3-stage TMA -> mbarrier -> tcgen05.mma pipeline:
//   warp 4 = TMA producer (one elected lane), warp 5 = MMA issuer, warps 0..3 = epilogue
//   (TMEM -> registers -> global).  96 KB of shared memory and 128 TMEM columns: two CTAs per SM.
// * Operands use the SWIZZLE_NONE core-matrix layout verified by tools/umma_probe.cu.  A 3-D tensor
//   map (8 elements, rows, groups-of-8) makes ONE TMA box land a whole operand stage in exactly that
//   layout, for both majors:
//     K-major  operand, global [rows][ld]  (K contiguous):  dims (8, rows, K/8),  box (8, 128, 8)
//        -> smem (k/8)*2048 + r*16 + (k%8)*2        LBO 2048, SBO 128, k-step +4096
//     MN-major operand, global [K][ld]     (MN contiguous): dims (8, K, MN/8),    box (8, 64, 16)
//        -> smem (m/8)*1024 + k*16 + (m%8)*2        LBO 128,  SBO 1024, k-step +256
//   Out-of-range rows / columns are zero-filled by the TMA engine, so ragged edges need no branches.
// * A may come from two sources concatenated along K (e.g. [fc2 | emb1] for cat_layer).
#pragma once
#include "common.cuh"
#include "umma_ptx.cuh"
#include <cuda.h>

namespace lw {

constexpr int BM = 128, BN = 128, BK = 64, NSTAGE = 3;
constexpr int STAGE_BYTES = (BM + BN) * BK * 2;             // 32 KB
constexpr int GEMM_SMEM = NSTAGE * STAGE_BYTES + 1024;      // + barriers
constexpr int GEMM_THREADS = 192;
constexpr float LS = 256.0f, INV_LS = 1.0f / 256.0f;

enum Epi { EPI_RELU_F16 = 0, EPI_GATE_F16 = 1, EPI_F32 = 2, EPI_ATOMIC = 3 };

struct GemmArgs {
  int M, N;                 // valid output extent
  int K1, K2;               // K taken from A source 1 / source 2 (multiples of 16; K2 may be 0)
  int b_k0;                 // K offset into B for this launch's first chunk (used with split-K on points)
  int ksplit;               // K elements per blockIdx.z slice (0 = no split)
  // epilogue
  const float* bias;        // [N] (EPI_RELU_F16)
  __half* out16; int ldo;   // fp16 output
  const __half* gate; int ldg;          // EPI_GATE_F16: activation whose sign gates the gradient
  const float* r1_row; const float* r1_col; int r1_stride;   // optional rank-1 term: acc += r1_row[m*r1_stride] * r1_col[n]
  float* out32; int ld32; int accumulate;                    // EPI_F32
  float* gdst; int ldgd; int ldgn; int n_lo; int n_valid; int ones_col; float* gbias;
                            // EPI_ATOMIC: G[m*ldgd + n*ldgn] += acc for n_lo <= n < n_valid; n == ones_col -> gbias[m]
  float scale;              // multiplies the accumulator (loss-scale removal)
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(map), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int A_MN, int B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_gemm_umma(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
            const __grid_constant__ CUtensorMap mapB, GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty = full + NSTAGE;
  uint64_t* done = empty + NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // K range of this CTA (split-K along the reduction for the weight-gradient GEMMs)
  int k_begin = 0, k_end = g.K1 + g.K2;
  if (g.ksplit > 0) { k_begin = blockIdx.z * g.ksplit; k_end = min(k_end, k_begin + g.ksplit); }
  const int n_chunks = (k_end - k_begin + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(done, 1);
    ptx::mbar_init_fence();
  }
  if (warp == 5) { ptx::tmem_alloc(tmem_slot, 128); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c % NSTAGE;
        if (c >= NSTAGE) ptx::mbar_wait(&empty[s], ((c / NSTAGE) - 1) & 1);
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + BM * BK * 2;
        ptx::mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        const int k = k_begin + c * BK;                     // global K position of this chunk
        // A: source 1 covers [0, K1), source 2 covers [K1, K1+K2); chunks never straddle (K1 % 64 == 0 or K2 == 0)
        const bool second = (g.K2 > 0) && (k >= g.K1);
        const CUtensorMap* ma = second ? &mapA2 : &mapA1;
        const int ka = second ? k - g.K1 : k;
        if (A_MN) tma_load_3d(sa, ma, &full[s], 0, ka, m0 / 8);           // dims (8, K, M/8)
        else      tma_load_3d(sa, ma, &full[s], 0, m0, ka / 8);           // dims (8, M, K/8)
        const int kb = g.b_k0 + k;
        if (B_MN) tma_load_3d(sb, &mapB, &full[s], 0, kb, n0 / 8);
        else      tma_load_3d(sb, &mapB, &full[s], 0, n0, kb / 8);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::idesc_f16(BM, BN, A_MN, B_MN);
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c % NSTAGE;
        ptx::mbar_wait(&full[s], (c / NSTAGE) & 1);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES), sb = sa + BM * BK * 2;
        const int klen = min(BK, k_end - (k_begin + c * BK));
#pragma unroll 1
        for (int ks = 0; ks * 16 < klen; ++ks) {
          const uint64_t ad = A_MN ? ptx::smem_desc(sa + ks * 256, 128, 1024) : ptx::smem_desc(sa + ks * 4096, 2048, 128);
          const uint64_t bd = B_MN ? ptx::smem_desc(sb + ks * 256, 128, 1024) : ptx::smem_desc(sb + ks * 4096, 2048, 128);
          ptx::umma_f16(tm, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[s]);
      }
      ptx::umma_commit(done);
    }
  } else {
    // ===================== epilogue: TMEM lane = output row =====================
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after();
    const int m = m0 + tid;
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c32 = 0; c32 < BN; c32 += 32) {
      if (n0 + c32 >= g.N) break;                           // warp-uniform
      float v[32];
      ptx::tmem_ld32(taddr + c32, v);
      ptx::tmem_ld_wait();
      if (m >= g.M) continue;
      const int nb = n0 + c32;
      if (EPI == EPI_RELU_F16) {
        __half* o = g.out16 + (size_t)m * g.ldo + nb;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (nb + j >= g.N) break;
          uint32_t h[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a = fmaxf(v[j + 2 * q] * g.scale + g.bias[nb + j + 2 * q], 0.f);
            const float b = fmaxf(v[j + 2 * q + 1] * g.scale + g.bias[nb + j + 2 * q + 1], 0.f);
            __half2 hh = __floats2half2_rn(a, b);
            h[q] = *reinterpret_cast<uint32_t*>(&hh);
          }
          *reinterpret_cast<uint4*>(o + j) = make_uint4(h[0], h[1], h[2], h[3]);
        }
      } else if (EPI == EPI_GATE_F16) {
        __half* o = g.out16 + (size_t)m * g.ldo + nb;
        const __half* gt = g.gate + (size_t)m * g.ldg + nb;
        const float r1 = g.r1_row ? g.r1_row[(size_t)m * g.r1_stride] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (nb + j >= g.N) break;
          const uint4 gv = *reinterpret_cast<const uint4*>(gt + j);
          const __half* gh = reinterpret_cast<const __half*>(&gv);
          uint32_t h[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float a = v[j + 2 * q] * g.scale, b = v[j + 2 * q + 1] * g.scale;
            if (g.r1_row) { a = fmaf(r1, g.r1_col[nb + j + 2 * q], a); b = fmaf(r1, g.r1_col[nb + j + 2 * q + 1], b); }
            a = (__half2float(gh[2 * q]) > 0.f) ? fminf(fmaxf(a, -60000.f), 60000.f) : 0.f;
            b = (__half2float(gh[2 * q + 1]) > 0.f) ? fminf(fmaxf(b, -60000.f), 60000.f) : 0.f;
            __half2 hh = __floats2half2_rn(a, b);
            h[q] = *reinterpret_cast<uint32_t*>(&hh);
          }
          *reinterpret_cast<uint4*>(o + j) = make_uint4(h[0], h[1], h[2], h[3]);
        }
      } else if (EPI == EPI_F32) {
        float* o = g.out32 + (size_t)m * g.ld32 + nb;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (nb + j < g.N) o[j] = (g.accumulate ? o[j] : 0.f) + v[j] * g.scale;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = nb + j;
          if (n >= g.n_lo && n < g.n_valid) atomicAdd(g.gdst + (size_t)m * g.ldgd + (size_t)n * g.ldgn, v[j] * g.scale);
          else if (n == g.ones_col && g.gbias) atomicAdd(g.gbias + m, v[j] * g.scale);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tm, 128);
}

// ---- host: tensor maps ----------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// fp16 matrix with `rows` rows of `cols` contiguous elements, row pitch `ld` elements.
//   mn_major = 0: the matrix is [rows = M or N][cols = K]   (K contiguous)   -> box (8, 128, 8)
//   mn_major = 1: the matrix is [rows = K][cols = M or N]   (MN contiguous)  -> box (8, 64, 16)
static bool make_operand_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int mn_major) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {8, (cuuint64_t)rows, (cuuint64_t)(cols / 8)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 16};
  cuuint32_t box[3] = {8, (cuuint32_t)(mn_major ? BK : BM), (cuuint32_t)(mn_major ? BM / 8 : BK / 8)};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Operand { const void* base; long long rows, cols, ld; };

template <int A_MN, int B_MN, int EPI>
static cudaError_t launch_gemm(const Operand& a1, const Operand& a2, const Operand& b, GemmArgs g, int m_tiles, int n_tiles,
                               int z, cudaStream_t st) {
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_umma<A_MN, B_MN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e != cudaSuccess) return e;
    attr[dev & 63] = true;
  }
  CUtensorMap mA1, mA2, mB;
  if (!make_operand_map(&mA1, a1.base, a1.rows, a1.cols, a1.ld, A_MN)) return cudaErrorInvalidValue;
  if (a2.base) { if (!make_operand_map(&mA2, a2.base, a2.rows, a2.cols, a2.ld, A_MN)) return cudaErrorInvalidValue; }
  else mA2 = mA1;
  if (!make_operand_map(&mB, b.base, b.rows, b.cols, b.ld, B_MN)) return cudaErrorInvalidValue;
  k_gemm_umma<A_MN, B_MN, EPI><<<dim3(m_tiles, n_tiles, z), GEMM_THREADS, GEMM_SMEM, st>>>(mA1, mA2, mB, g);
  return cudaGetLastError();
}

}  // namespace lw
