// Generic tcgen05 GEMM used by the layer-wise path for wide models (hidden 128 / 256: background
// model, iMAP whole-scene model), where one object's weights and a tile's activations no longer fit
// in shared memory together and the layers run as separate launches over all points.
//
//   D[M][N] (+)= A[M][K] * B[N][K]^T        fp16 operands, fp32 accumulate in TMEM
//
// * 128 x 128 output tile per CTA, K in chunks of 64, 3-stage TMA -> mbarrier -> tcgen05.mma pipeline:
//   warp 4 = TMA producer (one elected lane), warp 5 = MMA issuer, warps 0..3 = epilogue
//   (TMEM -> registers -> global).  96 KB of shared memory and 128 TMEM columns: two CTAs per SM.
// * Operand stages use the 128-byte-swizzle canonical layouts, written by TMA (CU_TENSOR_MAP_SWIZZLE_128B,
//   128-byte inner box so the TMA engine moves full lines; a first version with 16-byte inner boxes into the
//   SWIZZLE_NONE layout was TMA-bound at ~150 TFLOP/s):
//     K-major  operand, global [rows][ld]  (K contiguous):  2-D box (64 k, 128 rows)  -> [128 rows][128 B]
//        descriptor SWIZZLE_128B, SBO 1024 (8 rows), k-step +32 B inside the swizzle atom
//     MN-major operand, global [K][ld]     (MN contiguous): two 2-D boxes (64 mn, 64 k) -> two [64 k][128 B] panels
//        descriptor SWIZZLE_128B, LBO 8192 (next 64-wide panel), SBO 1024 (8 k-rows), k-step +2048 B
//   Out-of-range rows / columns are zero-filled by the TMA engine, so ragged edges need no branches.
// * A may come from two sources concatenated along K (e.g. [fc2 | emb1] for cat_layer).
#pragma once
#include "common.cuh"
#include "umma_ptx.cuh"
#include <cuda.h>
#include <algorithm>

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
  int ksplit;               // K elements per z slice (0 = no split)
  int mt, nt, zt;           // tile counts (filled by launch_gemm)
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

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(map), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// SWIZZLE_128B shared-memory descriptor (layout type 2)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return ptx::smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);
}

template <int A_MN, int B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_gemm_umma(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
            const __grid_constant__ CUtensorMap mapB, GemmArgs g) {
  // Persistent: each CTA walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...  of the (z, m, n) tile space
  // (n fastest, so the N-tiles of one row block run back to back and share A in L2).  Two TMEM accumulators
  // (2 x 128 columns) let the MMAs of tile i+1 overlap the epilogue of tile i.
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;          // [2] accumulator ready for the epilogue
  uint64_t* tempty = tfull + 2;              // [2] accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles_total = g.mt * g.nt * g.zt;
  const int k_total = g.K1 + g.K2;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], 128); }
    ptx::mbar_init_fence();
  }
  if (warp == 5) { ptx::tmem_alloc(tmem_slot, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *tmem_slot;

#define TILE_COORDS(T)                                                              \
  const int tz = (T) / (g.mt * g.nt), trem = (T) - tz * (g.mt * g.nt);              \
  const int m0 = (trem / g.nt) * BM, n0 = (trem % g.nt) * BN;                       \
  const int k_begin = g.ksplit > 0 ? tz * g.ksplit : 0;                             \
  const int k_end = g.ksplit > 0 ? min(k_total, k_begin + g.ksplit) : k_total;      \
  const int n_chunks = (k_end - k_begin + BK - 1) / BK;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      uint32_t cc = 0;                        // running chunk counter across tiles
      for (int t = blockIdx.x; t < n_tiles_total; t += gridDim.x) {
        TILE_COORDS(t)
        for (int c = 0; c < n_chunks; ++c, ++cc) {
          const int s = cc % NSTAGE;
          if (cc >= NSTAGE) ptx::mbar_wait(&empty[s], ((cc / NSTAGE) - 1) & 1);
          unsigned char* sa = smem + s * STAGE_BYTES;
          unsigned char* sb = sa + BM * BK * 2;
          ptx::mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          const int k = k_begin + c * BK;                     // global K position of this chunk
          // A: source 1 covers [0, K1), source 2 covers [K1, K1+K2); chunks never straddle (K1 % 64 == 0 or K2 == 0)
          const bool second = (g.K2 > 0) && (k >= g.K1);
          const CUtensorMap* ma = second ? &mapA2 : &mapA1;
          const int ka = second ? k - g.K1 : k;
          if (A_MN) { tma_load_2d(sa, ma, &full[s], m0, ka); tma_load_2d(sa + 8192, ma, &full[s], m0 + 64, ka); }   // dims (M, K)
          else      tma_load_2d(sa, ma, &full[s], ka, m0);                                                          // dims (K, M)
          const int kb = g.b_k0 + k;
          if (B_MN) { tma_load_2d(sb, &mapB, &full[s], n0, kb); tma_load_2d(sb + 8192, &mapB, &full[s], n0 + 64, kb); }
          else      tma_load_2d(sb, &mapB, &full[s], kb, n0);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::idesc_f16(BM, BN, A_MN, B_MN);
      uint32_t cc = 0, it = 0;
      for (int t = blockIdx.x; t < n_tiles_total; t += gridDim.x, ++it) {
        TILE_COORDS(t)
        (void)m0; (void)n0;
        const uint32_t acc = it & 1u;
        if (it >= 2) { ptx::mbar_wait(&tempty[acc], ((it >> 1) - 1) & 1); ptx::tc_fence_after(); }
        for (int c = 0; c < n_chunks; ++c, ++cc) {
          const int s = cc % NSTAGE;
          ptx::mbar_wait(&full[s], (cc / NSTAGE) & 1);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES), sb = sa + BM * BK * 2;
          const int klen = min(BK, k_end - (k_begin + c * BK));
#pragma unroll 1
          for (int ks = 0; ks * 16 < klen; ++ks) {
            const uint64_t ad = A_MN ? desc_sw128(sa + ks * 2048, 8192, 1024) : desc_sw128(sa + ks * 32, 16, 1024);
            const uint64_t bd = B_MN ? desc_sw128(sb + ks * 2048, 8192, 1024) : desc_sw128(sb + ks * 32, 16, 1024);
            ptx::umma_f16(tm + acc * BN, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty[s]);
        }
        ptx::umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ===================== epilogue: TMEM lane = output row =====================
    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles_total; t += gridDim.x, ++it) {
    TILE_COORDS(t)
    (void)n_chunks; (void)k_end;
    const uint32_t acc = it & 1u;
    ptx::mbar_wait(&tfull[acc], (it >> 1) & 1);
    ptx::tc_fence_after();
    const int m = m0 + tid;
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + acc * BN;
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
    ptx::tc_fence_before();
    ptx::mbar_arrive(&tempty[acc]);            // this thread has drained its lanes of the accumulator
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tm, 256);
#undef TILE_COORDS
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
//   mn_major = 0: the matrix is [rows = M or N][cols = K]   (K contiguous)   -> box (64 k, 128 rows)
//   mn_major = 1: the matrix is [rows = K][cols = M or N]   (MN contiguous)  -> box (64 mn, 64 k), two per stage
static bool make_operand_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int mn_major) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)(mn_major ? BK : BM)};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
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
  static int n_sm[64] = {};
  if (!n_sm[dev & 63]) cudaDeviceGetAttribute(&n_sm[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  g.mt = m_tiles; g.nt = n_tiles; g.zt = z;
  const long long total = (long long)m_tiles * n_tiles * z;
  const int grid = (int)std::min<long long>(total, 2LL * n_sm[dev & 63]);
  k_gemm_umma<A_MN, B_MN, EPI><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(mA1, mA2, mB, g);
  return cudaGetLastError();
}

}  // namespace lw
