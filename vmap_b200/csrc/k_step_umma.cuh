// K1 (tensor-core flavour): fused PE -> MLP -> render -> loss -> backward for H = 32,
// fp16 operands / fp32 accumulate on the 5th-gen tensor cores (tcgen05.mma, TMEM
// accumulators), weights staged per object with one bulk async copy (TMA engine).
//
// CTA = 512 threads: two "point groups" of 256 threads.  A group owns
// a tile of up to 128 sample points (whole rays); two threads share a point (= one TMEM lane),
// splitting the accumulator columns / PE directions, so the per-stage epilogue latency halves
// and each SM sub-partition has 4 resident compute warps.  Each group walks its tile through
// 12 MMA stages in lock-step: threads write the next operand rows to
// shared memory, meet at the group's named barrier, one elected lane of the group's first warp
// issues the stage's tcgen05.mma batch and commits to the group's mbarrier, threads read the
// accumulator back with tcgen05.ld.  The two groups are independent, so one group's epilogue
// can overlap the other's MMAs.  Weight gradients accumulate in TMEM across all the
// tiles a CTA owns for an object and are flushed once (fp32 atomics) per (CTA, object).
//
// Shared-memory operand layout: SWIZZLE_NONE 8x8 core matrices.  An activation block is
// [feature-group (8 feats)][point][8 feats] halves, i.e. element (p, f) sits at
// (f/8)*2048 + p*16 + (f%8)*2.  The same bytes serve as a K-major A operand (forward /
// dgrad: M = points, K = features) and as an MN-major operand (wgrad: M or N = features,
// K = points) -- only the descriptor's LBO/SBO swap roles.  Weight matrices W[o][k] are
// stored the same way, (k/8)*512 + o*16 + (k%8)*2, and serve as K-major B (forward) and
// MN-major B (dgrad: N = k, K = o).
//
// Feature order inside the embedding blocks is chosen for the PE recurrence (one sincos
// per direction, then angle doubling), and each block carries a constant-1 column so the
// bias gradients fall out of the wgrad MMAs; the weight image (written by the fused Adam
// kernel) is permuted to match.  dY operands carry a static loss scale of 2^8.
//
// Reference arithmetic: embedding.py:82-91, model.py:54-85, render_rays.py:4-96,
// loss.py:5-62 and their autograd backward (train.py:293-324).
#pragma once
#include <string>
#include "common.cuh"
#include "k_step_fp32.cuh"
#include "umma_ptx.cuh"

#define UMMA_MAX_S 16

// Optional cycle trace (bring-up / profiling builds only: -DVMB_TRACE): CTA 0 records clock64()
// at every stage boundary of its first tiles into a global buffer (see tools/trace_umma.py).
#ifdef VMB_TRACE
__device__ long long g_vmb_trace[4][256];
#define TR(row, slot) do { if (blockIdx.x == 0 && (slot) < 256) g_vmb_trace[row][slot] = clock64(); } while (0)
#else
#define TR(row, slot) do { } while (0)
#endif

namespace um {

constexpr int GT = 256;                  // threads per point group
constexpr int NT = 2 * GT;                // no dedicated issuer warps: warp 0 of a group issues its MMAs
constexpr int FGB = 2048;                 // bytes of one 8-feature group for 128 points
// feature-group index of each block inside a group's activation region
constexpr int FG_HC = 0, FG_DH = 4, FG_FC1 = 6, FG_FC2 = 10, FG_E1 = 14, FG_FC3 = 26, FG_FC4 = 30, FG_E2 = 34;
constexpr int FG_TOTAL = 40;
constexpr int ACT_BYTES = FG_TOTAL * FGB;                 // 81920 per group
// weight image (bytes)
constexpr int IMG_WIN = 0, IMG_WM1 = 6144, IMG_WCAT = 8192, IMG_WM2 = 16384, IMG_WCL = 18432;
constexpr int IMG_WA16 = 23552, IMG_WOC16 = 24576, IMG_F32 = 25600;
// fp32 section (float index relative to IMG_F32)
constexpr int F_BIN = 0, F_BM1 = 32, F_BCAT = 64, F_BM2 = 96, F_BCL = 128, F_BA = 160, F_BOC = 161, F_DIRS = 168;
constexpr int IMG_BYTES = 26624;
// shared memory map
constexpr int SM_ACT0 = 0, SM_ACT1 = ACT_BYTES, SM_W = 2 * ACT_BYTES, SM_G = SM_W + IMG_BYTES;
// per-group fp32 scratch: 12 rows of 128 floats, per-ray broadcast slots, dB accumulators, dproj
constexpr int R_OCC = 0, R_F = 1, R_C0 = 2, R_C1 = 3, R_C2 = 4, R_Z = 5, R_W = 6, R_T = 7, R_GW = 8, R_T0 = 9, R_T1 = 10, R_T2 = 11;
constexpr int GS_RB = 12 * 512, GS_DBS = GS_RB + 5 * 512, GS_DPR = GS_DBS + 256;   // rb = [5][128] per-ray gradients
constexpr int GS_BYTES = GS_DPR + 21 * 512;               // 19712
constexpr int SM_MISC = SM_G + 2 * GS_BYTES;
constexpr int SMEM_BYTES = SM_MISC + 256;                 // 230144 <= 227 KB (232448)
// TMEM columns
constexpr int WG_IN = 0, WG_M1 = 32, WG_CAT = 64, WG_M2 = 96, WG_CL = 128, WG_A = 160, WG_OC = 176;
constexpr int ACC0 = 192, ACC_STRIDE = 160;               // per group: A[0,32) B[32,64) E[64,160)
constexpr float LS = 256.0f, INV_LS = 1.0f / 256.0f, HMAX = 60000.0f;

// ---- column maps ------------------------------------------------------------------------
// emb1 block (96 cols): 0 = const 1, 1..3 = xyz/scale, 4..7 = dir 20 (k=0..3),
// 8i+e (i=1..10) = dir 2(i-1)+e/4, k=e%4; 88..95 = 0.
__host__ __device__ inline int emb1_col_to_j(int c) {     // -> reference emb index, -2 ones, -1 pad
  if (c == 0) return -2;
  if (c < 4) return c - 1;
  if (c < 8) return 3 + (c - 4) * VMB_NDIRS + 20;
  if (c >= 88) return -1;
  const int i = c >> 3, e = c & 7;
  return 3 + (e & 3) * VMB_NDIRS + 2 * (i - 1) + (e >> 2);
}
__host__ __device__ inline int j_to_emb1_col(int j) {
  if (j < 3) return 1 + j;
  const int k = (j - 3) / VMB_NDIRS, d = (j - 3) % VMB_NDIRS;
  if (d == 20) return 4 + k;
  return 8 * (d / 2 + 1) + (d & 1) * 4 + k;
}
// emb2 block (48 cols): 8i+e (i=0..4) = dir 4i+e/2, k=4+e%2; 40,41 = dir 20 (k=4,5); 42 = const 1.
__host__ __device__ inline int emb2_col_to_j2(int c) {    // -> index into the reference's emb[87:], -2 ones, -1 pad
  if (c == 42) return -2;
  if (c > 42) return -1;
  int d, k;
  if (c >= 40) { d = 20; k = 4 + (c - 40); } else { d = 4 * (c >> 3) + ((c & 7) >> 1); k = 4 + (c & 1); }
  return 3 + k * VMB_NDIRS + d - VMB_E1;
}
__host__ __device__ inline int j2_to_emb2_col(int j2) {
  const int j = j2 + VMB_E1;
  const int k = (j - 3) / VMB_NDIRS, d = (j - 3) % VMB_NDIRS;
  if (d == 20) return 40 + (k - 4);
  return 8 * (d / 4) + (d & 3) * 2 + (k - 4);
}
// half index of W[o][c] in a 32-row matrix / of W[j][o] in a 16-row (heads) matrix
__host__ __device__ inline int widx32(int base_bytes, int o, int c) { return (base_bytes + (c >> 3) * 512 + o * 16 + (c & 7) * 2) >> 1; }
__host__ __device__ inline int widx16(int base_bytes, int j, int o) { return (base_bytes + (o >> 3) * 256 + j * 16 + (o & 7) * 2) >> 1; }

// wgrad accumulator (block 0..4, lane) -> (param index of out-column 0, stride per out-column), or -1
__device__ __forceinline__ int wg_base(const VmbLayout& L, int blk, int lane, int& ld) {
  ld = 1;
  switch (blk) {
    case 0: {
      if (lane >= 96) return -1;
      const int j = emb1_col_to_j(lane);
      if (j == -2) return L.o_bin;
      if (j < 0) return -1;
      ld = VMB_E1; return L.o_Win + j;
    }
    case 1:
      if (lane < 32) { ld = 32; return L.o_Wm1 + lane; }
      return lane == 64 ? L.o_bm1 : -1;
    case 2: {
      if (lane < 32) { ld = 32 + VMB_E1; return L.o_Wcat + lane; }
      const int j = emb1_col_to_j(lane - 32);
      if (j == -2) return L.o_bcat;
      if (j < 0) return -1;
      ld = 32 + VMB_E1; return L.o_Wcat + 32 + j;
    }
    case 3:
      if (lane < 32) { ld = 32; return L.o_Wm2 + lane; }
      return lane == 64 + 42 ? L.o_bm2 : -1;
    default: {
      if (lane < 32) { ld = 32 + L.e2; return L.o_Wcl + lane; }
      if (lane >= 80) return -1;
      const int j2 = emb2_col_to_j2(lane - 32);
      if (j2 == -2) return L.o_bcl;
      if (j2 < 0) return -1;
      ld = 32 + L.e2; return L.o_Wcl + 32 + j2;
    }
  }
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// sigmoid on the fast path: MUFU.EX2 + MUFU.RCP (rel. error ~1e-6, far below the fp16 operand noise)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ uint32_t relu_h2(uint32_t x) {
  __half2 h = __hmax2(*reinterpret_cast<__half2*>(&x), __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
// dy * (h > 0), packed
__device__ __forceinline__ uint32_t gate_h2(uint32_t dy, uint32_t h) {
  const __half2 m = __hgt2(*reinterpret_cast<__half2*>(&h), __float2half2_rn(0.f));
  __half2 r = __hmul2(*reinterpret_cast<__half2*>(&dy), m);
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bring-up safety net: a protocol bug traps (CUDA error) instead of hanging the GPU.
#ifndef VMB_SPIN_LIMIT
#define VMB_SPIN_LIMIT 50000000u
#endif
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < VMB_SPIN_LIMIT; ++i)
    if (ptx::mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }

struct Misc {
  uint64_t done[2], wbar;
  uint32_t tmem_base;
  int on[3];
  int abort_flag;
};

// ---- the MMA issuer: one elected lane of a converged warp issues a group's tcgen05.mma ------
struct Issuer {
  uint32_t a16, w16, tm;      // (activation base, weight base) >> 4, TMEM base

  // descriptors: lo word = start address >> 4 | (LBO >> 4) << 16, hi word = SBO >> 4 | version 1.
  // Everything but the base is a compile-time constant, so each descriptor costs one add.
  static __device__ __forceinline__ uint64_t mk(uint32_t base16, uint32_t off, uint32_t lbo, uint32_t sbo) {
    const uint32_t lo = base16 + (off >> 4) + ((lbo >> 4) << 16);
    const uint32_t hi = (sbo >> 4) | 0x4000u;
    return ((uint64_t)hi << 32) | lo;
  }
  __device__ __forceinline__ uint64_t a_k(int, int fg, int ks) const { return mk(a16, fg * FGB + ks * 4096, 2048, 128); }
  __device__ __forceinline__ uint64_t x_mn(int, int fg, int ks) const { return mk(a16, fg * FGB + ks * 256, 128, 2048); }
  __device__ __forceinline__ uint64_t w_k(int off, int ks) const { return mk(w16, off + ks * 1024, 512, 128); }
  __device__ __forceinline__ uint64_t w16_k(int off, int ks) const { return mk(w16, off + ks * 512, 256, 128); }
  __device__ __forceinline__ uint64_t w_mn(int off, int ks) const { return mk(w16, off + ks * 256, 128, 512); }
  __device__ __forceinline__ uint64_t w16_mn(int off) const { return mk(w16, off, 128, 256); }

  // the wgrad accumulators are zeroed with tcgen05.st at segment boundaries, so every wgrad MMA accumulates
  __device__ __forceinline__ void wgrad(int g, int col, int, int fgA, int fgB, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ptx::umma_f16(tm + col, x_mn(g, fgA, ks), x_mn(g, fgB, ks), idesc, 1u);
  }

  __device__ __forceinline__ void stage(int g, int st) {
    constexpr uint32_t KK32 = ptx::idesc_f16(128, 32, 0, 0), KK16 = ptx::idesc_f16(128, 16, 0, 0);
    constexpr uint32_t KM32 = ptx::idesc_f16(128, 32, 0, 1), KM48 = ptx::idesc_f16(128, 48, 0, 1), KM96 = ptx::idesc_f16(128, 96, 0, 1);
    constexpr uint32_t MM32 = ptx::idesc_f16(128, 32, 1, 1), MM16 = ptx::idesc_f16(128, 16, 1, 1);
    const uint32_t A = tm + ACC0 + g * ACC_STRIDE, B = A + 32, E = A + 64;
    switch (st) {
      case 0:   // in_layer: emb1 (K=96) -> A
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) ptx::umma_f16(A, a_k(g, FG_E1, ks), w_k(IMG_WIN, ks), KK32, ks > 0);
        break;
      case 1:   // mid1: fc1 -> B
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(B, a_k(g, FG_FC1, ks), w_k(IMG_WM1, ks), KK32, ks > 0);
        break;
      case 2:   // cat_layer: [fc2 | emb1] (K=128) -> A
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_f16(A, a_k(g, FG_FC2, ks), w_k(IMG_WCAT, ks), KK32, ks > 0);
        break;
      case 3:   // mid2: fc3 -> B
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(B, a_k(g, FG_FC3, ks), w_k(IMG_WM2, ks), KK32, ks > 0);
        break;
      case 4:   // color_linear: [fc4 | emb2] (K=80) -> A ; out_alpha: fc4 -> B[0..16) col 0
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) ptx::umma_f16(A, a_k(g, FG_FC4, ks), w_k(IMG_WCL, ks), KK32, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(B, a_k(g, FG_FC4, ks), w16_k(IMG_WA16, ks), KK16, ks > 0);
        break;
      case 5:   // out_color: hc -> B[0..16) cols 1..3 (accumulate onto alpha's tile)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(B, a_k(g, FG_HC, ks), w16_k(IMG_WOC16, ks), KK16, 1u);
        break;
      case 6:   // d_hc = dhead @ W_oc -> A ; wgrad out_color, out_alpha
        ptx::umma_f16(A, a_k(g, FG_DH, 0), w16_mn(IMG_WOC16), KM32, 0u);
        wgrad(g, WG_OC, 6, FG_HC, FG_DH, MM16);
        wgrad(g, WG_A, 5, FG_FC4, FG_DH, MM16);
        break;
      case 7:   // d_fc4 = dYc @ W_cl[:, :32] + dhead @ W_a -> B ; wgrad color_linear
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(B, a_k(g, FG_HC, ks), w_mn(IMG_WCL, ks), KM32, ks > 0);
        ptx::umma_f16(B, a_k(g, FG_DH, 0), w16_mn(IMG_WA16), KM32, 1u);
        wgrad(g, WG_CL, 4, FG_FC4, FG_HC, MM32);
        break;
      case 8:   // d_fc3 = dY4 @ W_m2 -> A ; wgrad mid2
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(g, FG_FC4, ks), w_mn(IMG_WM2, ks), KM32, ks > 0);
        wgrad(g, WG_M2, 3, FG_FC3, FG_FC4, MM32);
        break;
      case 9:   // d_fc2 = dY3 @ W_cat[:, :32] -> B ; d_emb1 = dY3 @ W_cat[:, 32:] -> E ; wgrad cat_layer
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(B, a_k(g, FG_FC3, ks), w_mn(IMG_WCAT, ks), KM32, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(E, a_k(g, FG_FC3, ks), w_mn(IMG_WCAT + 4 * 512, ks), KM96, ks > 0);
        wgrad(g, WG_CAT, 2, FG_FC2, FG_FC3, MM32);
        break;
      case 10:  // d_fc1 = dY2 @ W_m1 -> A ; wgrad mid1
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(g, FG_FC2, ks), w_mn(IMG_WM1, ks), KM32, ks > 0);
        wgrad(g, WG_M1, 1, FG_FC1, FG_FC2, MM32);
        break;
      default:  // 11: d_emb1 += dY1 @ W_in -> E ; d_emb2 = dYc @ W_cl[:, 32:] -> A[0..48) (dYc still sits in the hc block) ; wgrad in_layer
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(E, a_k(g, FG_FC1, ks), w_mn(IMG_WIN, ks), KM96, 1u);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(g, FG_HC, ks), w_mn(IMG_WCL + 4 * 512, ks), KM48, ks > 0);
        wgrad(g, WG_IN, 0, FG_E1, FG_FC1, MM32);
        break;
    }
  }
};

// sin(pi 2^k x), k = 0..5: one MUFU sin/cos pair, then angle doubling
__device__ __forceinline__ void sin_ladder(float proj, float (&s)[6]) {
  const float r = proj - 2.0f * rintf(0.5f * proj);          // exact: sin(pi x) has period 2
  s[0] = __sinf(VMB_PI_F * r);
  float c = __cosf(VMB_PI_F * r);
#pragma unroll
  for (int k = 1; k < 6; ++k) {
    const float s2 = s[k - 1] + s[k - 1];
    s[k] = s2 * c;
    c = fmaf(-s2, s[k - 1], 1.0f);
  }
}
// four independent sin ladders interleaved (ILP 4)
__device__ __forceinline__ void sin_ladder4(const float (&proj)[4], float (&s)[4][6]) {
  float c[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const float r = proj[d] - 2.0f * rintf(0.5f * proj[d]);
    s[d][0] = __sinf(VMB_PI_F * r);
    c[d] = __cosf(VMB_PI_F * r);
  }
#pragma unroll
  for (int k = 1; k < 6; ++k)
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const float s2 = s[d][k - 1] + s[d][k - 1];
      s[d][k] = s2 * c[d];
      c[d] = fmaf(-s2, s[d][k - 1], 1.0f);
    }
}
__device__ __forceinline__ void cos_ladder4(const float (&proj)[4], float (&c)[4][6]) {
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const float r = proj[d] - 2.0f * rintf(0.5f * proj[d]);
    c[d][0] = __cosf(VMB_PI_F * r);
  }
#pragma unroll
  for (int k = 1; k < 6; ++k)
#pragma unroll
    for (int d = 0; d < 4; ++d) c[d][k] = fmaf(c[d][k - 1] + c[d][k - 1], c[d][k - 1], -1.0f);
}
// cos(pi 2^k x), k = 0..5
__device__ __forceinline__ void cos_ladder(float proj, float (&c)[6]) {
  const float r = proj - 2.0f * rintf(0.5f * proj);
  c[0] = __cosf(VMB_PI_F * r);
#pragma unroll
  for (int k = 1; k < 6; ++k) c[k] = fmaf(c[k - 1] + c[k - 1], c[k - 1], -1.0f);
}

}  // namespace um

// ---------------------------------------------------------------------------------------
template <int SC>
__global__ void __launch_bounds__(um::NT, 1)
k_step_umma(StepParams a, VmbLayout L, const unsigned char* __restrict__ image, int tpo, int nr, long long T, int phase_delay) {
  using namespace um;
  extern __shared__ __align__(1024) unsigned char smem[];
  Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int S = SC ? SC : a.S, R = a.R;      // SC: compile-time samples/ray (unrolls the render scans)
  if (tid == 0) TR(1, 199);       // kernel entry

  if (tid == 0) {
    ptx::mbar_init(&misc->done[0], 1);  ptx::mbar_init(&misc->done[1], 1);
    ptx::mbar_init(&misc->wbar, 1);
    misc->abort_flag = 0;
    ptx::mbar_init_fence();
  }
  if (warp == 15) { ptx::tmem_alloc(&misc->tmem_base, 512); ptx::tmem_relinquish(); }
  if (tid < 3) {                      // render_rays.py:68-73: one empty mask anywhere zeroes the term for all
    int on = 1;
    if (!a.fwd_only) for (int i = 0; i < a.B; ++i) on &= (a.counts[i * 4 + tid] != 0);
    misc->on[tid] = on;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = misc->tmem_base;
  if (warp < 8) {                     // zero the persistent wgrad accumulators (192 columns x 128 lanes)
    const uint32_t zb = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 96;
#pragma unroll
    for (int c = 0; c < 6; ++c) ptx::tmem_st_zero16(zb + c * 16);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid == 0) TR(1, 200);     // setup done
  ptx::tc_fence_after();

  // this CTA's range of global tiles (tile = nr whole rays of one object)
  const long long gt_begin = (T * blockIdx.x) / gridDim.x, gt_end = (T * (blockIdx.x + 1)) / gridDim.x;
  uint32_t wpar = 0;                  // weight-barrier parity (one completion per segment)
  uint32_t ph = 0;                    // req/done parity of this thread's group

  for (long long gt = gt_begin; gt < gt_end;) {
    const int b = (int)(gt / tpo);
    const int t0 = (int)(gt - (long long)b * tpo);
    const int t1 = (int)min((long long)tpo, (long long)t0 + (gt_end - gt));
    gt += t1 - t0;

    // ---- stage this object's weight image: one bulk copy global -> shared --------------
    if (tid == 0) {
      ptx::mbar_arrive_expect_tx(&misc->wbar, IMG_BYTES);
      ptx::bulk_g2s(smem + SM_W, image + (size_t)b * IMG_BYTES, IMG_BYTES, &misc->wbar);
    }
    mbar_wait_or_trap(&misc->wbar, wpar);
    if (tid == 0) TR(1, 201);   // segment: weights landed
    wpar ^= 1;

    {
      // =========================== point groups ==========================================
      const int g = warp >> 3, tg = tid & (GT - 1);
      const int p = tg & 127, hsel = tg >> 7;          // point slot (= TMEM lane), column / direction half
      unsigned char* act = smem + (g ? SM_ACT1 : SM_ACT0);
      float* sc = reinterpret_cast<float*>(smem + SM_G + g * GS_BYTES);          // [12][128]
      float* rb = reinterpret_cast<float*>(smem + SM_G + g * GS_BYTES + GS_RB);  // [5][128]
      float* dbs = reinterpret_cast<float*>(smem + SM_G + g * GS_BYTES + GS_DBS);
      float* dpr = reinterpret_cast<float*>(smem + SM_G + g * GS_BYTES + GS_DPR);  // [21][128]
      const float* wf = reinterpret_cast<const float*>(smem + SM_W + IMG_F32);
      const float* Bd = wf + F_DIRS;
      const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t tA = tm + ACC0 + g * ACC_STRIDE + lane_base, tB = tA + 32, tE = tA + 64;
      const float isc = 1.0f / a.scale[b];
      const float inv_nd = a.fwd_only ? 0.f : 1.f / ((float)a.counts[b * 4 + 0] + 1e-10f);
      const float inv_no = a.fwd_only ? 0.f : 1.f / ((float)a.counts[b * 4 + 1] + 1e-10f);
      const float inv_ns = a.fwd_only ? 0.f : 1.f / ((float)a.counts[b * 4 + 2] + 1e-10f);
      const float on_d = misc->on[0] ? 1.f : 0.f, on_c = misc->on[1] ? 1.f : 0.f, on_o = misc->on[2] ? 1.f : 0.f;
      float ls_d = 0.f, ls_c = 0.f, ls_o = 0.f;
      const int np = nr * S;
      const int rl = p / S, sidx = p - rl * S;
      if (tg < 64) dbs[tg] = 0.f;
      group_bar(g);
      if (g == 1 && phase_delay > 0 && t0 == (int)(gt_begin - (long long)b * tpo)) {   // de-phase the two groups once
        const long long c0 = clock64();
        while (clock64() - c0 < phase_delay) {}
      }

      Issuer is;
      is.a16 = ptx::smem_u32(act) >> 4;
      is.w16 = ptx::smem_u32(smem + SM_W) >> 4;
      is.tm = tm;
      const bool issuer_warp = (warp & 7) == 0;
      // operands written -> group barrier -> warp 0 of the group issues the stage's MMAs and commits
      // to the group's mbarrier -> everyone waits for the accumulator
#define STAGE_SYNC(ST)                                 \
  do {                                                 \
    ptx::fence_async_smem();                           \
    ptx::tc_fence_before();                            \
    group_bar(g);                                      \
    TRG();                                             \
    if (issuer_warp) {                                 \
      ptx::tc_fence_after();                           \
      if (ptx::elect_one()) {                          \
        TR(2 + g, trs2); ++trs2;                       \
        is.stage(g, ST);                               \
        ptx::umma_commit(&misc->done[g]);              \
        TR(2 + g, trs2); ++trs2;                       \
      }                                                \
      __syncwarp();                                    \
    }                                                  \
    TRG();                                             \
    mbar_wait_or_trap(&misc->done[g], ph);             \
    ph ^= 1;                                           \
    ptx::tc_fence_after();                             \
  } while (0)
      // hidden-layer epilogue on this thread's 16 columns: acc + bias -> ReLU -> fp16 (2 x 16 B)
#define EPI_RELU(TADDR, BIAS_OFF, FG)                                                          \
  do {                                                                                         \
    float v[16];                                                                               \
    ptx::tmem_ld16((TADDR) + 16 * hsel, v);                                                    \
    ptx::tmem_ld_wait();                                                                       \
    const float4* bp = reinterpret_cast<const float4*>(wf + (BIAS_OFF) + 16 * hsel);           \
    uint32_t h[8];                                                                             \
    _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                            \
      const float4 bb = bp[q];                                                                 \
      h[2 * q] = relu_h2(pack_h2(v[4 * q] + bb.x, v[4 * q + 1] + bb.y));                       \
      h[2 * q + 1] = relu_h2(pack_h2(v[4 * q + 2] + bb.z, v[4 * q + 3] + bb.w));               \
    }                                                                                          \
    uint4* dst = reinterpret_cast<uint4*>(act + ((FG) + 2 * hsel) * FGB + p * 16);             \
    dst[0] = make_uint4(h[0], h[1], h[2], h[3]);                                               \
    dst[128] = make_uint4(h[4], h[5], h[6], h[7]);                                             \
  } while (0)
      // dgrad epilogue: dY = (h > 0) * acc, written over h (the stored fp16 activation is the mask)
#define EPI_DGRAD(TADDR, FG)                                                                   \
  do {                                                                                         \
    float v[16];                                                                               \
    ptx::tmem_ld16((TADDR) + 16 * hsel, v);                                                    \
    uint4* dst = reinterpret_cast<uint4*>(act + ((FG) + 2 * hsel) * FGB + p * 16);             \
    const uint4 o0 = dst[0], o1 = dst[128];                                                    \
    ptx::tmem_ld_wait();                                                                       \
    dst[0] = make_uint4(gate_h2(pack_h2(v[0], v[1]), o0.x), gate_h2(pack_h2(v[2], v[3]), o0.y),       \
                        gate_h2(pack_h2(v[4], v[5]), o0.z), gate_h2(pack_h2(v[6], v[7]), o0.w));      \
    dst[128] = make_uint4(gate_h2(pack_h2(v[8], v[9]), o1.x), gate_h2(pack_h2(v[10], v[11]), o1.y),   \
                          gate_h2(pack_h2(v[12], v[13]), o1.z), gate_h2(pack_h2(v[14], v[15]), o1.w)); \
  } while (0)

      // prefetched inputs of the next tile (global-load latency overlaps the current tile)
      float nx = 0.f, ny = 0.f, nz = 0.f, nzv = 0.f, n_gd = 0.f, n_c0 = 0.f, n_c1 = 0.f, n_c2 = 0.f;
      int n_sm = 0;
      auto prefetch = [&](int t) {
        nx = ny = nz = nzv = 0.f; n_gd = n_c0 = n_c1 = n_c2 = 0.f; n_sm = 0;
        if (t >= t1) return;
        const int r0n = t * nr;
        if (p < np && r0n + rl < R) {
          const size_t pi = (size_t)(r0n + rl) * S + sidx;
          const float* pp = a.pcs + (size_t)b * a.pcs_stride + pi * 3;
          nx = pp[0]; ny = pp[1]; nz = pp[2];
          if (hsel == 0 && !a.fwd_only) nzv = a.z[(size_t)b * a.z_stride + pi];
        }
        if (hsel == 1 && p < nr && r0n + p < R && !a.fwd_only) {      // ray threads
          const int ray = r0n + p;
          n_gd = a.gt_depth[(size_t)b * a.gt_depth_stride + ray];
          const float* gcp = a.gt_colour + (size_t)b * a.gt_colour_stride + (size_t)ray * 3;
          n_c0 = gcp[0]; n_c1 = gcp[1]; n_c2 = gcp[2];
          n_sm = (int)a.sem[(size_t)b * a.sem_stride + ray] | ((int)a.mask[(size_t)b * a.mask_stride + ray] << 8) | 0x10000;
        }
      };
      prefetch(t0 + g);

      int trs = 0, trs2 = 0;
#define TRG() do { if (tg == 0) { TR(g, trs); ++trs; } } while (0)
      for (int t = t0 + g; t < t1; t += 2) {
        const int r0 = t * nr;
        TRG();                                      // tile start
        // ---- E0: positional embedding (embedding.py:82-91) ------------------------------
        const float t0x = nx * isc, t1x = ny * isc, t2x = nz * isc, zv = nzv;
        const float gd = n_gd, gc0 = n_c0, gc1 = n_c1, gc2 = n_c2;
        const int smv = n_sm;
        prefetch(t + 2);
        TRG();                                      // E0: prefetch issued
        if (hsel == 0) {
          sc[R_Z * 128 + p] = zv; sc[R_T0 * 128 + p] = t0x; sc[R_T1 * 128 + p] = t1x; sc[R_T2 * 128 + p] = t2x;
        }
        {
          uint4* e1 = reinterpret_cast<uint4*>(act + FG_E1 * FGB + p * 16);
          uint4* e2 = reinterpret_cast<uint4*>(act + FG_E2 * FGB + p * 16);
          const int q0 = hsel ? 3 : 0, q1 = hsel ? 5 : 3;
#pragma unroll 1
          for (int q = q0; q < q1; ++q) {              // directions 4q .. 4q+3
            const float* bq = Bd + q * 12;
            float pj[4], sv[4][6];
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) pj[dd] = fmaf(bq[dd * 3 + 2], t2x, fmaf(bq[dd * 3 + 1], t1x, bq[dd * 3] * t0x));
            sin_ladder4(pj, sv);
            e1[(2 * q + 1) * 128] = make_uint4(pack_h2(sv[0][0], sv[0][1]), pack_h2(sv[0][2], sv[0][3]), pack_h2(sv[1][0], sv[1][1]), pack_h2(sv[1][2], sv[1][3]));
            e1[(2 * q + 2) * 128] = make_uint4(pack_h2(sv[2][0], sv[2][1]), pack_h2(sv[2][2], sv[2][3]), pack_h2(sv[3][0], sv[3][1]), pack_h2(sv[3][2], sv[3][3]));
            e2[q * 128] = make_uint4(pack_h2(sv[0][4], sv[0][5]), pack_h2(sv[1][4], sv[1][5]), pack_h2(sv[2][4], sv[2][5]), pack_h2(sv[3][4], sv[3][5]));
          }
          if (hsel) {
            // direction 20 shares chunk 0 of emb1 with [1, x, y, z] and chunk 5 of emb2 with the const-1 column
            float s[6];
            sin_ladder(fmaf(Bd[62], t2x, fmaf(Bd[61], t1x, Bd[60] * t0x)), s);
            e1[0] = make_uint4(pack_h2(1.0f, t0x), pack_h2(t1x, t2x), pack_h2(s[0], s[1]), pack_h2(s[2], s[3]));
            e2[5 * 128] = make_uint4(pack_h2(s[4], s[5]), pack_h2(1.0f, 0.f), 0u, 0u);
            e1[11 * 128] = make_uint4(0u, 0u, 0u, 0u);
            // zero this point's dhead row (cols 4..15 stay zero; 0..3 are written after the render)
            uint4* dh = reinterpret_cast<uint4*>(act + FG_DH * FGB + p * 16);
            dh[0] = make_uint4(0u, 0u, 0u, 0u); dh[128] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        TRG();                                      // E0: embedding written
        STAGE_SYNC(0);                               // st0: in_layer
        TRG();
        EPI_RELU(tA, F_BIN, FG_FC1);
        TRG();
        STAGE_SYNC(1);                               // st1: mid1
        TRG();
        EPI_RELU(tB, F_BM1, FG_FC2);
        TRG();
        STAGE_SYNC(2);                               // st2: cat_layer
        TRG();
        EPI_RELU(tA, F_BCAT, FG_FC3);
        TRG();
        STAGE_SYNC(3);                               // st3: mid2
        TRG();
        EPI_RELU(tB, F_BM2, FG_FC4);
        TRG();
        STAGE_SYNC(4);                               // st4: color_linear + out_alpha
        TRG();
        EPI_RELU(tA, F_BCL, FG_HC);
        TRG();
        STAGE_SYNC(5);                               // st5: out_color
        TRG();
        // ---- heads: alpha*10 -> sigmoid occupancy, colour sigmoid (model.py:77,83; render_rays.py:6)
        float occ = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (hsel == 0) {
          float v[8];
          tmem_ld8(tB, v);
          ptx::tmem_ld_wait();
          occ = fast_sigmoid((v[0] + wf[F_BA]) * 10.0f);
          c0 = fast_sigmoid(v[1] + wf[F_BOC + 0]); c1 = fast_sigmoid(v[2] + wf[F_BOC + 1]); c2 = fast_sigmoid(v[3] + wf[F_BOC + 2]);
          sc[R_OCC * 128 + p] = occ; sc[R_F * 128 + p] = 1.f - occ + 1e-10f;
          sc[R_C0 * 128 + p] = c0; sc[R_C1 * 128 + p] = c1; sc[R_C2 * 128 + p] = c2;
          if (a.fwd_only && p < np && r0 + rl < R) {     // eval_points (trainer.py:77-90): raw alpha*10, sigmoid colour per point
            const size_t n = (size_t)(r0 + rl) * S + sidx;
            a.out_alpha[(size_t)b * a.alpha_stride + n] = (v[0] + wf[F_BA]) * 10.0f;
            float* oc = a.out_colour + (size_t)b * a.colour_stride + n * 3;
            oc[0] = c0; oc[1] = c1; oc[2] = c2;
          }
        }
        TRG();
        group_bar(g);
        TRG();
        if (a.fwd_only) continue;
        // ---- per-ray: termination weights, rendered depth/colour/opacity, losses, ray gradients
        if (hsel == 1 && p < nr) {
          float gD = 0.f, gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gO = 0.f;
          if (smv & 0x10000) {
            const int ray = r0 + p, pb = p * S;
            float Tr = 1.f, D = 0.f, O = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int q = pb + s;
              const float w = sc[R_OCC * 128 + q] * Tr;                 // render_rays.py:34
              sc[R_W * 128 + q] = w; sc[R_T * 128 + q] = Tr;
              D = fmaf(w, sc[R_Z * 128 + q], D); O += w;
              C0 = fmaf(w, sc[R_C0 * 128 + q], C0); C1 = fmaf(w, sc[R_C1 * 128 + q], C1); C2 = fmaf(w, sc[R_C2 * 128 + q], C2);
              Tr *= sc[R_F * 128 + q];                                  // render_rays.py:29
            }
            float V = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) { const float dz = sc[R_Z * 128 + pb + s] - D; V = fmaf(sc[R_W * 128 + pb + s], dz * dz, V); }
            if (a.r_depth) a.r_depth[(size_t)b * R + ray] = D;
            if (a.r_var) a.r_var[(size_t)b * R + ray] = V;
            if (a.r_opacity) a.r_opacity[(size_t)b * R + ray] = O;
            if (a.r_colour) { float* rc = a.r_colour + ((size_t)b * R + ray) * 3; rc[0] = C0; rc[1] = C1; rc[2] = C2; }
            const int sv = smv & 0xff, mv = (smv >> 8) & 0xff;
            const float m_o = (sv != 0) ? 1.f : 0.f, m_s = (sv != 2) ? 1.f : 0.f, m_d = (mv != 0) ? m_o : 0.f;
            const float info = 1.f / (sqrtf(V) + 1e-4f);                // render_rays.py:74-79
            const float e_d = D - gd, e_o = O - m_o, e_c0 = C0 - gc0, e_c1 = C1 - gc1, e_c2 = C2 - gc2;
            ls_d += on_d * fabsf(e_d) * m_d * info * inv_nd;
            ls_c += on_c * (fabsf(e_c0) + fabsf(e_c1) + fabsf(e_c2)) * m_o * inv_no;
            ls_o += on_o * fabsf(e_o) * m_s * inv_ns;
            gD = LS * on_d * vmb_sign(e_d) * m_d * info * inv_nd;
            const float kc = LS * on_c * a.cs * m_o * inv_no;
            gC0 = kc * vmb_sign(e_c0); gC1 = kc * vmb_sign(e_c1); gC2 = kc * vmb_sign(e_c2);
            gO = LS * on_o * a.os * vmb_sign(e_o) * m_s * inv_ns;
          }
          rb[p] = gD; rb[128 + p] = gC0; rb[256 + p] = gC1; rb[384 + p] = gC2; rb[512 + p] = gO;
        }
        TRG();
        group_bar(g);
        TRG();
        if (!a.backward) continue;
        // ---- per-point: d(loss)/d(raw alpha, raw colour) through the termination product ---
        float Gs = 0.f, gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, wq = 0.f;
        const bool pv = (hsel == 0) && (p < np) && (r0 + rl < R);
        if (pv) {
          gC0 = rb[128 + rl]; gC1 = rb[256 + rl]; gC2 = rb[384 + rl];
          Gs = fmaf(rb[rl], zv, fmaf(gC0, c0, fmaf(gC1, c1, fmaf(gC2, c2, rb[512 + rl]))));
          wq = sc[R_W * 128 + p];
          sc[R_GW * 128 + p] = Gs * wq;
        }
        group_bar(g);
        if (pv) {
          float suffix = 0.f;                                            // sum_{k>s} G_k w_k
          for (int k = sidx + 1; k < S; ++k) suffix += sc[R_GW * 128 + p - sidx + k];
          const float docc = Gs * sc[R_T * 128 + p] - __fdividef(suffix, 1.f - occ + 1e-10f);
          const float da = 10.0f * docc * occ * (1.f - occ);             // model.py:77
          *reinterpret_cast<uint2*>(act + FG_DH * FGB + p * 16) =
              make_uint2(pack_h2(da, gC0 * wq * c0 * (1.f - c0)), pack_h2(gC1 * wq * c1 * (1.f - c1), gC2 * wq * c2 * (1.f - c2)));
        }
        STAGE_SYNC(6);                               // st6: d_hc (+ wgrad heads)
        TRG();
        EPI_DGRAD(tA, FG_HC);
        TRG();
        STAGE_SYNC(7);                               // st7: d_fc4 (+ wgrad color_linear)
        TRG();
        EPI_DGRAD(tB, FG_FC4);
        TRG();
        STAGE_SYNC(8);                               // st8: d_fc3 (+ wgrad mid2)
        TRG();
        EPI_DGRAD(tA, FG_FC3);
        TRG();
        STAGE_SYNC(9);                               // st9: d_fc2, d_emb1 part 1 (+ wgrad cat_layer)
        TRG();
        EPI_DGRAD(tB, FG_FC2);
        TRG();
        STAGE_SYNC(10);                               // st10: d_fc1 (+ wgrad mid1)
        TRG();
        EPI_DGRAD(tA, FG_FC1);
        TRG();
        STAGE_SYNC(11);                               // st11: d_emb1 part 2 -> E, d_emb2 -> A[0..48) (+ wgrad in_layer)
        TRG();
        // ---- PE backward: dproj_d = pi * sum_k 2^k g_{k,d} cos(pi 2^k proj_d) ------------
        {
          const int q0 = hsel ? 3 : 0, q1 = hsel ? 5 : 3;
#pragma unroll 1
          for (int q = q0; q < q1; ++q) {
            float g1a[8], g1b[8], g2[8];
            tmem_ld8(tE + 16 * q + 8, g1a);           // emb1 cols of directions 4q, 4q+1 (k = 0..3)
            tmem_ld8(tE + 16 * q + 16, g1b);          //                          4q+2, 4q+3
            tmem_ld8(tA + 8 * q, g2);                 // emb2 cols (k = 4, 5)
            ptx::tmem_ld_wait();
            const float* bq = Bd + q * 12;
            float pj[4], cv[4][6];
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) pj[dd] = fmaf(bq[dd * 3 + 2], t2x, fmaf(bq[dd * 3 + 1], t1x, bq[dd * 3] * t0x));
            cos_ladder4(pj, cv);
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) {
              const float* g1 = (dd < 2) ? (g1a + dd * 4) : (g1b + (dd - 2) * 4);
              float dp = g1[0] * cv[dd][0];
              dp = fmaf(2.f * g1[1], cv[dd][1], dp);
              dp = fmaf(4.f * g1[2], cv[dd][2], dp);
              dp = fmaf(8.f * g1[3], cv[dd][3], dp);
              dp = fmaf(16.f * g2[dd * 2], cv[dd][4], dp);
              dp = fmaf(32.f * g2[dd * 2 + 1], cv[dd][5], dp);
              dpr[(4 * q + dd) * 128 + p] = dp * VMB_PI_F;
            }
          }
          if (hsel) {                                 // direction 20: emb1 cols 4..7, emb2 cols 40, 41
            float g1[8], g2[8], c[6];
            tmem_ld8(tE, g1);
            tmem_ld8(tA + 40, g2);
            ptx::tmem_ld_wait();
            cos_ladder(fmaf(Bd[62], t2x, fmaf(Bd[61], t1x, Bd[60] * t0x)), c);
            float dp = g1[4] * c[0];
            dp = fmaf(2.f * g1[5], c[1], dp); dp = fmaf(4.f * g1[6], c[2], dp); dp = fmaf(8.f * g1[7], c[3], dp);
            dp = fmaf(16.f * g2[0], c[4], dp); dp = fmaf(32.f * g2[1], c[5], dp);
            dpr[20 * 128 + p] = dp * VMB_PI_F;
          }
        }
        TRG();
        ptx::tc_fence_before();
        group_bar(g);
        ptx::tc_fence_after();
        TRG();
        {   // dB[d][i] += sum_p dproj[d][p] * t_i[p]  (embedding.py:84): warp w owns outputs w, w+8, ...
          const int w8 = tg >> 5, ln = tg & 31;
          float acc[8];
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {           // 8 independent dot products per warp (ILP)
            const int o = min(w8 + 8 * i8, 62);
            const int d = o / 3, i = o - d * 3;
            const float* dp = dpr + d * 128 + ln;
            const float* tp = sc + (R_T0 + i) * 128 + ln;
            float v = dp[0] * tp[0];
            v = fmaf(dp[32], tp[32], v); v = fmaf(dp[64], tp[64], v); v = fmaf(dp[96], tp[96], v);
            acc[i8] = v;
          }
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1)
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) acc[i8] += __shfl_xor_sync(0xffffffffu, acc[i8], sft);
          if (ln < 8 && w8 + 8 * ln < 63) {
            float v = acc[0];
#pragma unroll
            for (int i8 = 1; i8 < 8; ++i8) v = (ln == i8) ? acc[i8] : v;
            dbs[w8 + 8 * ln] += v;
          }
        }
        group_bar(g);                               // scratch is free for the next tile
        TRG();
      }
#undef TRG
#undef STAGE_SYNC
#undef EPI_RELU
#undef EPI_DGRAD
      // ---- per-(CTA, object) flush of the partial sums ---------------------------------------
      ls_d = warp_sum(ls_d); ls_c = warp_sum(ls_c); ls_o = warp_sum(ls_o);
      if ((tid & 31) == 0 && a.loss_terms && hsel == 1) {
        atomicAdd(a.loss_terms + b * 4 + 0, ls_d); atomicAdd(a.loss_terms + b * 4 + 1, ls_c);
        atomicAdd(a.loss_terms + b * 4 + 2, ls_o);
        atomicAdd(a.loss_terms + b * 4 + 3, ls_d + a.cs * ls_c + a.os * ls_o);
      }
      if (a.backward && tg < 63) atomicAdd(a.grads + (size_t)b * L.stride + L.o_B + tg, dbs[tg] * INV_LS);
    }

    // ---- segment end: all MMAs have completed (each group waited on its last commit) -----
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) TR(1, 202);   // segment: all tiles done, before flush
    ptx::tc_fence_after();
    if (a.backward && warp < 8) {
      // flush the wgrad accumulators: TMEM -> registers -> fp32 atomics on the grad block
      float* G = a.grads + (size_t)b * L.stride;
      const int q = warp & 3, half = warp >> 2, lane = q * 32 + (tid & 31);
#pragma unroll 1
      for (int cc = 0; cc < 3; ++cc) {
        const int blk = half * 3 + cc;
        float v[32];
        ptx::tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + blk * 32, v);
        ptx::tmem_ld_wait();
        if (blk < 5) {
          int ld;
          const int base = wg_base(L, blk, lane, ld);
          if (base >= 0) {
            float* gp = G + base;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(gp + j * ld, v[j] * INV_LS);
          }
        } else {    // heads: col 0 = out_alpha, cols 17..19 = out_color rows
          if (lane < 32) {
            atomicAdd(G + L.o_Wa + lane, v[0] * INV_LS);
#pragma unroll
            for (int c = 0; c < 3; ++c) atomicAdd(G + L.o_Woc + c * 32 + lane, v[17 + c] * INV_LS);
          } else if (lane == 32 + 42) {
            atomicAdd(G + L.o_ba, v[0] * INV_LS);
          } else if (lane == 112) {
#pragma unroll
            for (int c = 0; c < 3; ++c) atomicAdd(G + L.o_boc + c, v[17 + c] * INV_LS);
          }
        }
        ptx::tmem_st_zero16(tm + ((uint32_t)(q * 32) << 16) + blk * 32);       // ready for the next object
        ptx::tmem_st_zero16(tm + ((uint32_t)(q * 32) << 16) + blk * 32 + 16);
      }
      ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) TR(1, 203);   // segment: after flush
    ptx::tc_fence_after();
  }

  if (tid == 0) TR(1, 204);       // kernel end
  if (warp == 15) ptx::tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int umma_image_bytes() { return um::IMG_BYTES; }

// param index -> location in the image: t >= 0 half index; t <= -2 float word index -(t+2); -1 none
static void umma_fill_image_index(const VmbLayout& L, int* idx) {
  using namespace um;
  for (int i = 0; i < L.P; ++i) idx[i] = -1;
  auto fslot = [](int f) { return -(IMG_F32 / 4 + f) - 2; };
  const int H = 32;
  for (int o = 0; o < H; ++o) {
    for (int j = 0; j < VMB_E1; ++j) idx[L.o_Win + o * VMB_E1 + j] = widx32(IMG_WIN, o, j_to_emb1_col(j));
    idx[L.o_bin + o] = fslot(F_BIN + o);
    for (int k = 0; k < H; ++k) idx[L.o_Wm1 + o * H + k] = widx32(IMG_WM1, o, k);
    idx[L.o_bm1 + o] = fslot(F_BM1 + o);
    for (int k = 0; k < H + VMB_E1; ++k)
      idx[L.o_Wcat + o * (H + VMB_E1) + k] = widx32(IMG_WCAT, o, k < H ? k : H + j_to_emb1_col(k - H));
    idx[L.o_bcat + o] = fslot(F_BCAT + o);
    for (int k = 0; k < H; ++k) idx[L.o_Wm2 + o * H + k] = widx32(IMG_WM2, o, k);
    idx[L.o_bm2 + o] = fslot(F_BM2 + o);
    for (int k = 0; k < H + L.e2; ++k)
      idx[L.o_Wcl + o * (H + L.e2) + k] = widx32(IMG_WCL, o, k < H ? k : H + j2_to_emb2_col(k - H));
    idx[L.o_bcl + o] = fslot(F_BCL + o);
    idx[L.o_Wa + o] = widx16(IMG_WA16, 0, o);
    for (int c = 0; c < 3; ++c) idx[L.o_Woc + c * H + o] = widx16(IMG_WOC16, 1 + c, o);
  }
  idx[L.o_ba] = fslot(F_BA);
  for (int c = 0; c < 3; ++c) idx[L.o_boc + c] = fslot(F_BOC + c);
  for (int i = 0; i < VMB_NDIRS * 3; ++i) idx[L.o_B + i] = fslot(F_DIRS + i);
}

static int umma_launch_step(const VmbLayout& L, const StepParams& sp, const void* image, cudaStream_t st, std::string& err) {
  using namespace um;
  if (L.H != 32 || L.nfreq != 6) { err = "UMMA step kernel: hidden must be 32 and n_freq 6"; return -4; }
  if (sp.S > UMMA_MAX_S) { err = "UMMA step kernel: n_samples > 16"; return -4; }
  static int n_sm_dev[64] = {};            // per device (one process may drive several GPUs)
  int dev = 0;
  cudaGetDevice(&dev);
  int& n_sm = n_sm_dev[dev & 63];
  if (n_sm == 0) {
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(k_step_umma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_umma<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_umma<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { n_sm = 0; err = std::string("cudaFuncSetAttribute(k_step_umma): ") + cudaGetErrorString(e); return -2; }
  }
  const int nr = 128 / sp.S;
  const int tpo = (sp.R + nr - 1) / nr;
  const long long T = (long long)tpo * sp.B;
  long long grid = (T + 1) / 2;
  if (grid > n_sm) grid = n_sm;
  if (grid < 1) grid = 1;
  const unsigned char* img = (const unsigned char*)image;
  static int phase_delay = -1;
  if (phase_delay < 0) { const char* e = getenv("VMB_PHASE_DELAY"); phase_delay = e ? atoi(e) : 0; }   // experiment knob: cycles by which group 1 is de-phased at start
  if (sp.S == 10)      k_step_umma<10><<<(unsigned)grid, NT, SMEM_BYTES, st>>>(sp, L, img, tpo, nr, T, phase_delay);
  else if (sp.S == 14) k_step_umma<14><<<(unsigned)grid, NT, SMEM_BYTES, st>>>(sp, L, img, tpo, nr, T, phase_delay);
  else                 k_step_umma<0><<<(unsigned)grid, NT, SMEM_BYTES, st>>>(sp, L, img, tpo, nr, T, phase_delay);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { err = std::string("k_step_umma launch: ") + cudaGetErrorString(e); return -2; }
  return 0;
}
