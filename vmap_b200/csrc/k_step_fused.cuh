// K1, round 2: ONE launch per optimisation step for hidden = 32.
//   mask counts (render_rays.py:68,86) -> PE -> MLP -> volume render -> losses -> backward
//   -> per-object gradient reduction in a fixed order -> AdamW + fp16 weight-image refresh
// fp16 operands / fp32 accumulate on tcgen05 (TMEM accumulators), weights staged per object with one bulk
// async copy.  CTA = 2 point groups x 256 threads; a group owns a tile of up to 128 sample points (whole rays; one TMEM
// lane = one point, two threads per point split the accumulator columns / PE directions) and walks it through 12
// dependent MMA stages (6 forward, 6 backward): threads write the stage's operands -> group barrier -> one elected
// lane of the group's first warp issues the tcgen05.mma batch and commits to the group's mbarrier -> threads read the
// accumulator back with tcgen05.ld.  The two groups are independent, so one group's epilogue overlaps the other's
// MMAs.  Shared-memory operand layout: SWIZZLE_NONE 8x8 core matrices, an activation block is
// [feature-group (8 feats)][point][8 feats] halves -- the same bytes serve as a K-major A operand (dgrad: M = points,
// K = features) and as an MN-major operand (wgrad: M or N = features, K = points).  Weight gradients accumulate in
// TMEM across all the tiles a CTA owns for an object.  What the round-1 / round-2 cycle traces showed to be on the
// critical path is off it:
//   * the forward MMAs read their A operand from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc: 16 cycles at N = 32
//     instead of 40, tools/umma_bench2.cu): PE and the layer epilogues tcgen05.st the fp16 rows beside the shared-memory
//     copy the weight-gradient MMAs need later, and no generic->async proxy fence sits on the forward chain;
//   * rays never straddle a warp (32/S rays per warp), so transmittance, the five rendered sums, the variance and
//     the backward suffix sum are warp-shuffle scans over the sample axis in registers -- no per-ray serial loop,
//     no fp32 scratch rows in shared memory;
//   * dY_l is no longer stored over h_l: one spare 8 KB block (Z) and the blocks whose readers have retired are
//     rotated (dYc->Z, dY4->HC, dY3->FC4, dY2->HC, dY1->FC3), so a stage commits after its 1-4 dgrad MMAs and the
//     8 wgrad MMAs of the layer are issued behind the commit and overlap the next epilogue;
//   * both head weight gradients and both head bias gradients come out of ONE wgrad ([fc4 | emb2 | hc] x dhead);
//   * the PE-direction gradient dB = dproj^T [x y z] is a wgrad MMA too (fp16 dproj block x the coordinate columns
//     of the embedding block) instead of a shared-memory reduction;
//   * K0 (mask counts) runs in the prologue, K2 (AdamW) in the last CTA to finish an object: every (CTA, object)
//     segment writes its gradient partial to its own row, the finisher adds the rows in segment order (bitwise
//     reproducible -- no floating-point atomics anywhere) and applies AdamW exactly as k_adamw does.
//
// Reference arithmetic: embedding.py:82-91, model.py:54-85, render_rays.py:4-96, loss.py:5-62, their autograd
// backward (train.py:293-324) and torch.optim.AdamW.step + zero_grad (train.py:325-326).
#pragma once
#include <string>
#include <cmath>
#include "common.cuh"
#include "k_step_fp32.cuh"
#include "umma_ptx.cuh"
#include "k_umma_image.cuh"     // column maps, weight-image layout, PE ladders, packed-half helpers (namespace um)

// Optional cycle trace (profiling builds only: -DVMB_TRACE), same buffer as the round-1 kernel
#ifdef VMB_TRACE
#define TRF(row, slot) do { if (blockIdx.x == 0 && (slot) < 256) g_vmb_trace[row][slot] = clock64(); } while (0)
#else
#define TRF(row, slot) do { } while (0)
#endif
// finer stamps inside the PE-forward phase of group 0 (row 2), profiling builds with -DVMB_TRACE -DVMB_TRACE_E0
#if defined(VMB_TRACE) && defined(VMB_TRACE_E0)
#define TRE() do { if (tid == 0) { TRF(2, tre); ++tre; } } while (0)
#else
#define TRE() do { } while (0)
#endif

struct FusedExtra {
  float* partials;            // [(B + grid)][stride] per-(CTA, object) gradient partials, row = blockIdx + object
  unsigned int* obj_done;     // [B] segments finished per object (non-cooperative), [B] skip flags, [2] grid arrive / depart (self-resetting)
  const int* counts_in;       // optional [B][4] external mask counts (ray-sharded iMAP: all-reduced by the caller)
  int* counts_pub;            // [B][4] scratch: cooperative launches count each object ONCE (CTA b mod grid) and publish here
  int fuse_adam;              // 1: the finisher applies AdamW; 0: it adds the reduced gradient into `grads`
  float* p; float* m; float* v;
  __half* image_out; const int* img_index; int img_halves;
  int* step_counter;          // optional [B] device step numbers (t = counter + 1, incremented here)
  float step_size, bc2_sqrt;  // host-computed bias corrections when step_counter == nullptr
  double lr, b1d, b2d;
  float log_b1, log_b2;       // ln(beta1), ln(beta2)
  const float2* bc_table;     // [bc_n] (1 - beta1^t, sqrt(1 - beta2^t)) built on the host in double precision, t = index
  int bc_n;
  float lr_wd, one_m_b1, b2, one_m_b2, eps;
  int guard_loss;
  int* status;
  int single_group;           // 1: only point group 0 walks tiles -> one in-order wgrad MMA stream per SM -> bitwise reproducible
  int cooperative;            // 1: cooperative launch (all CTAs resident): an object's CTAs share its reduction + update
  float* loss_sum;            // optional: sum over objects of the weighted loss totals, written by the last CTA to leave
};

namespace uf {

constexpr int GT = 256, NT = 512, FGB = 2048;
// feature-group index of each 8-feature block inside a group's activation region.  Contiguity that the MMAs rely on:
// [FC1|FC2|E1] (mid1 bias row = E1's constant-1 column, cat_layer K = 128), [FC3|FC4|E2|HC] (mid2 / alpha bias rows =
// E2's constant-1 column, color_linear K = 80, heads wgrad M = [fc4 | emb2 | hc]).
constexpr int FG_DH = 0, FG_FC1 = 2, FG_FC2 = 6, FG_E1 = 10, FG_FC3 = 22, FG_FC4 = 26, FG_E2 = 30, FG_HC = 36, FG_Z = 40, FG_DPR = 44;
constexpr int FG_TOTAL = 47;
constexpr int ACT_BYTES = FG_TOTAL * FGB;                         // 96256 per group
constexpr int SM_ACT0 = 0, SM_ACT1 = ACT_BYTES, SM_W = 2 * ACT_BYTES, SM_MISC = SM_W + um::IMG_BYTES;
constexpr int MISC_BYTES = 512;
constexpr int SM_CNT = SM_MISC + MISC_BYTES;                      // int [B][3] mask counts
constexpr int SMEM_MAX = 232448;
constexpr int MAX_OBJ_SMEM = (SMEM_MAX - SM_CNT) / 12;            // objects whose counts fit (1062)
// TMEM columns: persistent wgrad accumulators, then per group 160 columns:
//   backward: acc[0,48) | - | E[64,160) (embedding-gradient tile)
//   forward : acc[0,32) | hA[32,48) activation A-operand | alpha tile[48,64) | emb1 A-operand[64,112) | emb2 A-operand
//             [112,136) | colour tile[136,152)             (VMB_TS_FWD: the forward MMAs read A from tensor memory)
#ifndef VMB_TS_FWD
#define VMB_TS_FWD 1
#endif
constexpr int WG_IN = 0, WG_M1 = 32, WG_CAT = 64, WG_M2 = 96, WG_CL = 128, WG_HD = 160, WG_DB = 176;
constexpr int ACC0 = 192, ACC_STRIDE = 160, ACC_E = 64;
constexpr int TC_HA = 32, TC_E1A = 64, TC_E2A = 112;
constexpr int TC_ALPHA = VMB_TS_FWD ? 48 : 64, TC_COL = VMB_TS_FWD ? 136 : 80;
constexpr float LS = um::LS, INV_LS = um::INV_LS;

struct Misc {
  uint64_t done[2], wb[2], wbar;
  uint32_t tmem_base;
  int on[3];
  int fin, skip;
  float step_size, bc2_sqrt;
  float lsum[8][4];
};
static_assert(sizeof(Misc) <= MISC_BYTES, "Misc");

// cvt + ReLU in one instruction
__device__ __forceinline__ uint32_t pack_relu_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }

// wgrad accumulator (block 0..4, lane) -> (param index of out-column 0, stride per out-column), or -1
__device__ __forceinline__ int wg_base(const VmbLayout& L, int blk, int lane, int& ld) {
  ld = 1;
  switch (blk) {
    case 0: {                                   // A = E1: lanes = emb1 columns
      if (lane >= 96) return -1;
      const int j = um::emb1_col_to_j(lane);
      if (j == -2) return L.o_bin;
      if (j < 0) return -1;
      ld = VMB_E1; return L.o_Win + j;
    }
    case 1:                                     // A = [FC1 | FC2 | E1...]: lane 64 = constant 1
      if (lane < 32) { ld = 32; return L.o_Wm1 + lane; }
      return lane == 64 ? L.o_bm1 : -1;
    case 2: {                                   // A = [FC2 | E1]
      if (lane < 32) { ld = 32 + VMB_E1; return L.o_Wcat + lane; }
      const int j = um::emb1_col_to_j(lane - 32);
      if (j == -2) return L.o_bcat;
      if (j < 0) return -1;
      ld = 32 + VMB_E1; return L.o_Wcat + 32 + j;
    }
    case 3:                                     // A = [FC3 | FC4 | E2...]: lane 64 + 42 = constant 1
      if (lane < 32) { ld = 32; return L.o_Wm2 + lane; }
      return lane == 64 + 42 ? L.o_bm2 : -1;
    default: {                                  // A = [FC4 | E2]
      if (lane < 32) { ld = 32 + L.e2; return L.o_Wcl + lane; }
      if (lane >= 80) return -1;
      const int j2 = um::emb2_col_to_j2(lane - 32);
      if (j2 == -2) return L.o_bcl;
      if (j2 < 0) return -1;
      ld = 32 + L.e2; return L.o_Wcl + 32 + j2;
    }
  }
}

// ---- MMA issue: one elected lane of the group's first warp ------------------------------------------------------
struct Issuer {
  uint32_t a16, w16, tm, acc;      // (activation base, weight base) >> 4, TMEM base, this group's accumulator base

  static __device__ __forceinline__ uint64_t mk(uint32_t base16, uint32_t off, uint32_t lbo, uint32_t sbo) {
    const uint32_t lo = base16 + (off >> 4) + ((lbo >> 4) << 16);
    const uint32_t hi = (sbo >> 4) | 0x4000u;
    return ((uint64_t)hi << 32) | lo;
  }
  __device__ __forceinline__ uint64_t a_k(int fg, int ks) const { return mk(a16, fg * FGB + ks * 4096, 2048, 128); }   // K-major A: M = points
  __device__ __forceinline__ uint64_t x_mn(int fg, int ks) const { return mk(a16, fg * FGB + ks * 256, 128, 2048); }   // MN-major: M/N = features, K = points
  __device__ __forceinline__ uint64_t w_k(int off, int ks) const { return mk(w16, off + ks * 1024, 512, 128); }
  __device__ __forceinline__ uint64_t w16_k(int off, int ks) const { return mk(w16, off + ks * 512, 256, 128); }
  __device__ __forceinline__ uint64_t w_mn(int off, int ks) const { return mk(w16, off + ks * 256, 128, 512); }
  __device__ __forceinline__ uint64_t w16_mn(int off) const { return mk(w16, off, 128, 256); }

  // dW (+)= X^T dY over the tile's 128 points: A = X block (features on M), B = dY block (features on N)
  __device__ __forceinline__ void wgrad(int col, int fgX, int fgDY, uint32_t idesc) const {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ptx::umma_f16(tm + col, x_mn(fgX, ks), x_mn(fgDY, ks), idesc, 1u);
  }

  // the MMAs a stage's epilogue waits for
  __device__ __forceinline__ void stage(int st) const {
    constexpr uint32_t KK32 = ptx::idesc_f16(128, 32, 0, 0), KK16 = ptx::idesc_f16(128, 16, 0, 0);
    constexpr uint32_t KM32 = ptx::idesc_f16(128, 32, 0, 1), KM48 = ptx::idesc_f16(128, 48, 0, 1), KM96 = ptx::idesc_f16(128, 96, 0, 1);
    const uint32_t A = acc, E = acc + ACC_E;
    switch (st) {
#if VMB_TS_FWD
      case 0:   // in_layer: emb1 (K = 96), A from tensor memory
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) ptx::umma_f16_ts(A, acc + TC_E1A + ks * 8, w_k(um::IMG_WIN, ks), KK32, ks > 0);
        break;
      case 1:   // mid1: fc1
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16_ts(A, acc + TC_HA + ks * 8, w_k(um::IMG_WM1, ks), KK32, ks > 0);
        break;
      case 2:   // cat_layer: [fc2 | emb1] (K = 128)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          ptx::umma_f16_ts(A, ks < 2 ? acc + TC_HA + ks * 8 : acc + TC_E1A + (ks - 2) * 8, w_k(um::IMG_WCAT, ks), KK32, ks > 0);
        break;
      case 3:   // mid2: fc3
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16_ts(A, acc + TC_HA + ks * 8, w_k(um::IMG_WM2, ks), KK32, ks > 0);
        break;
      case 4:   // color_linear: [fc4 | emb2] (K = 80) -> A ; out_alpha: fc4 -> alpha tile column 0
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
          ptx::umma_f16_ts(A, ks < 2 ? acc + TC_HA + ks * 8 : acc + TC_E2A + (ks - 2) * 8, w_k(um::IMG_WCL, ks), KK32, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16_ts(acc + TC_ALPHA, acc + TC_HA + ks * 8, w16_k(um::IMG_WA16, ks), KK16, ks > 0);
        break;
      case 5:   // out_color: hc -> its own 16-column tile (columns 1..3), so alpha's tile can be read while this runs
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16_ts(acc + TC_COL, acc + TC_HA + ks * 8, w16_k(um::IMG_WOC16, ks), KK16, ks > 0);
        break;
#else
      case 0:   // in_layer: emb1 (K = 96)
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) ptx::umma_f16(A, a_k(FG_E1, ks), w_k(um::IMG_WIN, ks), KK32, ks > 0);
        break;
      case 1:   // mid1: fc1
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(FG_FC1, ks), w_k(um::IMG_WM1, ks), KK32, ks > 0);
        break;
      case 2:   // cat_layer: [fc2 | emb1] (K = 128)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_f16(A, a_k(FG_FC2, ks), w_k(um::IMG_WCAT, ks), KK32, ks > 0);
        break;
      case 3:   // mid2: fc3
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(FG_FC3, ks), w_k(um::IMG_WM2, ks), KK32, ks > 0);
        break;
      case 4:   // color_linear: [fc4 | emb2] (K = 80) -> A ; out_alpha: fc4 -> alpha tile column 0
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) ptx::umma_f16(A, a_k(FG_FC4, ks), w_k(um::IMG_WCL, ks), KK32, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(acc + TC_ALPHA, a_k(FG_FC4, ks), w16_k(um::IMG_WA16, ks), KK16, ks > 0);
        break;
      case 5:   // out_color: hc -> its own 16-column tile (columns 1..3), so alpha's tile can be read while this runs
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(acc + TC_COL, a_k(FG_HC, ks), w16_k(um::IMG_WOC16, ks), KK16, ks > 0);
        break;
#endif
      case 6:   // d_hc = dhead @ W_oc
        ptx::umma_f16(A, a_k(FG_DH, 0), w16_mn(um::IMG_WOC16), KM32, 0u);
        break;
      case 7:   // d_fc4 = dYc @ W_cl[:, :32] + dhead @ W_a
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(FG_Z, ks), w_mn(um::IMG_WCL, ks), KM32, ks > 0);
        ptx::umma_f16(A, a_k(FG_DH, 0), w16_mn(um::IMG_WA16), KM32, 1u);
        break;
      case 8:   // d_fc3 = dY4 @ W_m2                      (dY4 lives in HC)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(FG_HC, ks), w_mn(um::IMG_WM2, ks), KM32, ks > 0);
        break;
      case 9:   // d_fc2 = dY3 @ W_cat[:, :32]             (dY3 lives in FC4)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(FG_FC4, ks), w_mn(um::IMG_WCAT, ks), KM32, ks > 0);
        break;
      case 10:  // d_fc1 = dY2 @ W_m1                      (dY2 lives in HC)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(FG_HC, ks), w_mn(um::IMG_WM1, ks), KM32, ks > 0);
        break;
      default:  // 11: d_emb1 += dY1 @ W_in (dY1 in FC3) -> E ; d_emb2 = dYc @ W_cl[:, 32:] (dYc in Z) -> A[0..48)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(E, a_k(FG_FC3, ks), w_mn(um::IMG_WIN, ks), KM96, 1u);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(A, a_k(FG_Z, ks), w_mn(um::IMG_WCL + 4 * 512, ks), KM48, ks > 0);
        break;
    }
  }
  // issued BEHIND the stage's commit: nothing waits for these before the next stage's commit (in-order pipe)
  __device__ __forceinline__ void post(int st) const {
    constexpr uint32_t KM96 = ptx::idesc_f16(128, 96, 0, 1);
    constexpr uint32_t MM32 = ptx::idesc_f16(128, 32, 1, 1), MM16 = ptx::idesc_f16(128, 16, 1, 1);
    switch (st) {
      case 6:  wgrad(WG_HD, FG_FC4, FG_DH, MM16); break;            // out_alpha + out_color weights and biases
      case 7:  wgrad(WG_CL, FG_FC4, FG_Z, MM32); break;             // color_linear: [fc4 | emb2] x dYc
      case 8:  wgrad(WG_M2, FG_FC3, FG_HC, MM32); break;            // mid2: fc3 x dY4
      case 9:                                                        // d_emb1 = dY3 @ W_cat[:, 32:] -> E ; cat_layer wgrad
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ptx::umma_f16(acc + ACC_E, a_k(FG_FC4, ks), w_mn(um::IMG_WCAT + 4 * 512, ks), KM96, ks > 0);
        wgrad(WG_CAT, FG_FC2, FG_FC4, MM32);
        break;
      case 10: wgrad(WG_M1, FG_FC1, FG_HC, MM32); break;            // mid1: fc1 x dY2
      case 11: wgrad(WG_IN, FG_E1, FG_FC3, MM32); break;            // in_layer: emb1 x dY1
      default: break;
    }
  }
  __device__ __forceinline__ void dirs_wgrad() const {               // dB = dproj^T [1 x y z ...]
    constexpr uint32_t MM16 = ptx::idesc_f16(128, 16, 1, 1);
    wgrad(WG_DB, FG_DPR, FG_E1, MM16);
  }
};

// Work partition: CTA c owns tile pairs [begin[c], begin[c+1]) of the global list (object-major).  Built on the host
// (fused_partition) so that a CTA whose range crosses an object boundary -- it pays a second flush / weight load /
// pipeline fill -- gets correspondingly fewer pairs: no CTA's cost exceeds the even share by more than one pair.
constexpr int MAX_CTAS = 192;
struct Ranges { int begin[MAX_CTAS + 1]; };
__device__ __forceinline__ int cta_of_pair(const Ranges& rg, int G, int pair) {        // largest c with begin[c] <= pair
  int lo = 0, hi = G - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rg.begin[mid] <= pair) lo = mid; else hi = mid - 1; }
  return lo;
}

}  // namespace uf

// ---------------------------------------------------------------------------------------------------------------
template <int SC>
__global__ void __launch_bounds__(uf::NT, 1)
k_step_fused(StepParams a, FusedExtra x, VmbLayout L, const unsigned char* image, const __grid_constant__ uf::Ranges rg, int tpo, int npo, int nr, int rpw) {
  using namespace uf;
  extern __shared__ __align__(1024) unsigned char smem[];
  Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);
  int* cnt = reinterpret_cast<int*>(smem + SM_CNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = SC ? SC : a.S, R = a.R;
  if (tid == 0) TRF(3, 199);

  if (tid == 0) {
    ptx::mbar_init(&misc->done[0], 1); ptx::mbar_init(&misc->done[1], 1);
    ptx::mbar_init(&misc->wb[0], 1);   ptx::mbar_init(&misc->wb[1], 1);
    ptx::mbar_init(&misc->wbar, 1);
    misc->on[0] = misc->on[1] = misc->on[2] = 1;
    ptx::mbar_init_fence();
  }
  if (warp == 15) { ptx::tmem_alloc(&misc->tmem_base, 512); ptx::tmem_relinquish(); }
  if (tid == 0) TRF(3, 210);          // barriers initialised
  const int G = gridDim.x;
  const long long gt_begin = rg.begin[blockIdx.x], gt_end = rg.begin[blockIdx.x + 1];
  __syncthreads();
  if (tid == 0 && gt_begin < gt_end) {     // first object's weight image: in flight while the mask counts run
    ptx::mbar_arrive_expect_tx(&misc->wbar, um::IMG_BYTES);
    ptx::bulk_g2s(smem + SM_W, image + (size_t)(gt_begin / npo) * um::IMG_BYTES, um::IMG_BYTES, &misc->wbar);
  }
  if (tid == 0) TRF(3, 211);          // TMEM allocated (first __syncthreads passed), first weight copy issued
  // ---- K0 in the prologue: mask counts of EVERY object (the any-empty early-out couples them, render_rays.py:68-73).
  // Cooperative launch: CTA c counts objects c, c + grid, ... ONCE and publishes counts + empty flags + a "published"
  // counter in global memory; every CTA starts its tiles at once and acquires the counts right before its first
  // volume render (thousands of cycles later: the wait is free).  Otherwise every CTA counts everything (148-fold
  // redundant reads of the label / mask bytes: ~3 us of L2 traffic on the launch's critical path).
  const bool pub_counts = x.cooperative && !x.counts_in && !a.fwd_only;
  unsigned int* gbar = x.obj_done + 2 * a.B;            // [0] arrive, [1] depart, [2] objects published, [3..5] empty flags
  if (pub_counts) {
    for (int b = blockIdx.x; b < a.B; b += G) {
      if (tid < 3) cnt[tid] = 0;
      __syncthreads();
      const unsigned char* s = a.sem + (size_t)b * a.sem_stride;
      const unsigned char* m = a.mask + (size_t)b * a.mask_stride;
      const bool vec = (((size_t)s | (size_t)m) & 3) == 0;
      const int nw = vec ? (R >> 2) : 0;
      auto nzb = [](uint32_t v) { v |= v >> 4; v |= v >> 2; v |= v >> 1; return v & 0x01010101u; };
      int nd = 0, no = 0, ns = 0;
      for (int w = tid; w < nw; w += NT) {               // four rays per 32-bit load
        const uint32_t sv = __ldg(reinterpret_cast<const uint32_t*>(s) + w), mv = __ldg(reinterpret_cast<const uint32_t*>(m) + w);
        const uint32_t o = nzb(sv);
        no += __popc(o); nd += __popc(o & nzb(mv)); ns += __popc(nzb(sv ^ 0x02020202u));
      }
      for (int r = (nw << 2) + tid; r < R; r += NT) {
        const int sv = s[r], mo = sv != 0;
        nd += (m[r] != 0) & mo; no += mo; ns += sv != 2;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        nd += __shfl_xor_sync(0xffffffffu, nd, o); no += __shfl_xor_sync(0xffffffffu, no, o); ns += __shfl_xor_sync(0xffffffffu, ns, o);
      }
      if (lane == 0) { atomicAdd(&cnt[0], nd); atomicAdd(&cnt[1], no); atomicAdd(&cnt[2], ns); }
      __syncthreads();
      if (tid == 0) {
        for (int k = 0; k < 3; ++k) {
          x.counts_pub[b * 4 + k] = cnt[k];
          if (cnt[k] == 0) atomicOr(&gbar[3 + k], 1u);
        }
        __threadfence();
        atomicAdd(&gbar[2], 1u);
      }
      __syncthreads();
    }
  } else if (!a.fwd_only) {
    if (x.counts_in) {
      for (int i = tid; i < a.B * 3; i += NT) {
        const int c = x.counts_in[(i / 3) * 4 + (i % 3)];
        cnt[i] = c;
        if (c == 0) misc->on[i % 3] = 0;
      }
    } else {
      // warp w counts objects w, w+16, ...; two objects at a time with every load of both in flight together (the
      // inputs are cold in HBM: one latency per pair of objects instead of one per 8 words)
      const int nw = R >> 2;
      auto finish = [&](int b, int nd, int no, int ns) {
        const unsigned char* s = a.sem + (size_t)b * a.sem_stride;
        const unsigned char* m = a.mask + (size_t)b * a.mask_stride;
        const bool vec = (((size_t)s | (size_t)m) & 3) == 0;
        for (int r = (vec ? (nw << 2) : 0) + lane; r < R; r += 32) {      // tail rays, or everything when unaligned
          const int sv = s[r], mo = sv != 0;
          nd += (m[r] != 0) & mo; no += mo; ns += sv != 2;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          nd += __shfl_xor_sync(0xffffffffu, nd, o); no += __shfl_xor_sync(0xffffffffu, no, o); ns += __shfl_xor_sync(0xffffffffu, ns, o);
        }
        if (lane == 0) {
          cnt[b * 3] = nd; cnt[b * 3 + 1] = no; cnt[b * 3 + 2] = ns;
          if (nd == 0) misc->on[0] = 0;
          if (no == 0) misc->on[1] = 0;
          if (ns == 0) misc->on[2] = 0;
        }
      };
      // every CTA needs every object's counts: each starts at a different object, so the grid does not hit the same
      // cache lines at the same moment
      const int shift = (int)((blockIdx.x * 7u) % (unsigned)a.B);
      for (int iA = warp; iA < a.B; iA += 2 * (NT / 32)) {
        const int iB = iA + NT / 32;
        const bool hasB = iB < a.B;
        const int bA = (iA + shift) % a.B, bB = hasB ? (iB + shift) % a.B : bA;
        const uint32_t* sA = reinterpret_cast<const uint32_t*>(a.sem + (size_t)bA * a.sem_stride);
        const uint32_t* mA = reinterpret_cast<const uint32_t*>(a.mask + (size_t)bA * a.mask_stride);
        const uint32_t* sB = reinterpret_cast<const uint32_t*>(a.sem + (size_t)(hasB ? bB : bA) * a.sem_stride);
        const uint32_t* mB = reinterpret_cast<const uint32_t*>(a.mask + (size_t)(hasB ? bB : bA) * a.mask_stride);
        const bool vA = (((size_t)sA | (size_t)mA) & 3) == 0, vB = hasB && (((size_t)sB | (size_t)mB) & 3) == 0;
        int cA[3] = {0, 0, 0}, cB[3] = {0, 0, 0};
        for (int w0 = 0; w0 < nw; w0 += 320) {                  // four rays per 32-bit load, ten loads per array per lane
          uint32_t sa[10], ma[10], sb[10], mb[10];
#pragma unroll
          for (int u = 0; u < 10; ++u) {
            const int w = w0 + u * 32 + lane;
            const bool in = w < nw;
            sa[u] = (in && vA) ? __ldg(sA + w) : 0u;  ma[u] = (in && vA) ? __ldg(mA + w) : 0u;
            sb[u] = (in && vB) ? __ldg(sB + w) : 0u;  mb[u] = (in && vB) ? __ldg(mB + w) : 0u;
          }
          // one bit per byte: nonzero(x) folds a byte's bits into bit 0 (labels are 0 / 1 / 2, masks any nonzero = true)
          auto nzb = [](uint32_t v) { v |= v >> 4; v |= v >> 2; v |= v >> 1; return v & 0x01010101u; };
#pragma unroll
          for (int u = 0; u < 10; ++u) {
            const bool in = w0 + u * 32 + lane < nw;
            const uint32_t oa = (in && vA) ? nzb(sa[u]) : 0u, ob = (in && vB) ? nzb(sb[u]) : 0u;
            cA[1] += __popc(oa); cA[0] += __popc(oa & nzb(ma[u])); cA[2] += (in && vA) ? __popc(nzb(sa[u] ^ 0x02020202u)) : 0;
            cB[1] += __popc(ob); cB[0] += __popc(ob & nzb(mb[u])); cB[2] += (in && vB) ? __popc(nzb(sb[u] ^ 0x02020202u)) : 0;
          }
        }
        finish(bA, cA[0], cA[1], cA[2]);
        if (hasB) finish(bB, cB[0], cB[1], cB[2]);
      }
    }
  }
  if (tid == 0) TRF(3, 212);          // this warp's mask counts done
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (tid == 0) TRF(3, 213);          // all counts done
  const uint32_t tm = misc->tmem_base;
  if (warp < 8) {                     // zero the persistent wgrad accumulators (192 columns x 128 lanes)
    const uint32_t zb = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 96;
#pragma unroll
    for (int c = 0; c < 6; ++c) ptx::tmem_st_zero16(zb + c * 16);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid == 0) TRF(3, 200);
  ptx::tc_fence_after();

  uint32_t wpar = 0;                  // weight-barrier parity (one completion per segment)
  uint32_t ph = 0, wph = 0;           // parities of this thread's group barriers: stage commits / deferred MMAs
  bool wb_pending = false;            // deferred MMAs of the previous tile not yet known complete

  // thread roles inside a group
  const int g = warp >> 3, tg = tid & (GT - 1);
  const int p = tg & 127, hsel = tg >> 7;               // point slot (= TMEM lane), column / direction half
  const int quad = (warp & 3);
  const int rw = lane / S, sidx = lane - rw * S;        // ray of this warp, sample index
  const bool lane_used = rw < rpw;
  const int ray_in_tile = quad * rpw + rw;
  const int seg_lo = lane - sidx;                        // first lane of this ray
  unsigned char* act = smem + (g ? SM_ACT1 : SM_ACT0);
  const float* wf = reinterpret_cast<const float*>(smem + SM_W + um::IMG_F32);
  const float* Bd = wf + um::F_DIRS;
  const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
  const uint32_t tA = tm + ACC0 + g * ACC_STRIDE + lane_base, tE = tA + ACC_E;
  Issuer is;
  is.a16 = ptx::smem_u32(act) >> 4;
  is.w16 = ptx::smem_u32(smem + SM_W) >> 4;
  is.tm = tm;
  is.acc = tm + ACC0 + g * ACC_STRIDE;
  const bool issuer_warp = (warp & 7) == 0;
  // loss weights of this thread's current object: (term enabled: no object has an empty mask) / (this object's mask count)
  const float on_d = misc->on[0] ? 1.f : 0.f, on_c = misc->on[1] ? 1.f : 0.f, on_o = misc->on[2] ? 1.f : 0.f;
  float w_d = 0.f, w_c = 0.f, w_o = 0.f;
  bool cnt_ready = !pub_counts;       // published counts acquired?

  // Reduce object b's partial rows over float4 columns [i_lo, i_hi) in segment order (the same sum every run) and either
  // apply AdamW (fuse_adam) or add the reduced gradient into `grads`.  `owner` writes the object's loss terms / status.
  // Called by all threads of the CTA; returns the object's skip flag (loss explosion guard, render_rays.py:88-90).
  auto finish_rows = [&](int b, int i_lo, int i_hi, bool owner) -> int {
    const int c_first = cta_of_pair(rg, G, b * npo), c_last = cta_of_pair(rg, G, (b + 1) * npo - 1);
    const int nseg = c_last - c_first + 1;
    const float* P0 = x.partials + (size_t)(c_first + b) * L.stride;
    const size_t row = (size_t)b * L.stride;
    const int n4 = L.stride >> 2;
    const bool upd = a.backward != 0;
    // (1) the object's loss terms (segment order), the explosion guard and the bias corrections
    if (warp == 0) {
      float s = 0.f;
      if (nseg <= 10) {                                 // up to ten segments are loaded by 30 lanes at once
        const int k = lane / 3, j = lane - 3 * k;
        const float v = (k < nseg) ? __ldcg(P0 + (size_t)k * L.stride + L.P + j) : 0.f;
        for (int kk = 0; kk < nseg; ++kk) s += __shfl_sync(0xffffffffu, v, kk * 3 + (lane < 3 ? lane : 0));
      } else if (lane < 3) {
        for (int k = 0; k < nseg; ++k) s += __ldcg(P0 + (size_t)k * L.stride + L.P + lane);
      }
      const float l_d = __shfl_sync(0xffffffffu, s, 0), l_c = __shfl_sync(0xffffffffu, s, 1), l_o = __shfl_sync(0xffffffffu, s, 2);
      if (lane == 0) {
        const float tot = l_d + a.cs * l_c + a.os * l_o;
        if (owner) { float* lt = a.loss_terms + b * 4; lt[0] = l_d; lt[1] = l_c; lt[2] = l_o; lt[3] = tot; }
        int bad = 0;                                    // render_rays.py:88-90: the reference aborts before the update
        if (x.guard_loss) {
          if (l_d > 100000.f || l_c > 100000.f || l_o > 100000.f) bad |= 1;
          if (!(tot == tot) || fabsf(tot) > 3.0e38f) bad |= 2;
          if (bad && x.status && owner) atomicOr(x.status, bad);
        }
        misc->skip = bad;
        if (x.fuse_adam && x.step_counter) {
          // bias corrections of step t from the host-built table (torch's double-precision scalars, rounded once);
          // beyond the table: 1 - beta^t = -expm1(t ln beta) in fp32 (cancellation-free, <= 3e-7 relative)
          const int t = x.step_counter[b] + 1;
          if (t < x.bc_n) {
            const float2 bc = x.bc_table[t];
            misc->step_size = (float)x.lr / bc.x;
            misc->bc2_sqrt = bc.y;
          } else {
            misc->step_size = (float)x.lr / (-expm1f((float)t * x.log_b1));
            misc->bc2_sqrt = sqrtf(-expm1f((float)t * x.log_b2));
          }
        } else {
          misc->step_size = x.step_size; misc->bc2_sqrt = x.bc2_sqrt;
        }
      }
    }
    __syncthreads();
    const int sk = misc->skip;
    // (2) reduce + update this CTA's columns (normally at most one float4 column per thread).  Deliberately compact
    //     code: it runs once per launch out of a cold instruction cache, where instruction fetch, not data, is the cost
    if (upd && !(sk && x.fuse_adam)) {
      const float step_size = misc->step_size, bc2_sqrt = misc->bc2_sqrt;
      for (int c4 = i_lo + tid; c4 < i_hi; c4 += NT) {
        const float4* src = reinterpret_cast<const float4*>(P0) + c4;
        float4 pp, mm = make_float4(0.f, 0.f, 0.f, 0.f), vv = mm;
        if (x.fuse_adam) {
          pp = *(reinterpret_cast<const float4*>(x.p + row) + c4);
          mm = *(reinterpret_cast<const float4*>(x.m + row) + c4);
          vv = *(reinterpret_cast<const float4*>(x.v + row) + c4);
        } else {
          pp = *(reinterpret_cast<const float4*>(a.grads + row) + c4);
        }
        // segment order = CTA order: the same sum every run; four rows in flight per batch
        float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int k0 = 0; k0 < nseg; k0 += 4) {
          float4 u[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) u[k] = (k0 + k < nseg) ? __ldcg(src + (size_t)(k0 + k) * n4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k0 + k < nseg) { gsum.x += u[k].x; gsum.y += u[k].y; gsum.z += u[k].z; gsum.w += u[k].w; }
        }
        const int e = c4 * 4;
        if (!x.fuse_adam) {
          float4 o = pp;
          o.x += gsum.x; o.y += gsum.y; o.z += gsum.z; o.w += gsum.w;
          if (e + 3 >= L.P) { if (e + 0 >= L.P) o.x = 0.f; if (e + 1 >= L.P) o.y = 0.f; if (e + 2 >= L.P) o.z = 0.f; o.w = 0.f; }
          *(reinterpret_cast<float4*>(a.grads + row) + c4) = o;
          continue;
        }
        // torch.optim.AdamW._single_tensor_adamw, op for op as k_adamw restates it
        float* pj = &pp.x; float* mj = &mm.x; float* vj = &vv.x; const float* gj = &gsum.x;
        __half* img = x.image_out ? x.image_out + (size_t)b * x.img_halves : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (e + j >= L.P) continue;
          float pw = pj[j] * x.lr_wd;
          const float m1 = mj[j] + (gj[j] - mj[j]) * x.one_m_b1;
          const float v1 = vj[j] * x.b2 + (x.one_m_b2 * gj[j]) * gj[j];
          const float denom = sqrtf(v1) / bc2_sqrt + x.eps;
          pw = pw - step_size * (m1 / denom);
          pj[j] = pw; mj[j] = m1; vj[j] = v1;
          if (img) {
            const int t = x.img_index[e + j];
            if (t >= 0) img[t] = __float2half_rn(pw);
            else if (t <= -2) reinterpret_cast<float*>(img)[-(t + 2)] = pw;
          }
        }
        *(reinterpret_cast<float4*>(x.p + row) + c4) = pp;
        *(reinterpret_cast<float4*>(x.m + row) + c4) = mm;
        *(reinterpret_cast<float4*>(x.v + row) + c4) = vv;
      }
    }
    __syncthreads();                                    // the next call's warp 0 overwrites misc->skip
    return sk;
  };

  // The work list is in units of TILE PAIRS (the two point groups of a CTA advance in lock-step rounds, one tile each):
  // every segment of an object is a whole number of rounds -- with a per-tile split a segment boundary inside a CTA
  // cost it an extra round.
  for (long long gt = gt_begin; gt < gt_end;) {
    const int b = (int)(gt / npo);
    const int p0 = (int)(gt - (long long)b * npo);
    const int p1 = (int)min((long long)npo, (long long)p0 + (gt_end - gt));
    gt += p1 - p0;
    const int t0 = 2 * p0, t1 = min(2 * p1, tpo);       // this segment's tiles of object b

    // ---- this object's weight image (one bulk copy global -> shared, issued before the previous segment's flush) ----
    um::mbar_wait_or_trap(&misc->wbar, wpar);
    if (tid == 0) TRF(3, 201);
    wpar ^= 1;

    float ls_d = 0.f, ls_c = 0.f, ls_o = 0.f;
    const int t_step = x.single_group ? 1 : 2;
    const int t_first = x.single_group ? (g == 0 ? t0 : t1) : t0 + g;
    {
    // ===================== compute threads of point group g ============================================================
    const float isc = 1.0f / a.scale[b];
    bool seg_w = false;               // w_d / w_c / w_o hold this segment's object?
    if (!a.fwd_only && !pub_counts) {
      w_d = on_d / ((float)cnt[b * 3 + 0] + 1e-10f);
      w_c = on_c / ((float)cnt[b * 3 + 1] + 1e-10f);
      w_o = on_o / ((float)cnt[b * 3 + 2] + 1e-10f);
      seg_w = true;
    }

    // operands written -> fence for the async proxy -> group barrier -> the elected lane of the group's first warp issues
    // the stage's MMAs, commits, then issues the layer's weight-gradient MMAs behind the commit -> wait for the accumulator
    // FWD_HANDOFF: a forward stage's MMAs read their A operand from tensor memory (tcgen05.st by the epilogue, no
    // generic->async proxy fence on the chain); the shared-memory copies written beside it are first read by
    // weight-gradient MMAs after stage 6's fence.
#if VMB_TS_FWD
#define FWD_HANDOFF() ptx::tmem_st_wait()
#else
#define FWD_HANDOFF() ptx::fence_async_smem()
#endif
#define STAGE_BEGIN(ST)                                \
  do {                                                 \
    if ((ST) < 6) FWD_HANDOFF(); else ptx::fence_async_smem(); \
    ptx::tc_fence_before();                            \
    group_bar(g);                                      \
    TRG();                                             \
    if (issuer_warp) {                                 \
      ptx::tc_fence_after();                           \
      if (ptx::elect_one()) {                          \
        is.stage(ST);                                  \
        ptx::umma_commit(&misc->done[g]);              \
        is.post(ST);                                   \
      }                                                \
      __syncwarp();                                    \
    }                                                  \
  } while (0)
#define STAGE_END()                                    \
  do {                                                 \
    um::mbar_wait_or_trap(&misc->done[g], ph);         \
    ph ^= 1;                                           \
    ptx::tc_fence_after();                             \
    TRG();                                             \
  } while (0)
#define STAGE(ST) do { STAGE_BEGIN(ST); STAGE_END(); } while (0)
    // hidden-layer epilogue on this thread's 16 columns: acc + bias -> ReLU -> fp16 (2 x 16 B)
#define EPI_RELU(BIAS_OFF, FG)                                                                 \
  do {                                                                                         \
    float v[16];                                                                               \
    ptx::tmem_ld16(tA + 16 * hsel, v);                                                         \
    const float4* bp = reinterpret_cast<const float4*>(wf + (BIAS_OFF) + 16 * hsel);           \
    const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];                               \
    ptx::tmem_ld_wait();                                                                       \
    uint32_t hh[8];                                                                            \
    hh[0] = pack_relu_h2(v[0] + b0.x, v[1] + b0.y);   hh[1] = pack_relu_h2(v[2] + b0.z, v[3] + b0.w);     \
    hh[2] = pack_relu_h2(v[4] + b1.x, v[5] + b1.y);   hh[3] = pack_relu_h2(v[6] + b1.z, v[7] + b1.w);     \
    hh[4] = pack_relu_h2(v[8] + b2.x, v[9] + b2.y);   hh[5] = pack_relu_h2(v[10] + b2.z, v[11] + b2.w);   \
    hh[6] = pack_relu_h2(v[12] + b3.x, v[13] + b3.y); hh[7] = pack_relu_h2(v[14] + b3.z, v[15] + b3.w);   \
    if (VMB_TS_FWD) ptx::tmem_st8(tA + TC_HA + 8 * hsel, hh);      /* next layer's A operand */              \
    uint4* dst = reinterpret_cast<uint4*>(act + ((FG) + 2 * hsel) * FGB + p * 16);   /* wgrad's copy */      \
    dst[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]);                                           \
    dst[128] = make_uint4(hh[4], hh[5], hh[6], hh[7]);                                         \
  } while (0)
    // dgrad epilogue: dY = (h > 0) * acc; h is read from its own block (FG_H), dY goes to a block whose readers retired
#define EPI_DGRAD(FG_H, FG_OUT)                                                                \
  do {                                                                                         \
    float v[16];                                                                               \
    ptx::tmem_ld16(tA + 16 * hsel, v);                                                         \
    const uint4* hsrc = reinterpret_cast<const uint4*>(act + ((FG_H) + 2 * hsel) * FGB + p * 16); \
    const uint4 o0 = hsrc[0], o1 = hsrc[128];                                                  \
    ptx::tmem_ld_wait();                                                                       \
    uint4* dst = reinterpret_cast<uint4*>(act + ((FG_OUT) + 2 * hsel) * FGB + p * 16);         \
    dst[0] = make_uint4(um::gate_h2(um::pack_h2(v[0], v[1]), o0.x), um::gate_h2(um::pack_h2(v[2], v[3]), o0.y),       \
                        um::gate_h2(um::pack_h2(v[4], v[5]), o0.z), um::gate_h2(um::pack_h2(v[6], v[7]), o0.w));      \
    dst[128] = make_uint4(um::gate_h2(um::pack_h2(v[8], v[9]), o1.x), um::gate_h2(um::pack_h2(v[10], v[11]), o1.y),   \
                          um::gate_h2(um::pack_h2(v[12], v[13]), o1.z), um::gate_h2(um::pack_h2(v[14], v[15]), o1.w)); \
  } while (0)

    // prefetched inputs of the next tile (global-load latency overlaps the current tile)
    float nx = 0.f, ny = 0.f, nz = 0.f, nzv = 0.f, n_gd = 0.f, n_c0 = 0.f, n_c1 = 0.f, n_c2 = 0.f;
    // (nothing computes on the loaded values here: the first use of a load stalls the warp for the memory latency)
    int n_sem = 0, n_msk = 0;
    bool n_live = false;
    // this object's base pointers, formed once per segment (the kernel-parameter loads and 64-bit stride products
    // otherwise sit, with their latencies, in every tile's prefetch)
    const float* pcs_b = a.pcs + (size_t)b * a.pcs_stride + (size_t)sidx * 3;
    const float* z_b = a.z + (size_t)b * a.z_stride + sidx;
    const float* gd_b = a.gt_depth + (size_t)b * a.gt_depth_stride;
    const float* gc_b = a.gt_colour + (size_t)b * a.gt_colour_stride;
    const unsigned char* sem_b = a.sem + (size_t)b * a.sem_stride;
    const unsigned char* msk_b = a.mask + (size_t)b * a.mask_stride;
    const bool want_targets = hsel == 0 && !a.fwd_only;
    auto prefetch = [&](int t) {
      nx = ny = nz = nzv = 0.f; n_gd = n_c0 = n_c1 = n_c2 = 0.f; n_sem = 0; n_msk = 0; n_live = false;
      if (t >= t1 || !lane_used) return;
      const int ray = t * nr + ray_in_tile;
      if (ray >= R) return;
      const size_t pi = (size_t)ray * S;
      const float* pp = pcs_b + pi * 3;
      nx = pp[0]; ny = pp[1]; nz = pp[2];
      if (want_targets) {                               // the render / loss lanes: every lane of a ray reads the ray's targets
        nzv = z_b[pi];
        n_gd = gd_b[ray];
        const float* gcp = gc_b + (size_t)ray * 3;
        n_c0 = gcp[0]; n_c1 = gcp[1]; n_c2 = gcp[2];
        n_sem = sem_b[ray];
        n_msk = msk_b[ray];
      }
      n_live = true;                                      // live point
    };
    prefetch(t_first);

    int trs = 0;
#define TRG() do { if (tg == 0) { TRF(g, trs); ++trs; } } while (0)
    for (int t = t_first; t < t1; t += t_step) {
      const int ray = t * nr + ray_in_tile;
      TRG();                                            // tile start
      // ---- E0: positional embedding (embedding.py:82-91) ----------------------------------------------------
      const float t0x = nx * isc, t1x = ny * isc, t2x = nz * isc, zv = nzv;
      const uint64_t tp0 = um::pk2(t0x, t0x), tp1 = um::pk2(t1x, t1x), tp2 = um::pk2(t2x, t2x);
      const float gd = n_gd, gc0 = n_c0, gc1 = n_c1, gc2 = n_c2;
      const int sv = n_sem, mv = n_msk;
      const bool live = n_live;
      [[maybe_unused]] int tre = 16 * ((t - t_first) / t_step);
      TRE();
      prefetch(t + t_step);
      TRE();
      {
        uint4* e1 = reinterpret_cast<uint4*>(act + FG_E1 * FGB + p * 16);
        uint4* e2 = reinterpret_cast<uint4*>(act + FG_E2 * FGB + p * 16);
        const int q0 = hsel ? 3 : 0, q1 = hsel ? 5 : 3;
        // software pipeline over the 4-direction chunks: the NEXT chunk's projections, range reduction and MUFU
        // sin / cos are issued before the CURRENT chunk's doubling recurrence, which hides their latency
        uint64_t s01, s23, c01, c23;
        {
          uint64_t pj01, pj23;
          um::project4(Bd, q0, tp0, tp1, tp2, pj01, pj23);
          um::sincos4_x2(pj01, pj23, s01, s23, c01, c23);
        }
#pragma unroll 1
        for (int q = q0; q < q1; ++q) {                // directions 4q .. 4q+3
          float sv[4][6];
          uint64_t ns01 = 0, ns23 = 0, nc01 = 0, nc23 = 0;
          if (q + 1 < q1) {
            uint64_t pj01, pj23;
            um::project4(Bd, q + 1, tp0, tp1, tp2, pj01, pj23);
            um::sincos4_x2(pj01, pj23, ns01, ns23, nc01, nc23);
          }
          um::sin_doubling4_x2(s01, s23, c01, c23, sv);
          s01 = ns01; s23 = ns23; c01 = nc01; c23 = nc23;
          TRE();
          if (q == q0 && wb_pending) {                  // the previous tile's deferred MMAs still read E1 / DPR / FC3
            um::mbar_wait_or_trap(&misc->wb[g], wph); wph ^= 1; wb_pending = false;
          }
          const uint4 ua = make_uint4(um::pack_h2(sv[0][0], sv[0][1]), um::pack_h2(sv[0][2], sv[0][3]), um::pack_h2(sv[1][0], sv[1][1]), um::pack_h2(sv[1][2], sv[1][3]));
          const uint4 ub = make_uint4(um::pack_h2(sv[2][0], sv[2][1]), um::pack_h2(sv[2][2], sv[2][3]), um::pack_h2(sv[3][0], sv[3][1]), um::pack_h2(sv[3][2], sv[3][3]));
          const uint4 uc = make_uint4(um::pack_h2(sv[0][4], sv[0][5]), um::pack_h2(sv[1][4], sv[1][5]), um::pack_h2(sv[2][4], sv[2][5]), um::pack_h2(sv[3][4], sv[3][5]));
          e1[(2 * q + 1) * 128] = ua; e1[(2 * q + 2) * 128] = ub; e2[q * 128] = uc;
          if (VMB_TS_FWD) {       // the same 8-feature chunks as A-operand columns (4 packed columns per chunk)
            ptx::tmem_st4(tA + TC_E1A + (2 * q + 1) * 4, ua.x, ua.y, ua.z, ua.w);
            ptx::tmem_st4(tA + TC_E1A + (2 * q + 2) * 4, ub.x, ub.y, ub.z, ub.w);
            ptx::tmem_st4(tA + TC_E2A + q * 4, uc.x, uc.y, uc.z, uc.w);
          }
          TRE();
        }
        if (hsel) {
          // direction 20 shares chunk 0 of emb1 with [1, x, y, z] and chunk 5 of emb2 with the const-1 column
          float s[6];
          um::sin_ladder(fmaf(Bd[2 * um::DIRS_PITCH + 20], t2x, fmaf(Bd[um::DIRS_PITCH + 20], t1x, Bd[20] * t0x)), s);
          const uint4 u0 = make_uint4(um::pack_h2(1.0f, t0x), um::pack_h2(t1x, t2x), um::pack_h2(s[0], s[1]), um::pack_h2(s[2], s[3]));
          const uint4 u5 = make_uint4(um::pack_h2(s[4], s[5]), um::pack_h2(1.0f, 0.f), 0u, 0u);
          e1[0] = u0;
          e2[5 * 128] = u5;
          e1[11 * 128] = make_uint4(0u, 0u, 0u, 0u);
          if (VMB_TS_FWD) {
            ptx::tmem_st4(tA + TC_E1A, u0.x, u0.y, u0.z, u0.w);
            ptx::tmem_st4(tA + TC_E2A + 20, u5.x, u5.y, u5.z, u5.w);
            ptx::tmem_st4(tA + TC_E1A + 44, 0u, 0u, 0u, 0u);
          }
          // zero this point's dhead row (cols 4..15 stay zero; 0..3 are written after the render)
          uint4* dh = reinterpret_cast<uint4*>(act + FG_DH * FGB + p * 16);
          dh[0] = make_uint4(0u, 0u, 0u, 0u); dh[128] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      TRG();                                            // E0 done
      STAGE(0);                                         // in_layer
      EPI_RELU(um::F_BIN, FG_FC1);
      TRG();
      STAGE(1);                                         // mid1
      EPI_RELU(um::F_BM1, FG_FC2);
      TRG();
      STAGE(2);                                         // cat_layer
      EPI_RELU(um::F_BCAT, FG_FC3);
      TRG();
      STAGE(3);                                         // mid2
      EPI_RELU(um::F_BM2, FG_FC4);
      TRG();
      STAGE(4);                                         // color_linear + out_alpha
      EPI_RELU(um::F_BCL, FG_HC);
      TRG();
      STAGE_BEGIN(5);                                   // out_color runs while the alpha-only part of the render is computed
      // ---- heads + volume render + losses + ray gradients, all in registers of the hsel == 0 warps ---------------
      // rays never straddle a warp: the scans over the sample axis are warp shuffles
      auto raysum = [&](float v) {                      // per-ray sum: guarded tree reduction to the ray's first lane, broadcast
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          if (off < S) { const float u = __shfl_down_sync(0xffffffffu, v, off); if (sidx + off < S) v += u; }
        }
        return __shfl_sync(0xffffffffu, v, seg_lo);
      };
      float araw = 0.f, occ = 0.f, fr = 1.f, Tr = 1.f, w = 0.f, D = 0.f, O = 0.f, V = 0.f;
      if (hsel == 0) {                                  // alpha's tile was complete with stage 4
        if (!seg_w && pub_counts) {                     // cooperative launch: acquire the published mask counts (first render only)
          if (!cnt_ready) {
            const volatile unsigned int* pubd = gbar + 2;
            unsigned int spins = 0;
            while (*pubd < (unsigned int)a.B) { if (++spins > 2000000000u) __trap(); }
            __threadfence();
            cnt_ready = true;
          }
          const volatile unsigned int* ef = gbar + 3;
          const volatile int* cp = x.counts_pub + b * 4;
          w_d = (ef[0] ? 0.f : 1.f) / ((float)cp[0] + 1e-10f);
          w_c = (ef[1] ? 0.f : 1.f) / ((float)cp[1] + 1e-10f);
          w_o = (ef[2] ? 0.f : 1.f) / ((float)cp[2] + 1e-10f);
          seg_w = true;
        }
        float hv[8];
        um::tmem_ld8(tA + TC_ALPHA, hv);
        ptx::tmem_ld_wait();
        araw = (hv[0] + wf[um::F_BA]) * 10.0f;                                   // model.py:77
        occ = um::fast_sigmoid(araw);                                            // render_rays.py:6
        if (!a.fwd_only) {
          if (!live) occ = 0.f;
          // termination: T_s = prod_{j<s} (1 - occ_j + 1e-10) (render_rays.py:29), inclusive scan then shift
          fr = 1.f - occ + 1e-10f;
          float inc = fr;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            if (off < S) { const float u = __shfl_up_sync(0xffffffffu, inc, off); if (sidx >= off) inc *= u; }
          }
          Tr = __shfl_up_sync(0xffffffffu, inc, 1);
          if (sidx == 0) Tr = 1.f;
          w = occ * Tr;                                                           // render_rays.py:34
          D = raysum(w * zv); O = raysum(w);
          const float dz = zv - D;
          V = raysum(w * dz * dz);                                                // render_rays.py:47-51 (detached)
        }
      }
      STAGE_END();
      if (hsel == 0) {
        float hv[8];
        um::tmem_ld8(tA + TC_COL, hv);
        ptx::tmem_ld_wait();
        const float c0 = um::fast_sigmoid(hv[1] + wf[um::F_BOC + 0]), c1 = um::fast_sigmoid(hv[2] + wf[um::F_BOC + 1]),
                    c2 = um::fast_sigmoid(hv[3] + wf[um::F_BOC + 2]);
        if (a.fwd_only) {                               // eval_points (trainer.py:77-90): raw alpha*10, sigmoid colour per point
          if (live) {
            const size_t n = (size_t)ray * S + sidx;
            a.out_alpha[(size_t)b * a.alpha_stride + n] = araw;
            float* oc = a.out_colour + (size_t)b * a.colour_stride + n * 3;
            oc[0] = c0; oc[1] = c1; oc[2] = c2;
          }
        } else {
          const float C0 = raysum(w * c0), C1 = raysum(w * c1), C2 = raysum(w * c2);
          const float m_o = (live && sv != 0) ? 1.f : 0.f, m_s = (live && sv != 2) ? 1.f : 0.f, m_d = (mv != 0) ? m_o : 0.f;
          const float info = 1.f / (sqrtf(V) + 1e-4f);                            // render_rays.py:74-79
          const float e_d = D - gd, e_o = O - m_o, e_c0 = C0 - gc0, e_c1 = C1 - gc1, e_c2 = C2 - gc2;
          if (live && sidx == 0) {
            if (a.r_depth) a.r_depth[(size_t)b * R + ray] = D;
            if (a.r_var) a.r_var[(size_t)b * R + ray] = V;
            if (a.r_opacity) a.r_opacity[(size_t)b * R + ray] = O;
            if (a.r_colour) { float* rc = a.r_colour + ((size_t)b * R + ray) * 3; rc[0] = C0; rc[1] = C1; rc[2] = C2; }
            ls_d += w_d * fabsf(e_d) * m_d * info;
            ls_c += w_c * (fabsf(e_c0) + fabsf(e_c1) + fabsf(e_c2)) * m_o;
            ls_o += w_o * fabsf(e_o) * m_s;
          }
          if (a.backward) {
            const float gD = LS * w_d * vmb_sign(e_d) * m_d * info;
            const float kc = LS * w_c * a.cs * m_o;
            const float gC0 = kc * vmb_sign(e_c0), gC1 = kc * vmb_sign(e_c1), gC2 = kc * vmb_sign(e_c2);
            const float gO = LS * w_o * a.os * vmb_sign(e_o) * m_s;
            // d(loss)/d(occ_s) through the termination product: G_s T_s - (sum_{k>s} G_k w_k) / (1 - occ_s + 1e-10)
            const float Gs = fmaf(gD, zv, fmaf(gC0, c0, fmaf(gC1, c1, fmaf(gC2, c2, gO))));
            float suf = Gs * w;                                                   // inclusive suffix scan
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
              if (off < S) { const float u = __shfl_down_sync(0xffffffffu, suf, off); if (sidx + off < S) suf += u; }
            }
            float sx = __shfl_down_sync(0xffffffffu, suf, 1);
            if (sidx + 1 >= S) sx = 0.f;
            const float docc = Gs * Tr - __fdividef(sx, fr);
            const float da = 10.0f * docc * occ * (1.f - occ);                    // model.py:77
            if (live)
              *reinterpret_cast<uint2*>(act + FG_DH * FGB + p * 16) =
                  make_uint2(um::pack_h2(da, gC0 * w * c0 * (1.f - c0)), um::pack_h2(gC1 * w * c1 * (1.f - c1), gC2 * w * c2 * (1.f - c2)));
          }
        }
      }
      TRG();                                            // heads + render done
      if (a.fwd_only || !a.backward) { group_bar(g); continue; }
      STAGE(6);                                         // d_hc            (+ heads wgrad behind the commit)
      EPI_DGRAD(FG_HC, FG_Z);                           // dYc -> Z
      TRG();
      STAGE(7);                                         // d_fc4           (+ wgrad color_linear)
      EPI_DGRAD(FG_FC4, FG_HC);                         // dY4 -> HC
      TRG();
      STAGE(8);                                         // d_fc3           (+ wgrad mid2)
      EPI_DGRAD(FG_FC3, FG_FC4);                        // dY3 -> FC4
      TRG();
      STAGE(9);                                         // d_fc2           (+ d_emb1 part 1, wgrad cat_layer)
      EPI_DGRAD(FG_FC2, FG_HC);                         // dY2 -> HC
      TRG();
      STAGE(10);                                        // d_fc1           (+ wgrad mid1)
      EPI_DGRAD(FG_FC1, FG_FC3);                        // dY1 -> FC3
      TRG();
      STAGE(11);                                        // d_emb1 part 2 -> E, d_emb2 -> A[0..48)   (+ wgrad in_layer)
      // ---- PE backward: dproj_d = pi * sum_k 2^k g_{k,d} cos(pi 2^k proj_d), written as an fp16 block --------------
      {
        const int q0 = hsel ? 3 : 0, q1 = hsel ? 5 : 3;
        uint64_t c01, c23;
        {
          uint64_t pj01, pj23;
          um::project4(Bd, q0, tp0, tp1, tp2, pj01, pj23);
          um::cos4_x2(pj01, pj23, c01, c23);
        }
#pragma unroll 1
        for (int q = q0; q < q1; ++q) {
          float g1a[8], g1b[8], g2[8];
          um::tmem_ld8(tE + 16 * q + 8, g1a);          // emb1 cols of directions 4q, 4q+1 (k = 0..3)
          um::tmem_ld8(tE + 16 * q + 16, g1b);         //                          4q+2, 4q+3
          um::tmem_ld8(tA + 8 * q, g2);                // emb2 cols (k = 4, 5)
          float cv[4][6], dp[4];
          uint64_t nc01 = 0, nc23 = 0;
          if (q + 1 < q1) {                            // next chunk's front half before this chunk's recurrence
            uint64_t pj01, pj23;
            um::project4(Bd, q + 1, tp0, tp1, tp2, pj01, pj23);
            um::cos4_x2(pj01, pj23, nc01, nc23);
          }
          um::cos_doubling4_x2(c01, c23, cv);
          c01 = nc01; c23 = nc23;
          ptx::tmem_ld_wait();
#pragma unroll
          for (int dd = 0; dd < 4; ++dd) {
            const float* g1 = (dd < 2) ? (g1a + dd * 4) : (g1b + (dd - 2) * 4);
            float d = g1[0] * cv[dd][0];
            d = fmaf(2.f * g1[1], cv[dd][1], d);
            d = fmaf(4.f * g1[2], cv[dd][2], d);
            d = fmaf(8.f * g1[3], cv[dd][3], d);
            d = fmaf(16.f * g2[dd * 2], cv[dd][4], d);
            d = fmaf(32.f * g2[dd * 2 + 1], cv[dd][5], d);
            dp[dd] = d * VMB_PI_F;
          }
          // directions 4q..4q+3 = columns (4q)%8.. of feature group q/2
          *reinterpret_cast<uint2*>(act + (FG_DPR + (q >> 1)) * FGB + p * 16 + (q & 1) * 8) =
              make_uint2(um::pack_h2(dp[0], dp[1]), um::pack_h2(dp[2], dp[3]));
        }
        if (hsel) {                                     // direction 20: emb1 cols 4..7, emb2 cols 40, 41
          float g1[8], g2[8], c[6];
          um::tmem_ld8(tE, g1);
          um::tmem_ld8(tA + 40, g2);
          um::cos_ladder(fmaf(Bd[2 * um::DIRS_PITCH + 20], t2x, fmaf(Bd[um::DIRS_PITCH + 20], t1x, Bd[20] * t0x)), c);
          ptx::tmem_ld_wait();
          float d = g1[4] * c[0];
          d = fmaf(2.f * g1[5], c[1], d); d = fmaf(4.f * g1[6], c[2], d); d = fmaf(8.f * g1[7], c[3], d);
          d = fmaf(16.f * g2[0], c[4], d); d = fmaf(32.f * g2[1], c[5], d);
          *reinterpret_cast<uint2*>(act + (FG_DPR + 2) * FGB + p * 16 + 8) = make_uint2(um::pack_h2(d * VMB_PI_F, 0.f), 0u);
        }
      }
      TRG();                                            // PE backward done
      // dB (+)= dproj^T [x y z] as a wgrad MMA; its commit also covers wgrad in_layer issued behind stage 11
      ptx::fence_async_smem();
      ptx::tc_fence_before();
      group_bar(g);
      if (issuer_warp) {
        ptx::tc_fence_after();
        if (ptx::elect_one()) { is.dirs_wgrad(); ptx::umma_commit(&misc->wb[g]); }
        __syncwarp();
      }
      wb_pending = true;
      TRG();
    }
#undef TRG
#undef STAGE
#undef EPI_RELU
#undef EPI_DGRAD
    if (wb_pending) { um::mbar_wait_or_trap(&misc->wb[g], wph); wph ^= 1; wb_pending = false; }
    }   // compute threads
    if (a.fwd_only) {                                   // nothing to reduce; the weight buffer is free once both groups are done
      __syncthreads();
      if (tid == 0 && gt < gt_end) {
        ptx::mbar_arrive_expect_tx(&misc->wbar, um::IMG_BYTES);
        ptx::bulk_g2s(smem + SM_W, image + (size_t)(gt / npo) * um::IMG_BYTES, um::IMG_BYTES, &misc->wbar);
      }
      continue;
    }
    // ---- segment end: this CTA's partial sums for object b -----------------------------------------------------
    ls_d = warp_sum(ls_d); ls_c = warp_sum(ls_c); ls_o = warp_sum(ls_o);
    if (hsel == 0 && lane == 0) { float* l = misc->lsum[g * 4 + quad]; l[0] = ls_d; l[1] = ls_c; l[2] = ls_o; }
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) TRF(3, 202);
    ptx::tc_fence_after();
    if (tid == 0 && gt < gt_end) {          // every MMA of this segment has completed: the weight buffer is free --
      ptx::mbar_arrive_expect_tx(&misc->wbar, um::IMG_BYTES);        // the next object's image lands during the flush
      ptx::bulk_g2s(smem + SM_W, image + (size_t)(gt / npo) * um::IMG_BYTES, um::IMG_BYTES, &misc->wbar);
    }
    // the segment's gradient row is assembled in shared memory (group 0's activation region is idle now: the scattered
    // 4-byte stores of the accumulator -> parameter-index mapping cost nothing there) and leaves as coalesced 16-byte stores
    float* Pr = reinterpret_cast<float*>(smem + SM_ACT0);
    if (tid < 3) {                                      // fixed summation order -> reproducible loss terms
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += misc->lsum[w8][tid];
      Pr[L.P + tid] = s;
    }
    if (a.backward && warp < 8) {
      // TMEM -> registers -> this segment's row of the partial block (plain stores, no atomics)
      const int q = warp & 3, half = warp >> 2, ln = q * 32 + lane;
#pragma unroll 1
      for (int cc = 0; cc < 3; ++cc) {
        const int blk = half * 3 + cc;
        float v[32];
        ptx::tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + blk * 32, v);
        ptx::tmem_ld_wait();
        if (blk < 5) {
          int ld;
          const int base = wg_base(L, blk, ln, ld);
          if (base >= 0) {
            float* gp = Pr + base;
#pragma unroll
            for (int j = 0; j < 32; ++j) gp[j * ld] = v[j] * INV_LS;
          }
        } else {    // heads tile: A = [fc4 | emb2 | hc], B = dhead (col 0 = alpha, 1..3 = colour); then the dB tile
          if (ln < 32) {
            Pr[L.o_Wa + ln] = v[0] * INV_LS;
          } else if (ln == 32 + 42) {                   // emb2's constant-1 column: all four head biases
            Pr[L.o_ba] = v[0] * INV_LS;
#pragma unroll
            for (int c = 0; c < 3; ++c) Pr[L.o_boc + c] = v[1 + c] * INV_LS;
          } else if (ln >= 80 && ln < 112) {
#pragma unroll
            for (int c = 0; c < 3; ++c) Pr[L.o_Woc + c * 32 + (ln - 80)] = v[1 + c] * INV_LS;
          }
          if (ln < VMB_NDIRS) {                         // dB[d][i]: A = dproj block, B = emb1 cols [1, x, y, z, ...]
#pragma unroll
            for (int i = 0; i < 3; ++i) Pr[L.o_B + ln * 3 + i] = v[16 + 1 + i] * INV_LS;
          }
        }
        ptx::tmem_st_zero16(tm + ((uint32_t)(q * 32) << 16) + blk * 32);       // ready for the next object
        ptx::tmem_st_zero16(tm + ((uint32_t)(q * 32) << 16) + blk * 32 + 16);
      }
      ptx::tmem_st_wait();
    }
    __syncthreads();
    {
      float4* dst = reinterpret_cast<float4*>(x.partials + (size_t)(blockIdx.x + b) * L.stride);
      const float4* src = reinterpret_cast<const float4*>(Pr);
      for (int i = tid; i < (L.stride >> 2); i += NT) dst[i] = src[i];
    }
    // ---- reduce + update ---------------------------------------------------------------------------------------------
    // cooperative launch: after the LAST segment (below, outside this loop), all CTAs share the reduction of all objects.
    // otherwise: the last segment of the object to arrive reduces that object's rows here.
    if (!x.cooperative) {
      __threadfence();
      __syncthreads();
      const int c_first = cta_of_pair(rg, G, b * npo), c_last = cta_of_pair(rg, G, (b + 1) * npo - 1);
      if (tid == 0) misc->fin = (atomicAdd(&x.obj_done[b], 1u) + 1u == (unsigned int)(c_last - c_first + 1)) ? 1 : 0;
      __syncthreads();
      if (misc->fin) {
        __threadfence();
        const int skip = finish_rows(b, 0, L.stride >> 2, true);
        if (tid == 0) {
          x.obj_done[b] = 0u;
          if (x.fuse_adam && x.step_counter && a.backward && !skip) x.step_counter[b] += 1;
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) TRF(3, 203);
    ptx::tc_fence_after();
  }

  if (x.cooperative && !a.fwd_only) {
    // ---- grid-wide finish: every CTA is resident (cooperative launch), all of them end their tiles within one round
    // of each other, and the reduction of ALL objects' partial rows + AdamW is split evenly over the grid (less than one
    // float4 of the parameter block per thread) instead of one object per SM on the kernel's tail.
    // gbar[0] arrive, gbar[1] depart;  x.obj_done[B + b] = skip flag of object b
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      atomicAdd(&gbar[0], 1u);
      unsigned int spins = 0;
      while (atomicAdd(&gbar[0], 0u) < (unsigned int)G) { __nanosleep(32); if (++spins > 400000000u) __trap(); }
    }
    __syncthreads();
    __threadfence();
    if (tid == 0) TRF(3, 205);
    const int n4 = L.stride >> 2;
    if (tid == 0) TRF(3, 206);
    const long long W = (long long)a.B * n4;
    const long long lo = (W * blockIdx.x) / G, hi = (W * (blockIdx.x + 1)) / G;
    for (int b = (int)(lo / n4); (long long)b * n4 < hi; ++b) {
      const int i_lo = (int)(max(lo, (long long)b * n4) - (long long)b * n4);
      const int i_hi = (int)(min(hi, (long long)(b + 1) * n4) - (long long)b * n4);
      const int skip = finish_rows(b, i_lo, i_hi, i_lo == 0);
      if (i_lo == 0 && tid == 0) x.obj_done[a.B + b] = (unsigned int)skip;
    }
    if (tid == 0) TRF(3, 207);          // this CTA's slices reduced + updated
    __threadfence();
    __syncthreads();
    if (tid == 0) TRF(3, 208);          // fenced
    if (tid == 0) misc->fin = (atomicAdd(&gbar[1], 1u) + 1u == (unsigned int)G) ? 1 : 0;
    __syncthreads();
    if (misc->fin) {                                    // the last CTA to leave: step numbers, re-arm the barrier
      __threadfence();
      if (x.fuse_adam && x.step_counter && a.backward)
        for (int b = tid; b < a.B; b += NT) if (x.obj_done[a.B + b] == 0u) x.step_counter[b] += 1;
      if (x.loss_sum && warp == 0) {                    // scalar loss of the step (loss.py:59-62), fixed summation order
        float s = 0.f;
        for (int b = lane; b < a.B; b += 32) s += __ldcg(a.loss_terms + b * 4 + 3);
        s = warp_sum(s);
        if (lane == 0) *x.loss_sum = s;
      }
      __syncthreads();
      if (tid == 0) { gbar[0] = 0u; gbar[1] = 0u; gbar[2] = 0u; gbar[3] = 0u; gbar[4] = 0u; gbar[5] = 0u; }
    }
  }

  if (tid == 0) TRF(3, 204);
  if (warp == 15) ptx::tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int fused_rows_needed(int n_obj, int n_sm) { return n_obj + n_sm; }

// Cost-aware split of B objects x npo tile pairs over G CTAs.  Starting a second object inside a CTA costs it a flush
// of the gradient partial, a weight load and a pipeline refill (~0.4 of a round measured), so pairs are laid out on a
// virtual axis x(p) = p + beta * (object of p) -- every object boundary is a gap of beta -- and that axis is split
// evenly: a CTA whose range crosses a boundary gets correspondingly fewer pairs.  beta is the largest of {1, 0.7, 0.4, 0}
// that does not raise the maximum number of rounds per CTA.
static void fused_partition(int B, int npo, int G, uf::Ranges& rg) {
  const long long T = (long long)B * npo;
  double beta = 0.0;
  if (T >= 2LL * G) {
    const long long rounds = (T + G - 1) / G;
    for (double cand : {1.0, 0.7, 0.4}) {
      if ((long long)std::ceil(((double)T + cand * (B - 1)) / G - 1e-9) <= rounds) { beta = cand; break; }
    }
  }
  const double V = (double)T + beta * (B - 1);
  int c = 0;
  rg.begin[0] = 0;
  for (long long p = 0; p < T; ++p) {
    const double x = (double)p + beta * (double)(p / npo);
    while (c < G - 1 && x >= V * (c + 1) / G && p > rg.begin[c] && T - p >= G - 1 - c) rg.begin[++c] = (int)p;
    if (T - p - 1 == G - 1 - c && c < G - 1) {            // one pair left per remaining CTA
      for (long long q = p + 1; q < T; ++q) rg.begin[++c] = (int)q;
      break;
    }
  }
  while (c < G) rg.begin[++c] = (int)T;
}

static int fused_launch_step(const VmbLayout& L, const StepParams& sp, const FusedExtra& fx, const void* image, int n_sm,
                             cudaStream_t st, std::string& err, bool* cooperative = nullptr) {
  using namespace uf;
  if (L.H != 32 || L.nfreq != 6) { err = "fused step kernel: hidden must be 32 and n_freq 6"; return -4; }
  if (sp.S < 1 || sp.S > 32) { err = "fused step kernel: n_samples must be in [1, 32]"; return -4; }
  if (!sp.fwd_only && sp.B > MAX_OBJ_SMEM) { err = "fused step kernel: too many objects for the in-kernel mask counts"; return -4; }
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int smem_bytes = SM_CNT + (sp.fwd_only ? 0 : sp.B * 12) + 16;
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_step_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_fused<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_fused<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_step_fused): ") + cudaGetErrorString(e); return -2; }
    attr_set[dev & 63] = true;
  }
  const int rpw = 32 / sp.S;                  // whole rays per warp: the sample axis never crosses a warp
  const int nr = 4 * rpw;
  const int tpo = (sp.R + nr - 1) / nr;       // tiles per object
  const int npo = (tpo + 1) / 2;              // rounds (tile pairs) per object
  const long long T = (long long)npo * sp.B;
  if (T > 0x7fffffffLL) { err = "fused step kernel: too many tiles"; return -1; }
  long long grid = T;
  if (grid > n_sm) grid = n_sm;
  if (grid > MAX_CTAS) grid = MAX_CTAS;
  if (grid < 1) grid = 1;
  Ranges rg;
  fused_partition(sp.B, npo, (int)grid, rg);
  const unsigned char* img = (const unsigned char*)image;
  FusedExtra fxl = fx;
  // every CTA is resident (grid <= #SMs, one CTA per SM) -- the cooperative attribute makes the runtime guarantee it,
  // which is what lets an object's CTAs wait for each other in the shared reduction
  static int coop_ok[64] = {};
  if (coop_ok[dev & 63] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
    coop_ok[dev & 63] = v ? 1 : -1;
  }
  fxl.cooperative = (coop_ok[dev & 63] == 1 && !sp.fwd_only && getenv("VMB_NO_COOP") == nullptr) ? 1 : 0;
  if (cooperative) *cooperative = fxl.cooperative != 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = fxl.cooperative;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e;
  if (sp.S == 10)      e = cudaLaunchKernelEx(&cfg, k_step_fused<10>, sp, fxl, L, img, rg, tpo, npo, nr, rpw);
  else if (sp.S == 14) e = cudaLaunchKernelEx(&cfg, k_step_fused<14>, sp, fxl, L, img, rg, tpo, npo, nr, rpw);
  else                 e = cudaLaunchKernelEx(&cfg, k_step_fused<0>, sp, fxl, L, img, rg, tpo, npo, nr, rpw);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { err = std::string("k_step_fused launch: ") + cudaGetErrorString(e); return -2; }
  return 0;
}
