// placeholder until the tcgen05 kernel lands
#pragma once
#include "common.cuh"
#include "k_step_fp32.cuh"
#include <string>
#define UMMA_MAX_S 16
static int umma_image_bytes() { return 32768; }
static void umma_fill_image_index(const VmbLayout& L, int* idx) { for (int i = 0; i < L.P; ++i) idx[i] = -1; }
static int umma_launch_step(const VmbLayout&, const StepParams&, const void*, cudaStream_t, std::string& err) {
  err = "UMMA step kernel not built"; return -4;
}
