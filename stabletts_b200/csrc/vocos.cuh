// Row kernels of the vocoder hand-off (vocos.cu); orchestration in vocos_api.cu.
#pragma once
#include "common.cuh"

namespace st {

struct DwLnArgs {
    const float* x = nullptr;        // (B, T, C) fp32
    const float* dw_w = nullptr;     // [7][C] depthwise taps (null: LayerNorm only)
    const float* dw_b = nullptr;     // [C]
    const float* ln_w = nullptr; const float* ln_b = nullptr;
    float* out_f32 = nullptr; bf16* out_hi = nullptr; bf16* out_lo = nullptr;
    int B = 0, T = 0, C = 0;
    float eps = 1e-6f;
};
cudaError_t launch_dwconv_ln(const DwLnArgs& a, cudaStream_t s);
cudaError_t launch_spectrum(const float* x, int Nh, int Kp, int K, int K2, long rows, float* s_f32, bf16* s_hi, bf16* s_lo,
                            cudaStream_t s);
cudaError_t launch_idft_basis(const float* window, int n_fft, int K, int K2, float* W, cudaStream_t s);
cudaError_t launch_overlap_add(const float* frames, const float* window, int B, int T, int n_fft, int hop, float* audio,
                               cudaStream_t s);

}  // namespace st
