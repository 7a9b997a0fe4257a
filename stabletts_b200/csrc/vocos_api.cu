// C-ABI + orchestration of the vocoder hand-off: the reference's Vocos
// (vocoders/vocos/models/model.py:11-20) on the mel this library's CFM path emits (api.py:76).
//
//   mel (B, n_mel, T) -> token-major split planes
//   embed Conv1d k=7 (backbone.py:30,50)                       conv-GEMM, 7 taps, K = n_mel, N = dim
//   LayerNorm(dim, eps 1e-6) (:31,51)                          row kernel
//   12 x ConvNeXtBlock (module.py:34-46):
//       depthwise k=7 conv + LayerNorm                         row kernel (one pass, split-bf16 out)
//       pwconv1 + exact GELU                                   GEMM K = dim, N = intermediate, EPI_GELU
//       pwconv2, * gamma, + residual                           GEMM K = intermediate, N = dim, EPI_GATE | EPI_RESID
//   final LayerNorm (backbone.py:43,55)                        row kernel
//   head Linear dim -> n_fft + 2 (head.py:96-101)              GEMM, log-magnitudes and phases in two 128-aligned column
//                                                              groups ([0, K) and [Kp, Kp + K), K = n_fft/2 + 1)
//   exp / clip / cos / sin (head.py:103-113)                   elementwise, emits the split [re | im] operand
//   irfft * window (head.py:62-63)                             ONE GEMM against the windowed inverse-DFT basis
//                                                              (oracle/vocoder_ref.py idft_basis: exact in float64)
//   fold / envelope / trim (head.py:66-81)                     4-frame gather
#include "handle.cuh"
#include "vocos.cuh"
#include <cmath>

using namespace st;

namespace st {

struct VocosState {
    st_vocos_dims d;
    int K = 0, Kp = 0, Nh = 0, K2 = 0;         // bins, phase column offset, padded head width, padded spectrum width
    GemmW embed, head, basis;
    std::vector<GemmW> pw1, pw2;
    std::vector<float*> dw_w, dw_b, ln_w, ln_b, gamma;
    float *norm_w = nullptr, *norm_b = nullptr, *fln_w = nullptr, *fln_b = nullptr, *window = nullptr;
    void* ws = nullptr; size_t ws_bytes = 0;
};

void vocos_free(st_handle* h) {
    VocosState* v = (VocosState*)h->vocos;
    if (!v) return;
    if (v->ws) cudaFree(v->ws);
    delete v;
    h->vocos = nullptr;
}

int vocos_finalize(st_handle* h, cudaStream_t s) {
    VocosState* v = (VocosState*)h->vocos;
    if (!v) return fail(h, "internal: vocoder state missing");
    const st_vocos_dims& d = v->d;
    const int L = d.n_layers, C = d.dim, I = d.intermediate;
    v->pw1.assign(L, GemmW()); v->pw2.assign(L, GemmW());
    v->dw_w.assign(L, nullptr); v->dw_b.assign(L, nullptr); v->ln_w.assign(L, nullptr); v->ln_b.assign(L, nullptr);
    v->gamma.assign(L, nullptr);
    if (pack_gemm(h, &v->embed, {"backbone.embed"}, C, d.n_mel, 7, 0, d.n_mel, true, s)) return 1;
    if (get_raw(h, "backbone.norm.weight", C, &v->norm_w) || get_raw(h, "backbone.norm.bias", C, &v->norm_b)) return 1;
    for (int l = 0; l < L; ++l) {
        const std::string p = "backbone.convnext." + std::to_string(l) + ".";
        float* dw;
        if (get_raw(h, p + "dwconv.weight", (int64_t)C * 7, &dw)) return 1;
        if (dev_alloc(h, &v->dw_w[l], (size_t)7 * C)) return 1;          // (C, 1, 7) -> [7][C]: float4 loads over channels
        ST_CUDA(launch_pack_conv(dw, v->dw_w[l], C, 1, 7, C, 0, 0, 1, s));
        if (get_raw(h, p + "dwconv.bias", C, &v->dw_b[l])) return 1;
        if (get_raw(h, p + "norm.weight", C, &v->ln_w[l]) || get_raw(h, p + "norm.bias", C, &v->ln_b[l])) return 1;
        if (get_raw(h, p + "gamma", C, &v->gamma[l])) return 1;
        if (pack_gemm(h, &v->pw1[l], {p + "pwconv1"}, I, C, 1, 0, C, true, s)) return 1;
        if (pack_gemm(h, &v->pw2[l], {p + "pwconv2"}, C, I, 1, 0, I, true, s)) return 1;
    }
    if (get_raw(h, "backbone.final_layer_norm.weight", C, &v->fln_w) || get_raw(h, "backbone.final_layer_norm.bias", C, &v->fln_b)) return 1;
    if (get_raw(h, "head.istft.window", d.n_fft, &v->window)) return 1;
    {   // head.out (n_fft + 2, dim): rows [0, K) = log-magnitudes, [K, 2K) = phases (chunk(2, dim=1), head.py:102) -> two
        // 128-aligned column groups of a zero-filled (Nh, dim) matrix
        float *w, *b;
        if (get_raw(h, "head.out.weight", (int64_t)2 * v->K * C, &w) || get_raw(h, "head.out.bias", 2 * v->K, &b)) return 1;
        GemmW& g = v->head;
        g.taps = 1; g.N = v->Nh; g.K = C;
        const size_t n = (size_t)v->Nh * C;
        if (dev_alloc(h, &g.f32, n) || dev_alloc(h, &g.hi, n) || dev_alloc(h, &g.lo, n) || dev_alloc(h, &g.bias, (size_t)v->Nh)) return 1;
        ST_CUDA(cudaMemsetAsync(g.f32, 0, n * 4, s));
        ST_CUDA(cudaMemsetAsync(g.bias, 0, (size_t)v->Nh * 4, s));
        ST_CUDA(launch_pack_conv(w, g.f32, v->K, C, 1, v->Nh, 0, 0, C, s));
        ST_CUDA(launch_pack_conv(w + (size_t)v->K * C, g.f32, v->K, C, 1, v->Nh, v->Kp, 0, C, s));
        ST_CUDA(cudaMemcpyAsync(g.bias, b, (size_t)v->K * 4, cudaMemcpyDeviceToDevice, s));
        ST_CUDA(cudaMemcpyAsync(g.bias + v->Kp, b + v->K, (size_t)v->K * 4, cudaMemcpyDeviceToDevice, s));
        ST_CUDA(launch_split(g.f32, g.hi, g.lo, (long)n, s));
    }
    {   // windowed inverse-DFT basis (n_fft outputs x K2)
        GemmW& g = v->basis;
        g.taps = 1; g.N = d.n_fft; g.K = v->K2;
        const size_t n = (size_t)d.n_fft * v->K2;
        if (dev_alloc(h, &g.f32, n) || dev_alloc(h, &g.hi, n) || dev_alloc(h, &g.lo, n)) return 1;
        ST_CUDA(launch_idft_basis(v->window, d.n_fft, v->K, v->K2, g.f32, s));
        ST_CUDA(launch_split(g.f32, g.hi, g.lo, (long)n, s));
    }
    return 0;
}

}  // namespace st

namespace {

struct VocosWs { Act mel, E, X, U, Hid, Hd, S, F; size_t bytes = 0; };

void layout_vocos_ws(const st_handle* h, const VocosState* v, VocosWs& w, void* base, int B, int T) {
    const st_vocos_dims& d = v->d;
    const bool tc = h->engine == ST_ENGINE_TCGEN05;
    const size_t rows = (size_t)B * T;
    Bump bp(base, 0);
    auto mk = [&](Act& a, int C, bool f32, bool split) {
        a.C = C;
        a.f32 = f32 ? bp.take<float>(rows * C) : nullptr;
        a.hi = split ? bp.take<bf16>(rows * C) : nullptr;
        a.lo = split ? bp.take<bf16>(rows * C) : nullptr;
    };
    mk(w.mel, d.n_mel, !tc, tc);
    mk(w.E, d.dim, true, false);
    mk(w.X, d.dim, true, false);
    mk(w.U, d.dim, !tc, tc);
    mk(w.Hid, d.intermediate, !tc, tc);
    mk(w.Hd, v->Nh, true, false);
    mk(w.S, v->K2, !tc, tc);
    mk(w.F, d.n_fft, true, false);
    w.bytes = bp.off + 256;
}

}  // namespace

extern "C" {

int st_create_vocos(const st_vocos_dims* dims, int device, st_handle** out) {
    if (!dims || !out) return fail(nullptr, "st_create_vocos: null argument");
    const st_vocos_dims& d = *dims;
    if (d.dim != 512 && d.dim != 768 && d.dim != 1024) return fail(nullptr, "Vocos dim must be 512, 768 or 1024 (reference VocosConfig: 768)");
    if (d.n_mel <= 0 || d.n_mel % 16) return fail(nullptr, "Vocos input_channels must be a positive multiple of 16");
    if (d.intermediate <= 0 || d.intermediate % 64) return fail(nullptr, "Vocos intermediate_dim must be a multiple of 64");
    if (d.n_layers <= 0 || d.n_layers > 64) return fail(nullptr, "Vocos num_layers out of range");
    if (d.hop <= 0 || d.n_fft <= 0 || d.n_fft % 128 || d.n_fft % d.hop || d.n_fft / d.hop > 16 || (d.n_fft - d.hop) % 2)
        return fail(nullptr, "Vocos n_fft must be a multiple of 128 and of hop_length, with at most 16 overlapping frames");
    // a CFM-estimator-shaped handle carries the device / engine / error plumbing; its dims are the reference ModelConfig's
    st_dims base = {80, 256, 1024, 4, 6, 3, 256};
    int rc = st_create(&base, device, out);
    if (rc) return rc;
    st_handle* h = *out;
    h->kind = 2;
    VocosState* v = new VocosState();
    v->d = d;
    v->K = d.n_fft / 2 + 1;
    v->Kp = (v->K + 127) / 128 * 128;          // phases start at a 128-aligned column
    v->Nh = 2 * v->Kp;
    v->K2 = 2 * ((v->K + 63) / 64 * 64);       // [re | im], each half padded to the GEMM's 64-channel K block
    h->vocos = v;
    return 0;
}

int st_vocos_forward(st_handle* h, const float* mel, float* audio, int B, int T, void* stream) {
    if (!h) return 1;
    ST_ENTER(h);
    if (h->kind != 2 || !h->vocos) return fail(h, "handle is not a Vocos vocoder");
    if (!h->finalized) return fail(h, "weights not finalized (call st_finalize_weights)");
    if (!mel || !audio) return fail(h, "st_vocos_forward: null pointer");
    if (B <= 0 || T <= 0 || B > 32767) return fail(h, "B and T must be positive");
    VocosState* v = (VocosState*)h->vocos;
    const st_vocos_dims& d = v->d;
    cudaStream_t s = (cudaStream_t)stream;
    VocosWs w;
    layout_vocos_ws(h, v, w, nullptr, B, T);
    if (w.bytes > v->ws_bytes) {
        if (v->ws) { ST_CUDA(cudaStreamSynchronize(s)); cudaFree(v->ws); v->ws = nullptr; v->ws_bytes = 0; }
        ST_CUDA(cudaMalloc(&v->ws, w.bytes));
        v->ws_bytes = w.bytes;
    }
    layout_vocos_ws(h, v, w, v->ws, B, T);
    const long rows = (long)B * T;
    auto base = [&](int flags) {
        GemmArgs g;
        g.BB = B; g.T = T; g.a_bmod = B; g.B = B; g.resid_clamp = B - 1; g.c_clamp = 0; g.flags = flags;
        return g;
    };
    ST_LAUNCH(launch_bct_to_btc(mel, w.mel.f32, w.mel.hi, w.mel.lo, B, d.n_mel, T, nullptr, s));
    {   // embed: Conv1d(n_mel -> dim, k = 7, padding 3) (backbone.py:30,50)
        GemmArgs g = base(EPI_BIAS);
        if (run_gemm(h, g, v->embed, &w.mel, nullptr, w.E, s)) return 1;
    }
    DwLnArgs ln;
    ln.B = B; ln.T = T; ln.C = d.dim; ln.eps = 1e-6f;
    ln.x = w.E.f32; ln.ln_w = v->norm_w; ln.ln_b = v->norm_b; ln.out_f32 = w.X.f32;
    ST_LAUNCH_P(ST_PROF_LN, 0, (double)rows * d.dim * 8, s, launch_dwconv_ln(ln, s));              // backbone.py:51
    for (int l = 0; l < d.n_layers; ++l) {                                                         // module.py:34-46
        DwLnArgs a;
        a.B = B; a.T = T; a.C = d.dim; a.eps = 1e-6f;
        a.x = w.X.f32; a.dw_w = v->dw_w[l]; a.dw_b = v->dw_b[l]; a.ln_w = v->ln_w[l]; a.ln_b = v->ln_b[l];
        a.out_f32 = w.U.f32; a.out_hi = w.U.hi; a.out_lo = w.U.lo;
        ST_LAUNCH_P(ST_PROF_LN, 0, (double)rows * d.dim * 8, s, launch_dwconv_ln(a, s));
        {
            GemmArgs g = base(EPI_BIAS | EPI_GELU);
            if (run_gemm(h, g, v->pw1[l], &w.U, nullptr, w.Hid, s, ST_PROF_GEMM_C1)) return 1;
        }
        {   // x = residual + gamma * pwconv2(h)
            GemmArgs g = base(EPI_BIAS | EPI_GATE | EPI_RESID);
            g.gate = v->gamma[l]; g.gate_bstride = 0; g.resid = w.X.f32;
            if (run_gemm(h, g, v->pw2[l], &w.Hid, nullptr, w.X, s, ST_PROF_GEMM_C2)) return 1;
        }
    }
    ln.x = w.X.f32; ln.ln_w = v->fln_w; ln.ln_b = v->fln_b; ln.out_f32 = w.U.f32; ln.out_hi = w.U.hi; ln.out_lo = w.U.lo;
    ST_LAUNCH_P(ST_PROF_LN, 0, (double)rows * d.dim * 8, s, launch_dwconv_ln(ln, s));              // backbone.py:55
    {   // head.out (head.py:101)
        GemmArgs g = base(EPI_BIAS);
        if (run_gemm(h, g, v->head, &w.U, nullptr, w.Hd, s)) return 1;
    }
    ST_LAUNCH(launch_spectrum(w.Hd.f32, v->Nh, v->Kp, v->K, v->K2, rows, w.S.f32, w.S.hi, w.S.lo, s));
    {   // frames = window * irfft(S) as one contraction
        GemmArgs g = base(0);
        if (run_gemm(h, g, v->basis, &w.S, nullptr, w.F, s)) return 1;
    }
    ST_LAUNCH(launch_overlap_add(w.F.f32, v->window, B, T, d.n_fft, d.hop, audio, s));
    return 0;
}

}  // extern "C"
