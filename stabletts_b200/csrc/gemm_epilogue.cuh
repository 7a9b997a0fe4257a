// Fused conv-GEMM epilogue shared by the 1-CTA (gemm_tc.cu) and 2-CTA (gemm_tc2.cu) tcgen05 kernels.
//
// One call drains this warp's share of a finished accumulator tile (its TMEM lane quarter = 32 frames,
// every second 32-column chunk):
//   Phase A: tcgen05.ld (thread = frame, 32 columns) -> XOR-swizzled shared-memory staging, no math.
//   Phase B: lane = (4 frames x 8 float4 column groups): every global access is a coalesced 128-byte
//            row segment (residual read, fp32 / split-bf16 writes); per-column vectors (bias, gate,
//            FiLM) sit in registers for the whole chunk, the per-frame mask for the whole tile.
//   v = acc + bias; [SiLU]; [partial RoPE on q/k head chunks, q pre-scaled for the exp2 softmax];
//   v = (gamma*v + beta) * mask * gate + resid  -> fp32 and/or split-bf16 planes.
#pragma once
#include "common.cuh"
#include <cstdlib>
#include "tc_ptx.cuh"

namespace st {

struct TcParams {
    int n_src, Cs0, Cs1, taps, N, a_bmod, BB, T;
    int m_tiles_per_b, n_tiles, total_tiles;
    int flags, B, film_H, c_clamp, resid_clamp, rope_H, tap_outer, dbg;
    long film_bstride, gate_bstride;
    const float *bias, *mask, *film, *gate, *resid, *rope_cs;
    float* out_f32; bf16* out_hi; bf16* out_lo;
};

inline void fill_tc_params(TcParams& p, const GemmArgs& g) {
    p.n_src = g.n_src; p.Cs0 = g.Cs[0]; p.Cs1 = g.Cs[1]; p.taps = g.taps; p.N = g.N; p.a_bmod = g.a_bmod; p.BB = g.BB; p.T = g.T;
    p.flags = g.flags; p.B = g.B; p.film_H = g.film_H; p.c_clamp = g.c_clamp; p.resid_clamp = g.resid_clamp; p.rope_H = g.rope_H;
    p.film_bstride = g.film_bstride; p.gate_bstride = g.gate_bstride;
    p.bias = g.bias; p.mask = g.mask; p.film = g.film; p.gate = g.gate; p.resid = g.resid; p.rope_cs = g.rope_cs;
    p.out_f32 = g.out_f32; p.out_hi = g.out_hi; p.out_lo = g.out_lo;
    static int tap_outer = -1;
    if (tap_outer < 0) { const char* e = getenv("STABLETTS_B200_TAP_OUTER"); tap_outer = (e && e[0] == '1') ? 1 : 0; }
    p.tap_outer = tap_outer;
    static int dbg = -1;            // TEMPORARY timing experiments (results are wrong when set): 1 = no global stores, 2 = no stores, no math
    if (dbg < 0) { const char* e = getenv("STABLETTS_B200_EPI_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
}

// softmax scale folded into q: 1/sqrt(64) * log2(e) (attention runs in the exp2 domain)
constexpr float kQScale = 0.125f * 1.4426950408889634f;

// Per-tile RoPE table prefetch for the QKV GEMM: the (cos, sin) pairs a lane needs depend only on its frames
// (t0 + 4*it + rs) and its pair group (c4 & 3), not on the column chunk, so they are loaded ONCE per tile --
// before the wait on the accumulator barrier, which hides their L2 latency behind the MMAs.
struct RopeRegs { float4 a[8], b[8]; };
__device__ __forceinline__ void epilogue_rope_prefetch(const TcParams& p, int t0, int lane, RopeRegs& r) {
    const int rs = lane >> 3, c4 = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int tc = min(t0 + it * 4 + rs, p.T - 1);
        const float* q = p.rope_cs + ((long)tc * 16 + (c4 & 3) * 4) * 2;
        r.a[it] = __ldg(reinterpret_cast<const float4*>(q));
        r.b[it] = __ldg(reinterpret_cast<const float4*>(q + 4));
    }
}

// bb: batch row, t0: first frame of this warp's 32-frame slab, n0: first column of the tile,
// tmem_acc: TMEM address of (lane quarter, accumulator column 0), stg: this warp's 4 KB staging.
// ROPE = true is the QKV variant (bias + partial RoPE + q pre-scale only; launch_gemm_tc rejects EPI_ROPE
// combined with SiLU/FiLM/mask/gate/residual); ROPE = false is everything else.  Two instances keep each
// one's registers and instruction footprint small; a launch only ever executes one of them.
template <int BN, bool ROPE>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, int bb, int t0, int n0, uint32_t tmem_acc, float4* stg,
                                              int eh, int lane, const RopeRegs* rr) {
    using namespace ptx;
    const int rs = lane >> 3, c4 = lane & 7;
    const int mb = bb % p.B;
    const long obase = (long)bb * p.T * p.N;
    // The two warps of a lane quarter split the 32-column chunks as {0,3,4,7,..} / {1,2,5,6,..}: with RoPE only
    // the even chunks (first half of every 64-wide head) carry the rotation, and this split gives each warp half of them.
    auto chunk_col = [eh](int kc) { return (2 * kc + ((kc & 1) ^ eh)) * 32; };   // increasing in kc
    uint32_t v[32];
    if (n0 + chunk_col(0) < p.N) tmem_ld32(tmem_acc + (uint32_t)chunk_col(0), v);

    float mrow[8];
    const float *film = nullptr, *gate = nullptr, *resid = nullptr;
    bool plain = true;
    if constexpr (!ROPE) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int t = t0 + it * 4 + rs;
            mrow[it] = ((p.flags & EPI_MASK) && t < p.T) ? __ldg(p.mask + (long)mb * p.T + t) : 1.f;
        }
        film = p.film + (long)mb * p.film_bstride;
        gate = p.gate + (long)min(bb, p.c_clamp) * p.gate_bstride;
        resid = p.resid + (long)min(bb, p.resid_clamp) * p.T * p.N;
        plain = (p.flags & (EPI_FILM | EPI_MASK | EPI_GATE | EPI_RESID)) == 0;
    }

#pragma unroll 1
    for (int kc = 0; kc < BN / 64; ++kc) {
        const int c0 = chunk_col(kc);
        if (n0 + c0 >= p.N) break;             // warp-uniform; later chunks lie further right
        const int nb = n0 + c0;                // chunk base column (multiple of 32), warp-uniform
        const int n = nb + c4 * 4;
        const bool col_ok = n < p.N;           // N % 4 == 0: a float4 column group is all-in or all-out
        // per-column vectors: issued before the TMEM wait so their latency overlaps it
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f), fg = g4, fb = b4, bp4 = b4;
        // RoPE applies to the first 32 dims of every 64-wide head of q and k (columns [0, 2H));
        // pairs (j, j+16) live in lanes c4 and c4^4 of the same frame (models/diffusion_transformer.py:173-198)
        const bool rope = ROPE && nb < 2 * p.rope_H && (nb & 63) == 0;
        const float post = (ROPE && nb < p.rope_H) ? kQScale : 1.0f;
        if (col_ok) {
            if (p.flags & EPI_BIAS) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            if constexpr (ROPE) {
                if (rope && (p.flags & EPI_BIAS)) bp4 = __ldg(reinterpret_cast<const float4*>(p.bias + (n ^ 16)));   // partner column
            } else {
                if (p.flags & EPI_GATE) g4 = __ldg(reinterpret_cast<const float4*>(gate + n));
                if (p.flags & EPI_FILM) {
                    fg = __ldg(reinterpret_cast<const float4*>(film + n));
                    fb = __ldg(reinterpret_cast<const float4*>(film + p.film_H + n));
                }
            }
        }
        float4 rv[8];
        if constexpr (!ROPE) {                 // residual rows of the whole chunk, also ahead of the TMEM wait
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int t = t0 + it * 4 + rs;
                rv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((p.flags & EPI_RESID) && col_ok && t < p.T)
                    rv[it] = __ldg(reinterpret_cast<const float4*>(resid + (long)t * p.N + n));
            }
        }
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
            stg[lane * 8 + (q ^ (lane & 7))] = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                           __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
        if (kc + 1 < BN / 64 && n0 + chunk_col(kc + 1) < p.N)      // next chunk's TMEM read flies during phase B
            tmem_ld32(tmem_acc + (uint32_t)chunk_col(kc + 1), v);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + rs;
            const int t = t0 + rl;
            const float4 sv = stg[rl * 8 + (c4 ^ (rl & 7))];
            if (p.dbg == 2) { if (sv.x == 1.2345e30f) stg[0].x = 1.f; continue; }     // TEMPORARY: phase A only
            float x[4] = {sv.x + b4.x, sv.y + b4.y, sv.z + b4.z, sv.w + b4.w};
            if constexpr (ROPE) {
                if (rope) {                    // warp-uniform branch
                    const float4 pv = stg[rl * 8 + ((c4 ^ 4) ^ (rl & 7))];      // partner dims (j +- 16) of the same frame
                    const float sgn = (c4 < 4) ? -1.f : 1.f;      // r_j = -x_{j+16} (j<16), +x_{j-16} (j>=16)
                    const float4 cs0 = rr->a[it], cs1 = rr->b[it];
                    x[0] = x[0] * cs0.x + sgn * (pv.x + bp4.x) * cs0.y;
                    x[1] = x[1] * cs0.z + sgn * (pv.y + bp4.y) * cs0.w;
                    x[2] = x[2] * cs1.x + sgn * (pv.z + bp4.z) * cs1.y;
                    x[3] = x[3] * cs1.z + sgn * (pv.w + bp4.w) * cs1.w;
                }
                x[0] *= post; x[1] *= post; x[2] *= post; x[3] *= post;
            } else {
                if (p.flags & EPI_SILU) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[e] = silu_f(x[e]);
                } else if (p.flags & EPI_GELU) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[e] = gelu_f(x[e]);
                }
                if (!plain) {                  // bias / SiLU only (cond_proj) skips the neutral FiLM·mask·gate+resid chain
                    const float m = mrow[it];
                    x[0] = (fg.x * x[0] + fb.x) * m * g4.x + rv[it].x;
                    x[1] = (fg.y * x[1] + fb.y) * m * g4.y + rv[it].y;
                    x[2] = (fg.z * x[2] + fb.z) * m * g4.z + rv[it].z;
                    x[3] = (fg.w * x[3] + fb.w) * m * g4.w + rv[it].w;
                }
            }
            if (t < p.T && col_ok && (p.dbg == 0 || x[0] == 1.2345e30f)) {
                const long o = obase + (long)t * p.N + n;
                if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = make_float4(x[0], x[1], x[2], x[3]);
                if (p.out_hi) {
                    uint32_t h01, l01, h23, l23;
                    split_bf16x2(x[0], x[1], h01, l01); split_bf16x2(x[2], x[3], h23, l23);
                    *reinterpret_cast<uint2*>(p.out_hi + o) = make_uint2(h01, h23);
                    *reinterpret_cast<uint2*>(p.out_lo + o) = make_uint2(l01, l23);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace st
