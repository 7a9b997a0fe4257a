// 2-CTA (cta_group::2) variant of the tcgen05 conv-GEMM engine: a CTA pair (one TPC, cluster of 2)
// computes a 256-frame x 256-channel tile with ONE tcgen05.mma.cta_group::2 stream issued by the
// leader CTA.  Each CTA stages only its own 128 A rows and HALF of the B tile (128 of the 256 output
// channels), so shared-memory fill traffic per MMA cycle is 42.7 B/clk/SM instead of 85 B/clk/SM for
// the 1-CTA 128x128 tile — the 1-CTA kernel is L2->SM bandwidth bound at ~50 % tensor-pipe
// utilisation with split-bf16 operands (profiles/r1a_gemm_tc_full.csv), this one is not.
//
// Same GemmArgs contract, same fused epilogue as gemm_tc.cu.  Barrier wiring:
//   full[s]        leader only   count 1: leader's producer arrive.expect_tx(2 x stage bytes); BOTH CTAs'
//                                TMA loads complete_tx on the leader's barrier (cta_group::2 TMA form)
//   empty[s]       each CTA      count 1: tcgen05.commit.cta_group::2 multicast from the leader's MMA thread
//   tmem_full[a]   each CTA      count 1: multicast commit after a tile's last k-block
//   tmem_empty[a]  leader only   count 8: the 4 epilogue warps of group a in each CTA (the peer arrives remotely, mapa)
#include "common.cuh"
#include "tc_ptx.cuh"
#include "gemm_epilogue.cuh"
#include <string>
#include <cstdlib>
#include <cstring>

namespace st {

bool tmap_encode_bf16(const void* ptr, int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                      CUtensorMap* out);
bool build_epi_maps(const GemmArgs& g, EpiMaps* em);
const char* gemm_tc_last_error();

namespace {

using namespace ptx;

constexpr int BM = 128;                 // rows per CTA (pair tile: 256)
constexpr int BK = 64;
constexpr int UK = 16;
constexpr int THREADS = 384;            // warps 0-3: TMA / MMA / TMEM-alloc / spare; warps 4-11: epilogue (two groups of four)
constexpr int TILE_BYTES = 128 * BK * 2;            // 16 KB: one A plane tile (128 frames x 64 channels)

// BN2 = pair tile width: 256 (each CTA stages 128 weight rows) or 128 (64 weight rows; twice as many, half as long
// tiles — chosen when that shortens the partial last wave, e.g. N = 256 outputs at cfg1: 250 tiles = 3.4 waves of 74
// CTA pairs -> 500 half tiles = 6.8 half waves)
// SHALLOW = 1: two main-loop stages instead of 3 / 4 — for GEMMs whose whole K fits a few k-blocks (O: 4, in_proj: 2) —
// which leaves 64 KB for per-warp residual tiles brought in by the TMA unit (2 x 4 KB per epilogue warp).  A thread reading
// its own residual row (8 x 16 B of one line) costs ~3000 cycles per chunk when nothing hides it (profiles/r2i_trace.log);
// the long k-tap convs hide it behind their main loop and keep the deep pipeline.
// PREC = 1: the opt-in two-pass FFN precision — ONE fp16 A plane, fp16 hi / lo weight planes, two MMAs per k-step
// (A16·Wlo + A16·Whi): a stage is A16 | Bh_hi | Bh_lo = 48 / 32 KB, so four stages fit where three did.
template <int BN2, int SHALLOW = 0, int PREC = 0> struct Cfg2 {
    static constexpr int B_TILE_BYTES = (BN2 / 2) * BK * 2;                  // B-half plane tile
    static constexpr int A_PLANES = PREC ? 1 : 2;
    static constexpr int STAGE_BYTES = A_PLANES * TILE_BYTES + 2 * B_TILE_BYTES;    // [A_hi, A_lo | A16], Bh_hi, Bh_lo
    static constexpr int STAGES = SHALLOW ? 2 : (PREC ? 4 : (BN2 == 256 ? 3 : 4));
    static constexpr int TMEM_COLS = 2 * BN2;                                // two accumulator stages
    static constexpr int STAGING_OFF = STAGES * STAGE_BYTES;                 // 1024-aligned: swizzled TMA-store tiles
    static constexpr int RESID_OFF = STAGING_OFF + EPI_WARPS * EPI_STAGE_BYTES;          // SHALLOW: 8 warps x 2 x 4 KB
    static constexpr int BAR_OFF = RESID_OFF + (SHALLOW ? EPI_WARPS * 2 * 4096 : 0);
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
};

struct Maps2 { CUtensorMap a_hi[2], a_lo[2], w_hi, w_lo; };

typedef TcParams Params2;

// One kernel instance per (tile width, epilogue mode): a combined kernel that switched over the modes at run time made
// ptxas keep every mode's register arrays in one allocation (2 KB of spills); separate instances also keep the
// instruction footprint of a launch small.
template <int BN2, int MODE, int SHALLOW, int PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ Maps2 maps, const __grid_constant__ EpiMaps em, const Params2 p) {
    using CF = Cfg2<BN2, SHALLOW, PREC>;
    constexpr int A_BYTES = CF::A_PLANES * TILE_BYTES;          // offset of the B tiles inside a stage
    constexpr int B_TILE_BYTES = CF::B_TILE_BYTES, STAGE_BYTES = CF::STAGE_BYTES, STAGES = CF::STAGES;
    constexpr int TMEM_COLS = CF::TMEM_COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CF::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* resid_bar = tmem_empty + 2;              // SHALLOW: [8 warps][2 buffers]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resid_bar + 2 * EPI_WARPS);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int kb0 = (p.Cs0 + BK - 1) / BK;
    const int kb1 = p.n_src > 1 ? (p.Cs1 + BK - 1) / BK : 0;
    const int kb_per_tap = kb0 + kb1;
    const int num_kb = p.taps * kb_per_tap;
    const int pad = p.taps / 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a_hi[0]); prefetch_tmap(&maps.a_lo[0]); prefetch_tmap(&maps.w_hi); prefetch_tmap(&maps.w_lo);
        if (p.n_src > 1) { prefetch_tmap(&maps.a_hi[1]); prefetch_tmap(&maps.a_lo[1]); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EPI_WARPS); }   // 4 warps of group i x 2 CTAs
        if (SHALLOW) for (int i = 0; i < 2 * EPI_WARPS; ++i) mbar_init(&resid_bar[i], 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync();                    // peer barriers initialised, both TMEM allocations done
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
                const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
                const int bb = m_tile / p.m_tiles_per_b, t0 = (m_tile % p.m_tiles_per_b) * (2 * BM) + (int)rank * BM;
                const int ab = bb % p.a_bmod, n0 = n_tile * BN2 + (int)rank * (BN2 / 2);
                // channel block OUTER, tap INNER: the k taps of one channel block read the same A rows shifted by one
                // frame, back to back, so taps 1.. hit L2 (tap-outer order re-read the whole A slab from HBM per tap)
                // (p.tap_outer = 1 restores the old order for A/B runs: STABLETTS_B200_TAP_OUTER=1)
                for (int it = 0; it < num_kb; ++it) {
                    {
                        const int kb = p.tap_outer ? it % kb_per_tap : it / p.taps;
                        const int tap = p.tap_outer ? it / kb_per_tap : it % p.taps;
                        const int src = kb >= kb0 ? 1 : 0;
                        const int kc = (src ? kb - kb0 : kb) * BK;
                        const int kw = (src ? p.Cs0 : 0) + kc;
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* s = smem + stage * STAGE_BYTES;
                        if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        tma_load_3d_2sm(&maps.a_hi[src], &full_bar[stage], s, kc, t0 + tap - pad, ab);          // PREC: the fp16 plane
                        if (!PREC) tma_load_3d_2sm(&maps.a_lo[src], &full_bar[stage], s + TILE_BYTES, kc, t0 + tap - pad, ab);
                        tma_load_2d_2sm(&maps.w_hi, &full_bar[stage], s + A_BYTES, kw, tap * p.N + n0);
                        tma_load_2d_2sm(&maps.w_lo, &full_bar[stage], s + A_BYTES + B_TILE_BYTES, kw, tap * p.N + n0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader) {
            constexpr uint32_t idesc = PREC ? make_idesc_f16(2 * BM, BN2) : make_idesc_bf16(2 * BM, BN2);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN2;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t a_hi = make_sw128_desc(sa), a_lo = make_sw128_desc(sa + TILE_BYTES);
                        const uint64_t b_hi = make_sw128_desc(sa + A_BYTES), b_lo = make_sw128_desc(sa + A_BYTES + B_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            const uint64_t adv = (uint64_t)((k * UK * 2) >> 4);
                            if (PREC) {                    // fp16 operands: A16·Wlo + A16·Whi (small term first)
                                umma_bf16_2sm(tmem_d, a_hi + adv, b_lo + adv, idesc, (kb | k) != 0);
                                umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, 1);
                            } else {
                                umma_bf16_2sm(tmem_d, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
                                umma_bf16_2sm(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
                                umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, 1);
                            }
                        }
                        umma_commit_2sm(&empty_bar[stage], 0b11);                       // both CTAs' stage s may be refilled
                        if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc], 0b11);   // both epilogues may drain
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue (both CTAs, own 128 rows; thread = frame) =================
        // warps 4-7 drain accumulator stage 0 (tiles 0, 2, 4, ... of this cluster), warps 8-11 stage 1 (tiles 1, 3, ...)
        const int wq = warp & 3, grp = (warp - 4) >> 2;
        const uint32_t stg = smem_u32(smem + CF::STAGING_OFF + (warp - 4) * EPI_STAGE_BYTES);
        if (lane == 0) {
            prefetch_tmap(&em.o_f32); prefetch_tmap(&em.o_hi); prefetch_tmap(&em.o_lo);
            if (MODE == EM_LN) { prefetch_tmap(&em.u_hi); prefetch_tmap(&em.u_lo); prefetch_tmap(&em.o2_f32); }
            if (SHALLOW) prefetch_tmap(&em.resid);
        }
        ResidPipe rp;
        if (SHALLOW) { rp.buf = smem_u32(smem + CF::RESID_OFF + (warp - 4) * 2 * 4096); rp.bar = resid_bar + 2 * (warp - 4); }
        // single-wave launch (at most one tile per cluster): both groups drain THE tile, chunk-interleaved (latency, not throughput)
        const bool split = MODE != EM_LN && !SHALLOW && p.total_tiles <= num_clusters;
        const int acc = split ? 0 : grp;
        int tile_it = split ? 0 : grp; uint32_t acc_phase = 0;
        for (int tile = cluster_id + (split ? 0 : grp * num_clusters); tile < p.total_tiles; tile += 2 * num_clusters, tile_it += 2) {
            const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
            const int bb = m_tile / p.m_tiles_per_b;
            const int t0 = (m_tile % p.m_tiles_per_b) * (2 * BM) + (int)rank * BM + wq * 32;
            const uint32_t tacc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN2);
            uint64_t* fb = &tmem_full[acc];
            const uint32_t ph = acc_phase;
            epilogue_tile<BN2, MODE, SHALLOW != 0>(p, em, bb, t0, n_tile * BN2, tacc, stg, lane, tile_it, rp, split ? grp : -1, [fb, ph]() { mbar_wait(fb, ph); tc_fence_after(); });
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);      // the LEADER's barrier gates the next MMA
            acc_phase ^= 1;
        }
        if (lane == 0) bulk_wait_read0();              // the TMA unit has read this warp's staging before the CTA exits (the
                                                       // global writes themselves complete with the grid)
    }

    tc_fence_before();
    cluster_sync();                    // nobody exits (or frees TMEM) while the peer can still signal it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
    }
}

std::string g_err2;
long long* g_epi_trace = nullptr;      // debug: [4 tiles][8 pass-1 + 8 pass-2 chunks][8] clock64 stamps of CTA 0 / warp 4 (STABLETTS_B200_EPI_TRACE=1)

}  // namespace

int gemm_tc2_read_trace(long long* host_out) {
    if (!g_epi_trace) return 1;
    return cudaMemcpy(host_out, g_epi_trace, 4 * 16 * 8 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
}

const char* gemm_tc2_last_error() { return g_err2.c_str(); }

// eligibility: wide outputs and enough pair tiles to fill the 74 TPCs
bool gemm_tc2_eligible(const GemmArgs& g, int num_sms) {
    if (g.N < 256 || g.N % 128) return false;
    const long pair_tiles = (long)g.BB * ((g.T + 255) / 256) * ((g.N + 255) / 256);
    return pair_tiles >= (num_sms / 2);
}

// pair-tile width: 128 when halving the tiles shortens the partial last wave by more than the ~3 % the narrower
// MMAs and the doubled A re-reads (L2 hits) cost; STABLETTS_B200_TC2_BN=128|256 forces one (A/B runs)
static int pick_bn2(const GemmArgs& g, int pairs) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("STABLETTS_B200_TC2_BN"); forced = e ? atoi(e) : 0; }
    if (g.ln) return 256;                          // fused LayerNorm: a CTA must own whole 256-channel rows
    if (forced == 128 || forced == 256) return forced;
    const long m_tiles = (long)g.BB * ((g.T + 2 * BM - 1) / (2 * BM));
    const long t256 = m_tiles * ((g.N + 255) / 256), t128 = m_tiles * ((g.N + 127) / 128);
    const double w256 = (double)((t256 + pairs - 1) / pairs), w128 = 0.515 * (double)((t128 + pairs - 1) / pairs);
    return w128 < 0.97 * w256 ? 128 : 256;
}

template <int BN2, int MODE, int SHALLOW, int PREC = 0>
static cudaError_t launch_tc2_inst(const Maps2& maps, const EpiMaps& em, const Params2& p, int pairs, cudaStream_t s) {
    static std::atomic<uint64_t> attr_done{0};      // one bit per device
    using CF = Cfg2<BN2, SHALLOW, PREC>;
    cudaError_t e = ensure_dyn_smem(gemm_tc2_kernel<BN2, MODE, SHALLOW, PREC>, CF::SMEM_BYTES, attr_done);
    if (e != cudaSuccess) { g_err2 = "cudaFuncSetAttribute(max dynamic smem) failed for gemm_tc2_kernel"; return e; }
    const int clusters = p.total_tiles < pairs ? p.total_tiles : pairs;
    return launch_k(gemm_tc2_kernel<BN2, MODE, SHALLOW, PREC>, dim3(2 * clusters), dim3(THREADS), (size_t)CF::SMEM_BYTES, s, maps, em, p);
}

// residual GEMMs whose K is a few k-blocks (O, in_proj) run the shallow-pipeline instance with TMA-loaded residual tiles
static bool use_shallow(const GemmArgs& g) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("STABLETTS_B200_SHALLOW"); env = (e && !strcmp(e, "0")) ? 0 : 1; }
    return env && (g.flags & EPI_RESID) && g.resid && (long)g.Ktot * g.taps <= 512;
}

template <int BN2>
static cudaError_t launch_tc2_bn(const Maps2& maps, const EpiMaps& em, Params2& p, const GemmArgs& g, int pairs, cudaStream_t s) {
    p.n_tiles = (g.N + BN2 - 1) / BN2;
    p.total_tiles = g.BB * p.m_tiles_per_b * p.n_tiles;
    const bool sh = use_shallow(g);
    if (g.prec) {                      // two-pass fp16 FFN convs: conv_1 (SiLU), conv_2 (residual, with or without the fused LayerNorm)
        switch (p.mode) {
            case EM_SILU:  return launch_tc2_inst<BN2, EM_SILU, 0, 1>(maps, em, p, pairs, s);
            case EM_LN:    return launch_tc2_inst<BN2, EM_LN, 0, 1>(maps, em, p, pairs, s);
            case EM_RESID: return launch_tc2_inst<BN2, EM_RESID, 0, 1>(maps, em, p, pairs, s);
            case EM_PLAIN: return launch_tc2_inst<BN2, EM_PLAIN, 0, 1>(maps, em, p, pairs, s);      // long-skip conv without the fused LayerNorm
            default: g_err2 = "the two-pass fp16 precision is built for the FFN and long-skip convs only"; return cudaErrorInvalidValue;
        }
    }
    switch (p.mode) {
        case EM_ROPE: return launch_tc2_inst<BN2, EM_ROPE, 0>(maps, em, p, pairs, s);
        case EM_LN:   return sh ? launch_tc2_inst<BN2, EM_LN, 1>(maps, em, p, pairs, s) : launch_tc2_inst<BN2, EM_LN, 0>(maps, em, p, pairs, s);
        case EM_SILU: return launch_tc2_inst<BN2, EM_SILU, 0>(maps, em, p, pairs, s);
        case EM_GELU: return launch_tc2_inst<BN2, EM_GELU, 0>(maps, em, p, pairs, s);
        case EM_RESID: return sh ? launch_tc2_inst<BN2, EM_RESID, 1>(maps, em, p, pairs, s) : launch_tc2_inst<BN2, EM_RESID, 0>(maps, em, p, pairs, s);
        default:      return launch_tc2_inst<BN2, EM_PLAIN, 0>(maps, em, p, pairs, s);
    }
}

cudaError_t launch_gemm_tc2(const GemmArgs& g, int num_sms, cudaStream_t s) {
    Maps2 maps;
    const int pairs = num_sms / 2;
    const int bn2 = pick_bn2(g, pairs);
    for (int i = 0; i < g.n_src; ++i) {            // (a 2-byte-element map serves bf16 and fp16 planes alike)
        if (!tmap_encode_bf16(g.A_hi[i], 3, (uint64_t)g.Cs[i], (uint64_t)g.T, (uint64_t)g.a_bmod, BK, BM, &maps.a_hi[i])) {
            g_err2 = gemm_tc_last_error(); return cudaErrorInvalidValue;
        }
        if (g.prec) maps.a_lo[i] = maps.a_hi[i];
        else if (!tmap_encode_bf16(g.A_lo[i], 3, (uint64_t)g.Cs[i], (uint64_t)g.T, (uint64_t)g.a_bmod, BK, BM, &maps.a_lo[i])) {
            g_err2 = gemm_tc_last_error(); return cudaErrorInvalidValue;
        }
    }
    if (g.n_src == 1) { maps.a_hi[1] = maps.a_hi[0]; maps.a_lo[1] = maps.a_lo[0]; }
    if (!tmap_encode_bf16(g.W_hi, 2, (uint64_t)g.Ktot, (uint64_t)g.taps * g.N, 1, BK, bn2 / 2, &maps.w_hi) ||
        !tmap_encode_bf16(g.W_lo, 2, (uint64_t)g.Ktot, (uint64_t)g.taps * g.N, 1, BK, bn2 / 2, &maps.w_lo)) {
        g_err2 = gemm_tc_last_error(); return cudaErrorInvalidValue;
    }
    if (g.ln && g.N != 256) { g_err2 = "fused LayerNorm needs N == 256 (one tile spans the whole row)"; return cudaErrorInvalidValue; }
    EpiMaps em;
    if (!build_epi_maps(g, &em)) { g_err2 = "epilogue store maps: missing output plane / " + std::string(gemm_tc_last_error()); return cudaErrorInvalidValue; }
    Params2 p;
    fill_tc_params(p, g);
    p.m_tiles_per_b = (g.T + 2 * BM - 1) / (2 * BM);
    if (getenv("STABLETTS_B200_EPI_TRACE")) {
        if (!g_epi_trace) cudaMalloc(&g_epi_trace, 4 * 16 * 8 * sizeof(long long));
        cudaMemsetAsync(g_epi_trace, 0, 4 * 16 * 8 * sizeof(long long), s);
        p.dbg = g_epi_trace;
    }
    return bn2 == 128 ? launch_tc2_bn<128>(maps, em, p, g, pairs, s) : launch_tc2_bn<256>(maps, em, p, g, pairs, s);
}

}  // namespace st
