// tcgen05 / TMA / TMEM conv-GEMM engine (sm_100a) — the product path for every dense contraction
// of the estimator (cond_proj, in_proj, QKV, O, conv_1, conv_2, long-skip convs, final_proj).
//
//   D[128 frames x BN channels] (fp32, TMEM) += A[128 x 64] (bf16, smem, K-major, SW128)
//                                              · B[BN x 64]^T (bf16, smem, K-major, SW128)
//
// * split-bf16 ("bf16x3"): every operand travels as hi = bf16(x), lo = bf16(x - hi); each k-step
//   issues Ahi·Bhi + Ahi·Blo + Alo·Bhi into the same fp32 TMEM accumulator (~16 mantissa bits;
//   plain bf16 cannot meet the 1e-3 parity bar).
// * k-tap Conv1d = taps shifted accumulating GEMMs: the A tile of tap j is the TMA box at frame
//   coordinate t0 + j - pad of a 3-D (C, T, batch) tensor map; frames outside [0, T) are zero-filled
//   by TMA — exactly the reference's zero padding at TENSOR edges (not utterance edges).
// * the U-Net long-skip concat is never materialised: k-blocks walk two A tensor maps.
// * persistent, warp-specialised CTA (one per SM): warp 0 = TMA producer, warp 1 = MMA issuer
//   (one elected thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers ->
//   fused bias/SiLU/FiLM/mask/gate/residual -> fp32 and/or split-bf16 global stores).  Two TMEM
//   accumulator stages let tile i's epilogue overlap tile i+1's MMAs.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <cstring>
#include <cstdlib>

namespace st {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;             // bf16 elements = 128 bytes = one SW128 row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;            // warps 0-3: TMA / MMA / TMEM-alloc / spare; warps 4-11: epilogue
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;     // 16 KB

struct TcMaps {
    CUtensorMap a_hi[2], a_lo[2], w_hi, w_lo;
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "selp.b32 %0, 1, 0, px;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row (1024 B) swizzle atoms.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type=SWIZZLE_128B(2) [61,64))
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

template <int BN> struct Cfg {
    static constexpr int B_TILE_BYTES = BN * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 2 : 3;
    static constexpr int TMEM_COLS = 2 * BN;                 // two accumulator stages (power of 2 >= 32)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * 4096 /*epilogue staging*/;
};

// ----------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full = empty_bar + C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    pdl_trigger();                     // successor may start its prologue now; it waits for us before touching memory
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb0 = (p.Cs0 + BLOCK_K - 1) / BLOCK_K;
    const int kb1 = p.n_src > 1 ? (p.Cs1 + BLOCK_K - 1) / BLOCK_K : 0;
    const int kb_per_tap = kb0 + kb1;
    const int num_kb = p.taps * kb_per_tap;
    const int pad = p.taps / 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a_hi[0]); prefetch_tmap(&maps.a_lo[0]); prefetch_tmap(&maps.w_hi); prefetch_tmap(&maps.w_lo);
        if (p.n_src > 1) { prefetch_tmap(&maps.a_hi[1]); prefetch_tmap(&maps.a_lo[1]); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                        // predecessors complete: operands / residuals are valid from here on

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
                const int bb = m_tile / p.m_tiles_per_b, t0 = (m_tile % p.m_tiles_per_b) * BLOCK_M;
                const int ab = bb % p.a_bmod, n0 = n_tile * BN;
                // channel block OUTER, tap INNER: the k taps of one channel block read the same A rows shifted by one
                // frame, back to back, so taps 1.. hit L2 (tap-outer order re-read the whole A slab from HBM per tap)
                // (p.tap_outer = 1 restores the old order for A/B runs: STABLETTS_B200_TAP_OUTER=1)
                for (int it = 0; it < num_kb; ++it) {
                    {
                        const int kb = p.tap_outer ? it % kb_per_tap : it / p.taps;
                        const int tap = p.tap_outer ? it / kb_per_tap : it % p.taps;
                        const int src = kb >= kb0 ? 1 : 0;
                        const int kc = (src ? kb - kb0 : kb) * BLOCK_K;          // channel offset inside the source
                        const int kw = (src ? p.Cs0 : 0) + kc;                   // column in the packed weight
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* s = smem + stage * C::STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                        tma_load_3d(&maps.a_hi[src], &full_bar[stage], s, kc, t0 + tap - pad, ab);
                        tma_load_3d(&maps.a_lo[src], &full_bar[stage], s + A_TILE_BYTES, kc, t0 + tap - pad, ab);
                        tma_load_2d(&maps.w_hi, &full_bar[stage], s + 2 * A_TILE_BYTES, kw, tap * p.N + n0);
                        tma_load_2d(&maps.w_lo, &full_bar[stage], s + 2 * A_TILE_BYTES + C::B_TILE_BYTES, kw, tap * p.N + n0);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc(BN);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t a_hi = make_sw128_desc(sa), a_lo = make_sw128_desc(sa + A_TILE_BYTES);
                    const uint64_t b_hi = make_sw128_desc(sa + 2 * A_TILE_BYTES);
                    const uint64_t b_lo = make_sw128_desc(sa + 2 * A_TILE_BYTES + C::B_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);   // +32 B per K step inside the 128 B row
                        umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);   // small terms first
                        umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    // commits are issued by the SAME thread that issued the MMAs
                    umma_commit(&empty_bar[stage]);                       // frees this smem stage when its MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ================= epilogue (8 warps: 4 TMEM lane quarters x 2 column halves) =================
        const int wq = warp & 3;                       // TMEM lane quarter this warp may access
        const int eh = (warp - 4) >> 2;                // two warps share a lane quarter: even / odd 32-column chunks
        float4* stg = reinterpret_cast<float4*>(smem + C::STAGES * C::STAGE_BYTES + 256) + (warp - 4) * 256;   // 4 KB per warp
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
            const int bb = m_tile / p.m_tiles_per_b, t0 = (m_tile % p.m_tiles_per_b) * BLOCK_M + wq * 32;
            const uint32_t tacc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN);
            if (p.flags & EPI_ROPE) {                  // kernel-uniform
                RopeRegs rr;
                epilogue_rope_prefetch(p, t0, lane, rr);
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                epilogue_tile<BN, true>(p, bb, t0, n_tile * BN, tacc, stg, eh, lane, &rr);
            } else {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                epilogue_tile<BN, false>(p, bb, t0, n_tile * BN, tacc, stg, eh, lane, nullptr);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
    }
}

// ----------------------------------------------------------------------------------------------
// host side: tensor-map construction (cached) and launch
// ----------------------------------------------------------------------------------------------
std::string g_err = "";
std::mutex g_mu;
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
bool g_attr_set[2] = {false, false};

struct MapKey {
    const void* ptr; uint64_t d0, d1, d2; uint32_t b0, b1; int rank;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && b0 == o.b0 && b1 == o.b1 && rank == o.rank;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h ^= k.d0 * 0x9E3779B97F4A7C15ull + (h << 6); h ^= k.d1 * 0xC2B2AE3D27D4EB4Full + (h >> 3);
        h ^= k.d2 * 0x165667B19E3779F9ull + (h << 9); h ^= ((size_t)k.b0 << 20) ^ ((size_t)k.b1 << 4) ^ (size_t)k.rank;
        return h;
    }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

bool ensure_encode() {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        g_err = "cuTensorMapEncodeTiled driver entry point unavailable";
        return false;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return true;
}

// bf16 tensor, dim0 contiguous.  rank 3: (d0, d1, d2) box (b0, b1, 1); rank 2: (d0, d1) box (b0, b1)
bool get_map(const void* ptr, int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, CUtensorMap* out) {
    MapKey key{ptr, d0, d1, d2, b0, b1, rank};
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return true; }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (rank " + std::to_string(rank) +
                ", dims " + std::to_string(d0) + "x" + std::to_string(d1) + "x" + std::to_string(d2) + ")";
        return false;
    }
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, m);
    *out = m;
    return true;
}

template <int BN>
cudaError_t launch_bn(const GemmArgs& g, int num_sms, cudaStream_t s) {
    using C = Cfg<BN>;
    TcMaps maps;
    for (int i = 0; i < g.n_src; ++i) {
        if (!get_map(g.A_hi[i], 3, (uint64_t)g.Cs[i], (uint64_t)g.T, (uint64_t)g.a_bmod, BLOCK_K, BLOCK_M, &maps.a_hi[i])) return cudaErrorInvalidValue;
        if (!get_map(g.A_lo[i], 3, (uint64_t)g.Cs[i], (uint64_t)g.T, (uint64_t)g.a_bmod, BLOCK_K, BLOCK_M, &maps.a_lo[i])) return cudaErrorInvalidValue;
    }
    if (g.n_src == 1) { maps.a_hi[1] = maps.a_hi[0]; maps.a_lo[1] = maps.a_lo[0]; }
    if (!get_map(g.W_hi, 2, (uint64_t)g.Ktot, (uint64_t)g.taps * g.N, 1, BLOCK_K, BN, &maps.w_hi)) return cudaErrorInvalidValue;
    if (!get_map(g.W_lo, 2, (uint64_t)g.Ktot, (uint64_t)g.taps * g.N, 1, BLOCK_K, BN, &maps.w_lo)) return cudaErrorInvalidValue;
    TcParams p;
    fill_tc_params(p, g);
    p.m_tiles_per_b = (g.T + BLOCK_M - 1) / BLOCK_M;
    p.n_tiles = (g.N + BN - 1) / BN;
    p.total_tiles = g.BB * p.m_tiles_per_b * p.n_tiles;
    constexpr int idx = BN == 256 ? 1 : 0;
    if (!g_attr_set[idx]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) { g_err = "cudaFuncSetAttribute(max dynamic smem) failed"; return e; }
        g_attr_set[idx] = true;
    }
    const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    return launch_k(gemm_tc_kernel<BN>, dim3(grid), dim3(NUM_THREADS), (size_t)C::SMEM_BYTES, s, maps, p);
}

}  // namespace

const char* gemm_tc_last_error() { return g_err.c_str(); }

// shared with attention_tc.cu (cached, mutex-free: callers serialise through their own launch mutex)
bool tmap_encode_bf16(const void* ptr, int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                      CUtensorMap* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!ensure_encode()) return false;
    return get_map(ptr, rank, d0, d1, d2, b0, b1, out);
}

bool gemm_tc2_eligible(const GemmArgs& g, int num_sms);
cudaError_t launch_gemm_tc2(const GemmArgs& g, int num_sms, cudaStream_t s);
const char* gemm_tc2_last_error();

// 0 = 1-CTA kernel only, 1 = 2-CTA kernel where eligible (default), 2 = 2-CTA whenever shapes allow (tests)
static int tc2_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("STABLETTS_B200_TC2");
        mode = !e ? 1 : (!strcmp(e, "force") ? 2 : (!strcmp(e, "0") ? 0 : 1));
    }
    return mode;
}

cudaError_t launch_gemm_tc(const GemmArgs& g, int num_sms, cudaStream_t s) {
    if (g.BB == 0 || g.T == 0) return cudaSuccess;
    if ((g.flags & EPI_ROPE) && (g.flags & (EPI_SILU | EPI_FILM | EPI_MASK | EPI_GATE | EPI_RESID))) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_err = "EPI_ROPE combines with EPI_BIAS only (the QKV epilogue variant)";
        return cudaErrorInvalidValue;
    }
    {
        const int mode = tc2_mode();
        bool ok = g.A_hi[0] && g.W_hi && g.N >= 256 && g.N % 128 == 0 && g.Ktot % 8 == 0 && g.Cs[0] % 8 == 0 &&
                  (g.n_src == 1 || (g.Cs[0] % BLOCK_K == 0 && g.Cs[1] % 8 == 0 && g.A_hi[1]));
        if (mode && ok && (mode == 2 || gemm_tc2_eligible(g, num_sms))) {
            cudaError_t e = launch_gemm_tc2(g, num_sms, s);
            if (e != cudaSuccess) { std::lock_guard<std::mutex> lk(g_mu); g_err = std::string("2-CTA kernel: ") + gemm_tc2_last_error(); }
            return e;
        }
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (!ensure_encode()) return cudaErrorNotSupported;
    for (int i = 0; i < g.n_src; ++i) {
        if (!g.A_hi[i] || !g.A_lo[i]) { g_err = "split-bf16 A planes missing"; return cudaErrorInvalidValue; }
        if (g.Cs[i] % 8) { g_err = "A channels must be a multiple of 8 (16-byte TMA stride)"; return cudaErrorInvalidValue; }
        if (i == 0 && g.n_src > 1 && g.Cs[0] % BLOCK_K) { g_err = "first concat source must be a multiple of 64 channels"; return cudaErrorInvalidValue; }
    }
    if (!g.W_hi || !g.W_lo || g.Ktot % 8 || g.N % 8) { g_err = "bad weight operand"; return cudaErrorInvalidValue; }
    // tile-width choice: wider tiles halve A re-reads; narrower tiles quantise better on 148 SMs
    const long m_tiles = (long)g.BB * ((g.T + BLOCK_M - 1) / BLOCK_M);
    bool use256 = false;
    if (g.N > 128) {
        const long t256 = m_tiles * ((g.N + 255) / 256), t128 = m_tiles * ((g.N + 127) / 128);
        const long w256 = (t256 + num_sms - 1) / num_sms * 2, w128 = (t128 + num_sms - 1) / num_sms;
        use256 = false && (w256 <= w128 + w128 / 8);   // 128x256 with 2 stages is TMA-latency bound (profiles/r1a); 2-CTA 256x256 is the planned wide tile
    }
    return use256 ? launch_bn<256>(g, num_sms, s) : launch_bn<128>(g, num_sms, s);
}

}  // namespace st
