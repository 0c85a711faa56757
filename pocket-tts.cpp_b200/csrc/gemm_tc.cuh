// gemm_tc.cuh — the tensor-core implementation of the generic GEMM family (see gemm.cuh) for sm_100a:
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) into a multi-stage shared-memory ring,
//   * tcgen05.mma (kind::f16, bf16 or f16 inputs, fp32 accumulate) issued by ONE thread, accumulator in TMEM,
//   * epilogue warps read TMEM with tcgen05.ld (32 lanes x 32 columns) and run the fused epilogue.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM lane groups
// (warp & 3) * 32). One 128 x BN output tile per CTA; up to two CTAs per SM so one tile's epilogue overlaps the
// other's loads/MMAs.
//
// A operand: channel-last activations [slot][rows][C]. A causal conv window (K = taps * C) is NOT materialised:
// k-block kb = (tap, c0) is loaded at row offset +tap, so a conv is the same kernel as a linear (taps == 1).
// 128 GEMM rows are fetched as 128/CH chunks of CH consecutive rows so that tiles may span slots (T % CH == 0).
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include <cuda.h>
#include <map>
#include <tuple>

namespace ptts {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                   "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format: version 1, layout type 2): rows are 128 B,
// 8-row groups are SBO = 1024 B apart; the k-slice inside the 128-byte atom is selected by advancing the start address.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;   // stride byte offset
    d |= 1ull << 46;                                // descriptor version (Blackwell)
    d |= 2ull << 61;                                // SWIZZLE_128B
    return d;
}

struct TcParams {
    int R, N, K;
    int kb_per_tap;       // C / 64
    int CH;               // rows per TMA chunk
    int cps;              // chunks per slot (T / CH)
    int total_chunks;     // ceil(R / CH) (plain) or n_slots * cps
    uint32_t idesc;
};

// Vectorised GENERIC epilogue: 8 consecutive columns of one row.
__device__ __forceinline__ void epi_vec8(const Epi& e, int row, int col, float (&v)[8], long long ro, long long r2, long long rr) {
    if (e.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + col)), b1 = __ldg(reinterpret_cast<const float4*>(e.bias + col + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (e.colscale) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.colscale + col)), b1 = __ldg(reinterpret_cast<const float4*>(e.colscale + col + 4));
        v[0] *= b0.x; v[1] *= b0.y; v[2] *= b0.z; v[3] *= b0.w; v[4] *= b1.x; v[5] *= b1.y; v[6] *= b1.z; v[7] *= b1.w;
    }
    if (e.rowmul) {
        const float* g = e.rowmul + (long long)row * e.rowmul_ld + col;
        const float4 b0 = *reinterpret_cast<const float4*>(g), b1 = *reinterpret_cast<const float4*>(g + 4);
        v[0] *= b0.x; v[1] *= b0.y; v[2] *= b0.z; v[3] *= b0.w; v[4] *= b1.x; v[5] *= b1.y; v[6] *= b1.z; v[7] *= b1.w;
    }
    if (e.resid) {
        const float* g = e.resid + rr + col;
        const float4 b0 = *reinterpret_cast<const float4*>(g), b1 = *reinterpret_cast<const float4*>(g + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (e.out) {
        float* g = e.out + ro + col;
        *reinterpret_cast<float4*>(g) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(g + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (e.out2_type != OUT2_NONE) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = apply_act(v[i], e.act);
        if (e.out2_type == OUT2_BF16) {
            __nv_bfloat162 p[4];
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
            *reinterpret_cast<uint4*>((__nv_bfloat16*)e.out2 + r2 + col) = *reinterpret_cast<uint4*>(p);
        } else {
            __half2 hi[4];
#pragma unroll
            for (int i = 0; i < 4; i++) hi[i] = __floats2half2_rn(a[2 * i], a[2 * i + 1]);
            *reinterpret_cast<uint4*>((__half*)e.out2 + r2 + col) = *reinterpret_cast<uint4*>(hi);
            if (e.out2_type == OUT2_F16_SPLIT) {
                __half2 lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 h = __half22float2(hi[i]);
                    lo[i] = __floats2half2_rn(a[2 * i] - h.x, a[2 * i + 1] - h.y);
                }
                *reinterpret_cast<uint4*>((__half*)e.out2 + r2 + col + e.split_off) = *reinterpret_cast<uint4*>(lo);
            }
        }
    }
}

template <int BN>
struct TcCfg {
    static constexpr int STAGES = (BN >= 128) ? 3 : 4;
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int W_BYTES = BN * 128;
    static constexpr int SMEM = STAGES * (A_BYTES + W_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                         const TcParams p, const Epi epi) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sW = base + STAGES * Cfg::A_BYTES;
    const uint32_t bars = sW + STAGES * Cfg::W_BYTES;          // full[STAGES] | empty[STAGES] | tmem_full | tmem_ptr
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull = bars + 16 * STAGES, tptr = tfull + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = blockIdx.x, tile_m = blockIdx.y;
    const int num_kb = p.K / 64;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const int cpt = 128 / p.CH;                         // chunks per tile
            const int g0 = tile_m * cpt;
            int valid = p.total_chunks - g0; if (valid > cpt) valid = cpt;
            const uint32_t bytes = (uint32_t)(valid * p.CH * 128 + Cfg::W_BYTES);
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const int tap = kb / p.kb_per_tap, c0 = (kb % p.kb_per_tap) * 64;
                mbar_expect_tx(full0 + 8 * s, bytes);
                for (int c = 0; c < valid; c++) {
                    const int g = g0 + c;
                    const int slot = g / p.cps, t0 = (g % p.cps) * p.CH;
                    tma_load_3d(sA + s * Cfg::A_BYTES + c * p.CH * 128, &tmA, full0 + 8 * s, c0, t0 + tap, slot);
                }
                tma_load_2d(sW + s * Cfg::W_BYTES, &tmW, full0 + 8 * s, kb * 64, tile_n * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint64_t ad = make_smem_desc_sw128(sA + s * Cfg::A_BYTES), bd = make_smem_desc_sw128(sW + s * Cfg::W_BYTES);
#pragma unroll
                for (int k = 0; k < 4; k++) tc_mma_f16(tmem_base, ad + 2 * k, bd + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
                tc_commit(empty0 + 8 * s);                       // frees the smem stage when these MMAs retire
            }
            tc_commit(tfull);                                    // accumulator complete
        }
    } else {
        // ===== epilogue warps =====
        const int ew = warp & 3;
        const int row = tile_m * 128 + ew * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const bool live = row < p.R;
        long long ro = 0, r2 = 0, rr = 0;
        if (live && epi.mode == EPI_GENERIC) {
            ro = epi.out ? epi.out_map.off(row, epi.rps) : 0;
            r2 = epi.out2 ? epi.out2_map.off(row, epi.rps) : 0;
            rr = epi.resid ? epi.resid_map.off(row, epi.rps) : 0;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tc_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
            if (live) {
                const int col0 = tile_n * BN + c0;
                if (epi.mode == EPI_GENERIC) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float w[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) w[i] = v[j + i];
                        epi_vec8(epi, row, col0 + j, w, ro, r2, rr);
                    }
                } else {
                    epi_apply<32>(epi, row, col0, v, p.N);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor-map cache + dispatch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcPlanCache {
    PFN_tmapEncodeTiled encode = nullptr;
    std::map<std::tuple<const void*, long long, long long, long long, long long, int, int>, CUtensorMap> maps;
    bool attr_set[3] = {false, false, false};
};

inline TcPlanCache* tc_plan_cache_create() {
    auto* c = new TcPlanCache;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) c->encode = (PFN_tmapEncodeTiled)fn;
    return c;
}
inline void tc_plan_cache_destroy(TcPlanCache* c) { delete c; }

struct TcGeom { int C, taps, T, CH, cps, n_slots, rows_per_slot_buf; long long slot_stride; bool ok; };

inline TcGeom tc_geometry(int R, int K, const RowMap& amap, int a_rps) {
    TcGeom g{}; g.ok = false;
    const long long C = amap.row_stride;
    if (C <= 0 || C % 64 != 0 || K % C != 0 || amap.base != 0) return g;
    g.C = (int)C; g.taps = (int)(K / C);
    if (amap.slot_stride == 0) {                       // plain matrix (one "slot")
        if (g.taps != 1) return g;
        g.T = R; g.CH = 128; g.cps = (R + 127) / 128; g.n_slots = 1; g.rows_per_slot_buf = R; g.slot_stride = (long long)R * C;
        g.ok = true; return g;
    }
    if (R % a_rps != 0 || amap.slot_stride % C != 0) return g;
    g.T = a_rps; g.n_slots = R / a_rps; g.slot_stride = amap.slot_stride; g.rows_per_slot_buf = (int)(amap.slot_stride / C);
    if (g.rows_per_slot_buf < g.T + g.taps - 1) return g;
    for (int ch : {128, 64, 32, 16}) if (g.T % ch == 0) { g.CH = ch; break; }
    if (!g.CH) return g;
    g.cps = g.T / g.CH;
    g.ok = true; return g;
}

inline int tc_pick_bn(int R, int N) {
    const int tiles_m = (R + 127) / 128;
    for (int bn : {128, 64, 32}) {
        if (N % bn != 0) continue;
        if (tiles_m * (N / bn) >= 120 || bn == 32) return bn;
    }
    return 0;
}

template <typename T>
inline bool tc_gemm_supported(int R, int N, int K, const RowMap& amap, int a_rps) {
    if (R < 64 || K % 64 != 0 || N % 32 != 0) return false;
    if (!tc_geometry(R, K, amap, a_rps).ok) return false;
    return tc_pick_bn(R, N) != 0;
}

inline const CUtensorMap* tc_get_map(TcPlanCache* c, const void* ptr, bool f16, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                                     const cuuint32_t* box) {
    auto key = std::make_tuple(ptr, (long long)dims[0], (long long)dims[1], (long long)(rank > 2 ? dims[2] : 1),
                               (long long)(rank > 2 ? strides_bytes[1] : 0), (int)(box[1] | (box[0] << 16)), (int)f16);
    auto it = c->maps.find(key);
    if (it != c->maps.end()) return &it->second;
    CUtensorMap m;
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = c->encode(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims,
                           strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "ptts_b200: cuTensorMapEncodeTiled failed (%d)\n", (int)r); abort(); }
    return &(c->maps[key] = m);
}

template <typename T>
inline int tc_gemm_launch(TcPlanCache* c, const T* A, RowMap amap, int a_rps, const T* W, int R, int N, int K, const Epi& epi, cudaStream_t stream) {
    if (!c || !c->encode) { fprintf(stderr, "ptts_b200: tensor-map encoder unavailable\n"); abort(); }
    constexpr bool f16 = std::is_same<T, __half>::value;
    const TcGeom g = tc_geometry(R, K, amap, a_rps);
    const int bn = tc_pick_bn(R, N);
    cuuint64_t adims[3] = {(cuuint64_t)g.C, (cuuint64_t)g.rows_per_slot_buf, (cuuint64_t)g.n_slots};
    cuuint64_t astr[2] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.slot_stride * 2};
    cuuint32_t abox[3] = {64, (cuuint32_t)g.CH, 1};
    const CUtensorMap* ta = tc_get_map(c, A, f16, 3, adims, astr, abox);
    cuuint64_t wdims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t wstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t wbox[2] = {64, (cuuint32_t)bn};
    const CUtensorMap* tw = tc_get_map(c, W, f16, 2, wdims, wstr, wbox);
    TcParams p; p.R = R; p.N = N; p.K = K; p.kb_per_tap = g.C / 64; p.CH = g.CH; p.cps = g.cps;
    p.total_chunks = (amap.slot_stride == 0) ? (R + 127) / 128 : g.n_slots * g.cps;
    // instruction descriptor (kind::f16): D=f32, A/B = bf16|f16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
    p.idesc = (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    dim3 grid(N / bn, (R + 127) / 128);
    const int bi = bn == 128 ? 0 : (bn == 64 ? 1 : 2);
    auto launch = [&](auto kern, int smem) {
        if (!c->attr_set[bi]) { PTTS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); c->attr_set[bi] = true; }
        kern<<<grid, 192, smem, stream>>>(*ta, *tw, p, epi);
    };
    if (bn == 128) launch(gemm_tc_kernel<128>, TcCfg<128>::SMEM);
    else if (bn == 64) launch(gemm_tc_kernel<64>, TcCfg<64>::SMEM);
    else launch(gemm_tc_kernel<32>, TcCfg<32>::SMEM);
    return 1;
}

}  // namespace ptts
