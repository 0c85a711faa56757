// Non-GEMM kernels of the per-frame path: row preparation (RoPE tables), LayerNorm(+AdaLN modulate),
// FlowLM decode/prefill attention, Mimi ring attention, flow-head glue, Mimi front end, state upkeep.
#pragma once
#include "common.cuh"

namespace ptts {

// ------------------------------------------------------------------------------------------------
// Row preparation
// ------------------------------------------------------------------------------------------------
// (cos, sin)(pos * freq_i) per row. freq tables are computed on the host exactly like the reference's
// two RoPE flavours (rope.h:36-38 for FlowLM, ggml_timestep_embedding for Mimi rope.h:8-20).
__global__ void rope_table_kernel(const int* __restrict__ row_pos, const float* __restrict__ freq, float2* __restrict__ cs, int R) {
    pdl_prologue();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * 32) return;
    const float rad = (float)row_pos[idx >> 5] * freq[idx & 31];
    cs[idx] = make_float2(cosf(rad), sinf(rad));
}

// Decode step, Mimi side: 16 rows per slot (positions mimi_off .. mimi_off+15) and their RoPE table.
__global__ void prepare_mimi_kernel(int slot0, int n, const int* __restrict__ mimi_off, const float* __restrict__ freq,
                                    int* __restrict__ mrow_slot, int* __restrict__ mrow_pos, float2* __restrict__ mcs) {
    pdl_prologue();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * M_T * 32) return;
    const int r = idx >> 5, i = idx & 31, s = slot0 + r / M_T, pos = mimi_off[s] + r % M_T;
    if (i == 0) { mrow_slot[r] = s; mrow_pos[r] = pos; }
    const float rad = (float)pos * freq[i];
    mcs[idx] = make_float2(cosf(rad), sinf(rad));
}

// Text prefill: x[r] = float(embed[token[r]])   (ggml_get_rows, reference conditioners/text.h:29-37)
__global__ void embed_gather_kernel(const __nv_bfloat16* __restrict__ table, const int* __restrict__ tokens, float* __restrict__ x, int R) {
    pdl_prologue();
    const int r = blockIdx.x;
    if (r >= R) return;
    const __nv_bfloat16* src = table + (long long)tokens[r] * D_MODEL;
    for (int i = threadIdx.x; i < D_MODEL; i += blockDim.x) x[(long long)r * D_MODEL + i] = __bfloat162float(src[i]);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (ggml_norm: mean, biased variance of deviations, 1/sqrtf(var+eps); reference src/torch.h:49-60,
// modules/mlp.h:52-64) with optional affine and optional AdaLN modulate y*(1+scale)+shift (mlp.h:3-9).
// Writes a low-precision copy (the A operand of the following GEMM) and/or f32.
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(C / 4) layernorm_kernel(const float* __restrict__ x, RowMap xmap, int rps, int R, float eps,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          const float* __restrict__ shift, const float* __restrict__ scale, int mod_ld,
                                                          __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32) {
    // One CTA (C/4 threads) per row, one float4 per thread: at decode batch sizes (R = 256) a warp-per-row layout leaves most SMs idle.
    pdl_prologue();
    constexpr int NW = C / 128;
    __shared__ float red[2][NW];
    const int row = blockIdx.x, col = threadIdx.x * 4, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (row >= R) return;
    const float4 a = *reinterpret_cast<const float4*>(x + xmap.off(row, rps) + col);
    float v[4] = {a.x, a.y, a.z, a.w};
    const float s1 = warp_sum(v[0] + v[1] + v[2] + v[3]);
    if (lane == 0) red[0][warp] = s1;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NW; i++) tot += red[0][i];
    const float mean = tot / C;
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) { v[i] -= mean; s2 += v[i] * v[i]; }
    s2 = warp_sum(s2);
    if (lane == 0) red[1][warp] = s2;
    __syncthreads();
    tot = 0.f;
#pragma unroll
    for (int i = 0; i < NW; i++) tot += red[1][i];
    const float rs = 1.0f / sqrtf(tot / C + eps);
    float y[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { y[i] = v[i] * rs; if (w) y[i] *= w[col + i]; if (b) y[i] += b[col + i]; }
    if (scale) {
        const float4 sc = *reinterpret_cast<const float4*>(scale + (long long)row * mod_ld + col), sh = *reinterpret_cast<const float4*>(shift + (long long)row * mod_ld + col);
        y[0] = y[0] * (sc.x + 1.f) + sh.x; y[1] = y[1] * (sc.y + 1.f) + sh.y; y[2] = y[2] * (sc.z + 1.f) + sh.z; y[3] = y[3] * (sc.w + 1.f) + sh.w;
    }
    if (out_bf16) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0], y[1]), p1 = __floats2bfloat162_rn(y[2], y[3]);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(out_bf16 + (long long)row * C + col) = pk;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (long long)row * C + col) = make_float4(y[0], y[1], y[2], y[3]);
}

// ------------------------------------------------------------------------------------------------
// FlowLM attention, one query row against its slot's cache rows [0, pos] (reference
// modules/transformer.h:157-199, src/torch.h:128-150: scale 1/8, causal by construction, softmax with f32
// probabilities, f32 PV). One CTA per (row, head). Baseline implementation; see attn_flow_split_kernel.
// ------------------------------------------------------------------------------------------------
template <typename KV>
__global__ void __launch_bounds__(128) attn_flow_kernel(const float* __restrict__ q, const KV* __restrict__ kc, const KV* __restrict__ vc,
                                                        long long kv_slot_stride, const int* __restrict__ row_slot,
                                                        const int* __restrict__ row_pos, __nv_bfloat16* __restrict__ out) {
    pdl_prologue();
    extern __shared__ float sc[];                 // [len] scores
    __shared__ float qs[D_HEAD];
    __shared__ float red[4];
    __shared__ float pv[4][D_HEAD];
    const int row = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int len = row_pos[row] + 1;
    const KV* K = kc + (long long)row_slot[row] * kv_slot_stride + h * D_HEAD;
    const KV* V = vc + (long long)row_slot[row] * kv_slot_stride + h * D_HEAD;
    if (tid < D_HEAD) qs[tid] = q[(long long)row * D_MODEL + h * D_HEAD + tid];
    __syncthreads();
    float mx = -INFINITY;
    for (int j = tid; j < len; j += 128) {
        const KV* kr = K + (long long)j * D_MODEL;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < D_HEAD; d += 8) {
            if constexpr (sizeof(KV) == 2) {
                const uint4 kv = *reinterpret_cast<const uint4*>(kr + d);
                const __nv_bfloat16* ke = reinterpret_cast<const __nv_bfloat16*>(&kv);
#pragma unroll
                for (int t = 0; t < 8; t++) acc = fmaf(__bfloat162float(ke[t]), qs[d + t], acc);
            } else {
                const float4 k0 = *reinterpret_cast<const float4*>(kr + d), k1 = *reinterpret_cast<const float4*>(kr + d + 4);
                acc = fmaf(k0.x, qs[d], acc); acc = fmaf(k0.y, qs[d + 1], acc); acc = fmaf(k0.z, qs[d + 2], acc); acc = fmaf(k0.w, qs[d + 3], acc);
                acc = fmaf(k1.x, qs[d + 4], acc); acc = fmaf(k1.y, qs[d + 5], acc); acc = fmaf(k1.z, qs[d + 6], acc); acc = fmaf(k1.w, qs[d + 7], acc);
            }
        }
        acc *= 0.125f;
        sc[j] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    if (lane == 0) red[wid] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < len; j += 128) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
    sum = warp_sum(sum);
    if (lane == 0) red[wid] = sum;
    __syncthreads();
    const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);
    float a0 = 0.f, a1 = 0.f;
    for (int j = wid; j < len; j += 4) {
        const float p = sc[j] * inv;
        const KV* vr = V + (long long)j * D_MODEL + 2 * lane;
        a0 = fmaf(p, to_f32<KV>(vr[0]), a0);
        a1 = fmaf(p, to_f32<KV>(vr[1]), a1);
    }
    pv[wid][2 * lane] = a0; pv[wid][2 * lane + 1] = a1;
    __syncthreads();
    if (tid < D_HEAD) {
        const float o = pv[0][tid] + pv[1][tid] + pv[2][tid] + pv[3][tid];
        out[(long long)row * D_MODEL + h * D_HEAD + tid] = __float2bfloat16_rn(o);
    }
}

// ------------------------------------------------------------------------------------------------
// FlowLM attention, split-KV streaming version for the bf16 cache (the dominant kernel at large batch: pure KV stream).
// One CTA per (KV split, query row), all 16 heads at once so that every cache row is read as one contiguous 2 KB line.
// Warp 8 lane 0 is the producer: it streams groups of 8 consecutive K rows and V rows (16 KB each, contiguous in the
// cache) into a 3-stage shared-memory ring (two CTAs per SM: 192 KB in flight per SM) with cp.async.bulk + mbarrier complete_tx. Consumer warp w owns key w of each
// stage: lane l, chunk i reads the 16 bytes at i*512 + l*16 of the row = 8 dims of head 4i + l/8, so a row is four
// conflict-free LDS.128 per lane; the dot products are finished with three shuffles inside each 8-lane group; softmax is
// the online (running max / sum) form in fp32. Partial (m, l, acc) per split go to a workspace and attn_flow_merge_kernel
// combines them (for one split the normalised bf16 output is written directly).
// Same math as attn_flow_kernel (reference modules/transformer.h:157-199, src/torch.h:128-150).
// ------------------------------------------------------------------------------------------------
constexpr int AF_STAGES = 3, AF_KEYS = 8, AF_ROW_BYTES = D_MODEL * 2, AF_STAGE_BYTES = 2 * AF_KEYS * AF_ROW_BYTES;   // 32 KB
constexpr int AF_SMEM = AF_STAGES * AF_STAGE_BYTES + 128;
constexpr int AF_MAX_SPLITS = 16;

__device__ __forceinline__ uint32_t af_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void af_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");   // hardware sleep until the phase flips
    } while (!ok);
}

__global__ void __launch_bounds__(288, 2) attn_flow_split_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                                                                 const __nv_bfloat16* __restrict__ vc, long long kv_slot_stride,
                                                                 const int* __restrict__ row_slot, const int* __restrict__ row_pos, int splits,
                                                                 float* __restrict__ ws_ml, float* __restrict__ ws_acc,
                                                                 __nv_bfloat16* __restrict__ out, int* __restrict__ merge_cnt) {
    pdl_prologue();
    extern __shared__ __align__(128) uint8_t af_smem[];
    const int row = blockIdx.y, split = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int len = row_pos[row] + 1;
    int chunk = (len + splits - 1) / splits; chunk = (chunk + AF_KEYS - 1) / AF_KEYS * AF_KEYS;
    const int j_begin = split * chunk, j_end = min(len, j_begin + chunk);
    const int n_keys = max(0, j_end - j_begin);
    const int n_stages_total = (n_keys + AF_KEYS - 1) / AF_KEYS;
    const uint32_t sbase = af_smem_u32(af_smem);
    const uint32_t bars = sbase + AF_STAGES * AF_STAGE_BYTES;      // full[4] | empty[4]
    if (threadIdx.x == 0) {
        for (int s = 0; s < AF_STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * (AF_STAGES + s)), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const long long slot_off = (long long)row_slot[row] * kv_slot_stride;

    if (warp == 8) {
        if (lane == 0) {
            // the KV stream is read exactly once per step: mark it evict-first so that it does not push the Mimi stream's operands
            // (re-read by many tiles) and the next GEMM's prefetched weights out of L2
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            const char* kg = reinterpret_cast<const char*>(kc + slot_off + (long long)j_begin * D_MODEL);
            const char* vg = reinterpret_cast<const char*>(vc + slot_off + (long long)j_begin * D_MODEL);
            for (int it = 0; it < n_stages_total; it++) {
                const int s = it % AF_STAGES; const uint32_t ph = (it / AF_STAGES) & 1;
                af_mbar_wait(bars + 8 * (AF_STAGES + s), ph ^ 1);
                const int nk = min(AF_KEYS, n_keys - it * AF_KEYS);
                const uint32_t bytes = (uint32_t)nk * AF_ROW_BYTES;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * s), "r"(2 * bytes) : "memory");
                const uint32_t dst = sbase + s * AF_STAGE_BYTES;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(dst), "l"(kg + (long long)it * AF_KEYS * AF_ROW_BYTES), "r"(bytes), "r"(bars + 8 * s), "l"(pol) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(dst + AF_KEYS * AF_ROW_BYTES), "l"(vg + (long long)it * AF_KEYS * AF_ROW_BYTES), "r"(bytes), "r"(bars + 8 * s), "l"(pol) : "memory");
            }
        }
    } else {
        // ---- consumers: lane owns dims d(i,e) = i*256 + lane*8 + e of head 4i + lane/8 ----
        float qr[4][8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float* qp = q + (long long)row * D_MODEL + i * 256 + lane * 8;
            const float4 a = *reinterpret_cast<const float4*>(qp), b = *reinterpret_cast<const float4*>(qp + 4);
            qr[i][0] = a.x * 0.125f; qr[i][1] = a.y * 0.125f; qr[i][2] = a.z * 0.125f; qr[i][3] = a.w * 0.125f;
            qr[i][4] = b.x * 0.125f; qr[i][5] = b.y * 0.125f; qr[i][6] = b.z * 0.125f; qr[i][7] = b.w * 0.125f;
        }
        float m[4], l[4], acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; i++) { m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
            for (int e = 0; e < 8; e++) acc[i][e] = 0.f; }
        for (int it = 0; it < n_stages_total; it++) {
            const int s = it % AF_STAGES; const uint32_t ph = (it / AF_STAGES) & 1;
            af_mbar_wait(bars + 8 * s, ph);
            const int nk = min(AF_KEYS, n_keys - it * AF_KEYS);
            if (warp < nk) {
                const uint8_t* kr = af_smem + s * AF_STAGE_BYTES + warp * AF_ROW_BYTES + lane * 16;
                const uint8_t* vr = kr + AF_KEYS * AF_ROW_BYTES;
                float sc[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint4 kv = *reinterpret_cast<const uint4*>(kr + i * 512);
                    const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
                    float a = 0.f;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        a = fmaf(__uint_as_float(w[t] << 16), qr[i][2 * t], a);
                        a = fmaf(__uint_as_float(w[t] & 0xffff0000u), qr[i][2 * t + 1], a);
                    }
                    a += __shfl_xor_sync(0xffffffffu, a, 4);
                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    sc[i] = a;
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float mn = fmaxf(m[i], sc[i]);
                    const float corr = expf(m[i] - mn), p = expf(sc[i] - mn);
                    m[i] = mn; l[i] = l[i] * corr + p;
                    const uint4 vv = *reinterpret_cast<const uint4*>(vr + i * 512);
                    const uint32_t w[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        acc[i][2 * t] = fmaf(acc[i][2 * t], corr, p * __uint_as_float(w[t] << 16));
                        acc[i][2 * t + 1] = fmaf(acc[i][2 * t + 1], corr, p * __uint_as_float(w[t] & 0xffff0000u));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * (AF_STAGES + s)) : "memory");
        }
        // ---- cross-warp merge through shared memory (ring is drained: every stage was consumed by all 8 warps) ----
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float* sm_m = reinterpret_cast<float*>(af_smem);               // [8][16]
        float* sm_l = sm_m + 8 * 16;                                   // [8][16]
        float* sm_a = sm_l + 8 * 16;                                   // [8][1024]
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if ((lane & 7) == 0) { sm_m[warp * 16 + 4 * i + (lane >> 3)] = m[i]; sm_l[warp * 16 + 4 * i + (lane >> 3)] = l[i]; }
            float* ap = sm_a + warp * D_MODEL + i * 256 + lane * 8;
            *reinterpret_cast<float4*>(ap) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            *reinterpret_cast<float4*>(ap + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = threadIdx.x;                                     // 0..255: dims 4t..4t+3, head t/16
        const int h = t >> 4;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; w++) M = fmaxf(M, sm_m[w * 16 + h]);
        float L = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const float mw = sm_m[w * 16 + h];
            const float sc = (mw == -INFINITY) ? 0.f : expf(mw - M);
            L = fmaf(sm_l[w * 16 + h], sc, L);
            const float4 a = *reinterpret_cast<const float4*>(sm_a + w * D_MODEL + 4 * t);
            o[0] = fmaf(a.x, sc, o[0]); o[1] = fmaf(a.y, sc, o[1]); o[2] = fmaf(a.z, sc, o[2]); o[3] = fmaf(a.w, sc, o[3]);
        }
        if (splits == 1) {
            const float inv = 1.0f / L;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0] * inv, o[1] * inv), p1 = __floats2bfloat162_rn(o[2] * inv, o[3] * inv);
            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(out + (long long)row * D_MODEL + 4 * t) = pk;
        } else {
            const long long wo = (long long)row * splits + split;
            if ((t & 15) == 0) { ws_ml[wo * 32 + h] = M; ws_ml[wo * 32 + 16 + h] = L; }
            *reinterpret_cast<float4*>(ws_acc + wo * D_MODEL + 4 * t) = make_float4(o[0], o[1], o[2], o[3]);
            if (merge_cnt) {
                // The LAST split CTA of a row to finish merges all of the row's partials (fixed split order: deterministic) and writes the
                // bf16 output, which removes the separate merge launch from every layer. Release/acquire through the per-row counter.
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                int* flag = reinterpret_cast<int*>(af_smem);
                if (t == 0) { const int prev = atomicAdd(merge_cnt + row, 1); *flag = (prev == splits - 1); if (prev == splits - 1) merge_cnt[row] = 0; }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (*flag) {
                    __threadfence();
                    float Mm = -INFINITY;
                    for (int sp = 0; sp < splits; sp++) Mm = fmaxf(Mm, __ldcg(ws_ml + ((long long)row * splits + sp) * 32 + h));
                    float Lm = 0.f, om[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int sp = 0; sp < splits; sp++) {
                        const long long w2 = (long long)row * splits + sp;
                        const float ms = __ldcg(ws_ml + w2 * 32 + h);
                        const float sc = (ms == -INFINITY) ? 0.f : expf(ms - Mm);
                        Lm = fmaf(__ldcg(ws_ml + w2 * 32 + 16 + h), sc, Lm);
                        const float4 a = __ldcg(reinterpret_cast<const float4*>(ws_acc + w2 * D_MODEL + 4 * t));
                        om[0] = fmaf(a.x, sc, om[0]); om[1] = fmaf(a.y, sc, om[1]); om[2] = fmaf(a.z, sc, om[2]); om[3] = fmaf(a.w, sc, om[3]);
                    }
                    const float inv = 1.0f / Lm;
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(om[0] * inv, om[1] * inv), p1 = __floats2bfloat162_rn(om[2] * inv, om[3] * inv);
                    uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                    *reinterpret_cast<uint2*>(out + (long long)row * D_MODEL + 4 * t) = pk;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Mimi ring attention: 16 queries of one slot x all 250 ring slots, additive 0/-inf bias taken from the
// reference's pattern (src/torch.h:168-221, called with the chunk's START offset, mimi_transformer.h:1198).
// mask_mode 0 = reference (non-causal quirk once offset > 250), 1 = ideal causal ring.
// q (bf16), K/V ring (bf16), probabilities rounded to bf16 (ggml bf16 mul_mat), f32 accumulation.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mimi_masked(int offset, int j, int c, int mask_mode) {
    if (mask_mode == 1) {
        const int last = offset + M_T - 1;
        const int p = last - (((last - c) % M_CTX + M_CTX) % M_CTX);   // newest position living in ring slot c
        return p < 0 || p > offset + j;
    }
    const int start = M_CTX * 2 - M_T;                                   // 484
    const int idx = (offset <= M_CTX ? start - offset : M_CTX - (offset % M_CTX)) + c;
    return idx >= start + 1 + j || (idx <= M_CTX - 1 && idx > M_CTX - 1 - (M_T - j - 1));
}

// One CTA (8 warps) per (slot, head). K and V of the head (250 x 64 bf16 each) are staged once in shared memory; warp w
// owns queries w and w+8. Scores: lane = ring slot (K rows read as 16-byte chunks, XOR-swizzled by row so that the 32 rows
// a warp touches spread over all banks), q broadcast from shared memory. Softmax per query row with warp shuffles, the
// probabilities are rounded to bf16 (ggml's bf16 mul_mat does that to the f32 operand), then PV with lane = 2 output dims.
constexpr int AM_SMEM = (2 * M_CTX * D_HEAD) * 2 + M_T * 256 * 4 + M_T * D_HEAD * 4;   // K,V bf16 + P f32 + Q f32 = 84,480 B

__global__ void __launch_bounds__(256) attn_mimi_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                                                        const __nv_bfloat16* __restrict__ vc, long long kv_slot_stride, int slot0,
                                                        const int* __restrict__ mimi_off, int mask_mode, __nv_bfloat16* __restrict__ out) {
    pdl_prologue();
    extern __shared__ __align__(16) uint8_t am_smem[];
    uint4* sK = reinterpret_cast<uint4*>(am_smem);                               // [250][8 chunks], chunk index XOR (row & 7)
    uint4* sV = sK + M_CTX * 8;                                                  // [250][8 chunks]
    float* sP = reinterpret_cast<float*>(sV + M_CTX * 8);                        // [16][256]
    float* sQ = sP + M_T * 256;                                                  // [16][64]
    const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int slot = slot0 + b;
    const int offset = mimi_off[slot];
    const uint4* Kg = reinterpret_cast<const uint4*>(kc + (long long)slot * kv_slot_stride + h * D_HEAD);   // row stride M_DIM*2 B = 64 uint4
    const uint4* Vg = reinterpret_cast<const uint4*>(vc + (long long)slot * kv_slot_stride + h * D_HEAD);
    for (int i = tid; i < M_CTX * 8; i += 256) {
        const int r = i >> 3, ch = i & 7;
        sK[r * 8 + (ch ^ (r & 7))] = Kg[(long long)r * (M_DIM / 8) + ch];
        sV[i] = Vg[(long long)r * (M_DIM / 8) + ch];
    }
    for (int i = tid; i < M_T * D_HEAD; i += 256) {
        const int t = i >> 6, d = i & 63;
        sQ[i] = __bfloat162float(q[((long long)b * M_T + t) * M_DIM + h * D_HEAD + d]);
    }
    __syncthreads();
    const int q0 = wid, q1 = wid + 8;
    // ---- scores (scale 1/8, additive 0/-inf bias) ----
    for (int c = lane; c < 256; c += 32) {
        float s0 = -INFINITY, s1 = -INFINITY;
        if (c < M_CTX) {
            const bool m0 = mimi_masked(offset, q0, c, mask_mode), m1 = mimi_masked(offset, q1, c, mask_mode);
            if (!(m0 && m1)) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int ch = 0; ch < 8; ch++) {
                    const uint4 kv = sK[c * 8 + (ch ^ (c & 7))];
                    const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
                    const float4 qa0 = *reinterpret_cast<const float4*>(sQ + q0 * 64 + ch * 8), qb0 = *reinterpret_cast<const float4*>(sQ + q0 * 64 + ch * 8 + 4);
                    const float4 qa1 = *reinterpret_cast<const float4*>(sQ + q1 * 64 + ch * 8), qb1 = *reinterpret_cast<const float4*>(sQ + q1 * 64 + ch * 8 + 4);
                    const float k0 = __uint_as_float(w[0] << 16), k1 = __uint_as_float(w[0] & 0xffff0000u), k2 = __uint_as_float(w[1] << 16), k3 = __uint_as_float(w[1] & 0xffff0000u);
                    const float k4 = __uint_as_float(w[2] << 16), k5 = __uint_as_float(w[2] & 0xffff0000u), k6 = __uint_as_float(w[3] << 16), k7 = __uint_as_float(w[3] & 0xffff0000u);
                    a0 = fmaf(k0, qa0.x, a0); a0 = fmaf(k1, qa0.y, a0); a0 = fmaf(k2, qa0.z, a0); a0 = fmaf(k3, qa0.w, a0);
                    a0 = fmaf(k4, qb0.x, a0); a0 = fmaf(k5, qb0.y, a0); a0 = fmaf(k6, qb0.z, a0); a0 = fmaf(k7, qb0.w, a0);
                    a1 = fmaf(k0, qa1.x, a1); a1 = fmaf(k1, qa1.y, a1); a1 = fmaf(k2, qa1.z, a1); a1 = fmaf(k3, qa1.w, a1);
                    a1 = fmaf(k4, qb1.x, a1); a1 = fmaf(k5, qb1.y, a1); a1 = fmaf(k6, qb1.z, a1); a1 = fmaf(k7, qb1.w, a1);
                }
                if (!m0) s0 = a0 * 0.125f;
                if (!m1) s1 = a1 * 0.125f;
            }
        }
        sP[q0 * 256 + c] = s0; sP[q1 * 256 + c] = s1;
    }
    __syncwarp();
    // ---- softmax rows q0, q1 (this warp wrote them) ----
#pragma unroll
    for (int rsel = 0; rsel < 2; rsel++) {
        float* pr = sP + (rsel ? q1 : q0) * 256;
        float mx = -INFINITY;
        for (int c = lane; c < 256; c += 32) mx = fmaxf(mx, pr[c]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int c = lane; c < 256; c += 32) { const float e = expf(pr[c] - mx); pr[c] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int c = lane; c < 256; c += 32) pr[c] = __bfloat162float(__float2bfloat16_rn(pr[c] * inv));
    }
    __syncwarp();
    // ---- PV: lane owns output dims 2*lane, 2*lane+1 ----
    float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
    const uint32_t* sV32 = reinterpret_cast<const uint32_t*>(sV);
    for (int c = 0; c < 248; c += 4) {
        const float4 p0 = *reinterpret_cast<const float4*>(sP + q0 * 256 + c), p1 = *reinterpret_cast<const float4*>(sP + q1 * 256 + c);
        const float pa[4] = {p0.x, p0.y, p0.z, p0.w}, pb[4] = {p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t vv = sV32[(c + u) * 32 + lane];
            const float v0 = __uint_as_float(vv << 16), v1 = __uint_as_float(vv & 0xffff0000u);
            o00 = fmaf(pa[u], v0, o00); o01 = fmaf(pa[u], v1, o01); o10 = fmaf(pb[u], v0, o10); o11 = fmaf(pb[u], v1, o11);
        }
    }
    for (int c = 248; c < M_CTX; c++) {
        const uint32_t vv = sV32[c * 32 + lane];
        const float v0 = __uint_as_float(vv << 16), v1 = __uint_as_float(vv & 0xffff0000u);
        const float pa = sP[q0 * 256 + c], pb = sP[q1 * 256 + c];
        o00 = fmaf(pa, v0, o00); o01 = fmaf(pa, v1, o01); o10 = fmaf(pb, v0, o10); o11 = fmaf(pb, v1, o11);
    }
    *reinterpret_cast<__nv_bfloat162*>(out + ((long long)b * M_T + q0) * M_DIM + h * D_HEAD + 2 * lane) = __floats2bfloat162_rn(o00, o01);
    *reinterpret_cast<__nv_bfloat162*>(out + ((long long)b * M_T + q1) * M_DIM + h * D_HEAD + 2 * lane) = __floats2bfloat162_rn(o10, o11);
}

// ------------------------------------------------------------------------------------------------
// Mimi ring attention on the warp-level tensor cores (mma.sync m16n8k16, bf16 x bf16 -> f32): the 16 queries of one (slot, head) are
// exactly one M = 16 tile: S = Q K^T as key tiles of 8, softmax, O = P V as key steps of 16 x 8 output tiles. No ldmatrix: every global
// access is a 16-byte vector load/store because the contraction index (and the output-dim index) may be permuted freely:
//   * QK^T: lane (g, t) loads K[key 8j+g][32p + 8t .. +7]; its 8 values are used as the k-slots of two k16 steps, and the
//     Q fragments are loaded with the same permutation;
//   * PV: output tile n holds dims {8c + n}, so lane (g, t) needs V[key][8g .. 8g+7] (one 16-byte load) for its four keys
//     16s + {2t, 2t+1, 8+2t, 9+2t}; PRMT interleaves key pairs. The lane ends up owning O[q][16t .. 16t+15]: two 16-byte stores.
// The NORMALISED probabilities are rounded to bf16 before P V (ggml's bf16 mul_mat rounds its f32 operand), as in attn_mimi_kernel.
// (tcgen05 cannot be used here: its minimum M is 64 and a CTA pair per 16-row problem would idle 3/4 of the datapath.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct MimiMaskRow { int lo; int a, b; };      // masked iff c >= lo || (a < c && c <= b) || c >= 250   (reference pattern, row j)
__device__ __forceinline__ MimiMaskRow mimi_mask_row(int offset, int j) {
    const int start = M_CTX * 2 - M_T;
    const int base = offset <= M_CTX ? start - offset : M_CTX - (offset % M_CTX);
    MimiMaskRow r; r.lo = start + 1 + j - base; r.a = M_CTX - 1 - (M_T - j - 1) - base; r.b = M_CTX - 1 - base;
    return r;
}
__device__ __forceinline__ bool mimi_masked_fast(const MimiMaskRow& r, int c) { return c >= r.lo || (c > r.a && c <= r.b) || c >= M_CTX; }

// ------------------------------------------------------------------------------------------------
// Mimi ring attention, four warps per (slot, head): warp w owns ring slots 64w .. 64w+63 (eight key tiles). K is read ONCE: the warp's
// 16 x 64 scores stay in registers while the row max / sum are combined across the four warps through shared memory, then P (normalised,
// rounded to bf16) x V for the warp's own keys and a fixed-order (deterministic) sum of the four partial outputs. A one-warp-per-(slot,
// head) version of the same fragments left ~14 warps per SM to hide HBM latency: 74 us per layer at 256 slots for 131 MB of K/V, vs 35 us.
// ------------------------------------------------------------------------------------------------
constexpr int AM4_LD = 68;                                   // padded row of the partial-output buffers (floats)

__global__ void __launch_bounds__(128) attn_mimi_mma4_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                                                             const __nv_bfloat16* __restrict__ vc, long long kv_slot_stride, int slot0,
                                                             const int* __restrict__ mimi_off, int mask_mode, __nv_bfloat16* __restrict__ out) {
    pdl_prologue();
    __shared__ float s_max[4][16], s_sum[4][16];
    __shared__ __align__(16) float s_o[3][16][AM4_LD];          // partial outputs of warps 1..3
    const int b = blockIdx.x / M_HEADS, h = blockIdx.x % M_HEADS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int slot = slot0 + b;
    const int offset = mimi_off[slot];
    const __nv_bfloat16* K = kc + (long long)slot * kv_slot_stride + h * D_HEAD;
    const __nv_bfloat16* V = vc + (long long)slot * kv_slot_stride + h * D_HEAD;
    uint32_t qa[4][4];
#pragma unroll
    for (int p = 0; p < 2; p++) {
        const uint4 lo = *reinterpret_cast<const uint4*>(q + ((long long)b * M_T + g) * M_DIM + h * D_HEAD + 32 * p + 8 * t);
        const uint4 hi = *reinterpret_cast<const uint4*>(q + ((long long)b * M_T + g + 8) * M_DIM + h * D_HEAD + 32 * p + 8 * t);
        qa[2 * p][0] = lo.x; qa[2 * p][1] = hi.x; qa[2 * p][2] = lo.y; qa[2 * p][3] = hi.y;
        qa[2 * p + 1][0] = lo.z; qa[2 * p + 1][1] = hi.z; qa[2 * p + 1][2] = lo.w; qa[2 * p + 1][3] = hi.w;
    }
    const MimiMaskRow mr0 = mimi_mask_row(offset, g), mr1 = mimi_mask_row(offset, g + 8);
    // ---- scores of this warp's 8 key tiles (kept in registers) ----
    float S[8][4];
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint4 kv[4][2];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int krow = min(64 * warp + 8 * (4 * half + u) + g, M_CTX - 1);        // keys 250..255 do not exist: clamp, mask below
            const __nv_bfloat16* kr = K + (long long)krow * M_DIM + 8 * t;
            kv[u][0] = *reinterpret_cast<const uint4*>(kr);
            kv[u][1] = *reinterpret_cast<const uint4*>(kr + 32);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float (&sc)[4] = S[4 * half + u];
            sc[0] = sc[1] = sc[2] = sc[3] = 0.f;
            mma_bf16_16816(sc, qa[0], kv[u][0].x, kv[u][0].y);
            mma_bf16_16816(sc, qa[1], kv[u][0].z, kv[u][0].w);
            mma_bf16_16816(sc, qa[2], kv[u][1].x, kv[u][1].y);
            mma_bf16_16816(sc, qa[3], kv[u][1].z, kv[u][1].w);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int c = 64 * warp + 8 * (4 * half + u) + 2 * t + e;
                bool m0, m1;
                if (mask_mode == 0) { m0 = mimi_masked_fast(mr0, c); m1 = mimi_masked_fast(mr1, c); }
                else { m0 = c >= M_CTX || mimi_masked(offset, g, c, mask_mode); m1 = c >= M_CTX || mimi_masked(offset, g + 8, c, mask_mode); }
                sc[e] = m0 ? -INFINITY : sc[e] * 0.125f;
                sc[2 + e] = m1 ? -INFINITY : sc[2 + e] * 0.125f;
            }
        }
    }
    // ---- row max over all 250 keys ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; j++) { mx0 = fmaxf(mx0, fmaxf(S[j][0], S[j][1])); mx1 = fmaxf(mx1, fmaxf(S[j][2], S[j][3])); }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    if (t == 0) { s_max[warp][g] = mx0; s_max[warp][g + 8] = mx1; }
    __syncthreads();
    mx0 = fmaxf(fmaxf(s_max[0][g], s_max[1][g]), fmaxf(s_max[2][g], s_max[3][g]));
    mx1 = fmaxf(fmaxf(s_max[0][g + 8], s_max[1][g + 8]), fmaxf(s_max[2][g + 8], s_max[3][g + 8]));
    // ---- exp and row sums (every row has at least its own position unmasked, so the max is finite) ----
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        S[j][0] = expf(S[j][0] - mx0); S[j][1] = expf(S[j][1] - mx0); S[j][2] = expf(S[j][2] - mx1); S[j][3] = expf(S[j][3] - mx1);
        l0 += S[j][0] + S[j][1]; l1 += S[j][2] + S[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if (t == 0) { s_sum[warp][g] = l0; s_sum[warp][g + 8] = l1; }
    __syncthreads();
    const float inv0 = 1.0f / (((s_sum[0][g] + s_sum[1][g]) + s_sum[2][g]) + s_sum[3][g]);
    const float inv1 = 1.0f / (((s_sum[0][g + 8] + s_sum[1][g + 8]) + s_sum[2][g + 8]) + s_sum[3][g + 8]);
    // ---- O_w = P_w V_w over this warp's 64 keys ----
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; n++) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
    for (int s4 = 0; s4 < 4; s4++) {
        uint32_t pa[4];
        {
            __nv_bfloat162 x;
            x = __floats2bfloat162_rn(S[2 * s4][0] * inv0, S[2 * s4][1] * inv0); pa[0] = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2bfloat162_rn(S[2 * s4][2] * inv1, S[2 * s4][3] * inv1); pa[1] = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2bfloat162_rn(S[2 * s4 + 1][0] * inv0, S[2 * s4 + 1][1] * inv0); pa[2] = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2bfloat162_rn(S[2 * s4 + 1][2] * inv1, S[2 * s4 + 1][3] * inv1); pa[3] = *reinterpret_cast<uint32_t*>(&x);
        }
        const int k0 = 64 * warp + 16 * s4 + 2 * t;
        const uint4 v0 = *reinterpret_cast<const uint4*>(V + (long long)min(k0, M_CTX - 1) * M_DIM + 8 * g);
        const uint4 v1 = *reinterpret_cast<const uint4*>(V + (long long)min(k0 + 1, M_CTX - 1) * M_DIM + 8 * g);
        const uint4 v2 = *reinterpret_cast<const uint4*>(V + (long long)min(k0 + 8, M_CTX - 1) * M_DIM + 8 * g);
        const uint4 v3 = *reinterpret_cast<const uint4*>(V + (long long)min(k0 + 9, M_CTX - 1) * M_DIM + 8 * g);
        const uint32_t a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w}, a2[4] = {v2.x, v2.y, v2.z, v2.w}, a3[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t b0_lo = __byte_perm(a0[w], a1[w], 0x5410), b1_lo = __byte_perm(a2[w], a3[w], 0x5410);
            const uint32_t b0_hi = __byte_perm(a0[w], a1[w], 0x7632), b1_hi = __byte_perm(a2[w], a3[w], 0x7632);
            mma_bf16_16816(o[2 * w], pa, b0_lo, b1_lo);
            mma_bf16_16816(o[2 * w + 1], pa, b0_hi, b1_hi);
        }
    }
    // ---- fixed-order sum of the four partial outputs: warps 1..3 publish, warp 0 adds them to its own (0 + 1 + 2 + 3) ----
    // o[n][e] = O[g][8(2t+e) + n], o[n][2+e] = O[g+8][...]: the lane owns dims 16t .. 16t+15 of rows g and g+8
    if (warp > 0) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                float* dst = &s_o[warp - 1][g + 8 * r][16 * t + 8 * e];
                *reinterpret_cast<float4*>(dst) = make_float4(o[0][2 * r + e], o[1][2 * r + e], o[2][2 * r + e], o[3][2 * r + e]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4][2 * r + e], o[5][2 * r + e], o[6][2 * r + e], o[7][2 * r + e]);
            }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            __nv_bfloat162 pk[8];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                float acc[8];
#pragma unroll
                for (int n = 0; n < 8; n++) acc[n] = o[n][2 * r + e];
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    const float* src = &s_o[w][g + 8 * r][16 * t + 8 * e];
                    const float4 x = *reinterpret_cast<const float4*>(src), y = *reinterpret_cast<const float4*>(src + 4);
                    acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w; acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
                }
#pragma unroll
                for (int n = 0; n < 8; n += 2) pk[e * 4 + n / 2] = __floats2bfloat162_rn(acc[n], acc[n + 1]);
            }
            __nv_bfloat16* dst = out + ((long long)b * M_T + g + 8 * r) * M_DIM + h * D_HEAD + 16 * t;
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(&pk[0]);
            *reinterpret_cast<uint4*>(dst + 8) = *reinterpret_cast<uint4*>(&pk[4]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Flow head glue
// ------------------------------------------------------------------------------------------------
// c = LN(h; out_norm) (bf16 copy for cond_embed) and EOS logit = out_eos(bf16(c)) + bias + 4
// (reference models/flow_lm.h:114-129). One CTA per row.
__global__ void __launch_bounds__(256) head_pre_kernel(const float* __restrict__ h, int R, const float* __restrict__ w, const float* __restrict__ b,
                                                       const __nv_bfloat16* __restrict__ w_eos, const float* __restrict__ b_eos,
                                                       __nv_bfloat16* __restrict__ c_bf16, float* __restrict__ eos) {
    pdl_prologue();
    __shared__ float red[3][8];
    const int row = blockIdx.x, col = threadIdx.x * 4, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (row >= R) return;
    const float4 a = *reinterpret_cast<const float4*>(h + (long long)row * D_MODEL + col);
    float v[4] = {a.x, a.y, a.z, a.w};
    const float s1 = warp_sum(v[0] + v[1] + v[2] + v[3]);
    if (lane == 0) red[0][warp] = s1;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += red[0][i];
    const float mean = tot / D_MODEL;
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) { v[i] -= mean; s2 += v[i] * v[i]; }
    s2 = warp_sum(s2);
    if (lane == 0) red[1][warp] = s2;
    __syncthreads();
    tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += red[1][i];
    const float rs = 1.0f / sqrtf(tot / D_MODEL + 1e-5f);
    float dot = 0.f;
    __nv_bfloat16 yb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float y = v[i] * rs * w[col + i];
        if (b) y += b[col + i];
        yb[i] = __float2bfloat16_rn(y);
        dot = fmaf(__bfloat162float(yb[i]), __bfloat162float(w_eos[col + i]), dot);
    }
    *reinterpret_cast<uint2*>(c_bf16 + (long long)row * D_MODEL + col) = *reinterpret_cast<uint2*>(yb);
    dot = warp_sum(dot);
    if (lane == 0) red[2][warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) d += red[2][i];
        eos[row] = d + (b_eos ? b_eos[0] : 0.f) + 4.0f;
    }
}

// noise -> (f32, bf16) per row. Noise is either injected by the caller (identical-noise parity runs, the
// reference's injection point is GraphContext::normal_, src/context.h:465-509) or drawn on the device from a
// counter-based generator keyed by (seed, slot, generation step): Philox-4x32-10 + Box-Muller, std = sqrt(temp).
__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One CTA (128 threads) per row: the first 32 threads draw the row's noise, then every thread computes 4 of the 512 outputs of
// input_proj(bf16(noise)) (32 -> 512, reference modules/mlp.h:236; bf16 operands, fp32 accumulation like every linear here).
__global__ void __launch_bounds__(128) noise_inproj_kernel(int slot0, int n, const float* __restrict__ injected, const unsigned long long* __restrict__ seed_ptr,
                                                           const float* __restrict__ temp, const int* __restrict__ gen_step, float* __restrict__ noise_f32,
                                                           const __nv_bfloat16* __restrict__ w_in_t, const float* __restrict__ b_in, float* __restrict__ xh) {
    pdl_prologue();
    __shared__ float zs[LDIM];
    const int r = blockIdx.x, slot = slot0 + r, i = threadIdx.x;
    if (r >= n) return;
    if (i < LDIM) {
        float z;
        if (injected) {
            z = injected[r * LDIM + i];
        } else {
            const float std = sqrtf(temp[slot]);
            if (std == 0.f) z = 0.f;
            else {
                uint32_t o[4];
                const unsigned long long seed = *seed_ptr;
                philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)slot, (uint32_t)gen_step[slot], (uint32_t)(i >> 1), 0x5054545Au, o);
                const float u1 = ((float)(o[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float u2 = ((float)(o[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float rad = sqrtf(-2.0f * logf(u1));
                float sn, cn; sincosf(6.28318530717958647692f * u2, &sn, &cn);
                z = ((i & 1) ? rad * sn : rad * cn) * std;
            }
        }
        noise_f32[r * LDIM + i] = z;
        zs[i] = __bfloat162float(__float2bfloat16_rn(z));
    }
    __syncthreads();
    // w_in_t is input_proj.weight transposed to [32][512]: thread i reads 8 contiguous bytes per input k, a warp 256 contiguous bytes
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < LDIM; k++) {
        const uint2 wv = __ldg(reinterpret_cast<const uint2*>(w_in_t + (long long)k * D_FLOW + i * 4));
        const float z = zs[k];
        acc[0] = fmaf(__uint_as_float(wv.x << 16), z, acc[0]); acc[1] = fmaf(__uint_as_float(wv.x & 0xffff0000u), z, acc[1]);
        acc[2] = fmaf(__uint_as_float(wv.y << 16), z, acc[2]); acc[3] = fmaf(__uint_as_float(wv.y & 0xffff0000u), z, acc[3]);
    }
    if (b_in) { const float4 b = *reinterpret_cast<const float4*>(b_in + i * 4); acc[0] += b.x; acc[1] += b.y; acc[2] += b.z; acc[3] += b.w; }
    *reinterpret_cast<float4*>(xh + (long long)r * D_FLOW + i * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// Decode-step entry of the FlowLM backbone, one CTA (256 threads) per utterance: h = input_linear(bf16(previous latent)) (32 -> 1024,
// reference models/flow_lm.h:99) followed by layer 0's norm1 (src/torch.h:49-60) -> bf16 A operand of the first in_proj.
__global__ void __launch_bounds__(256) flow_in_kernel(int slot0, int n, const __nv_bfloat16* __restrict__ lat_in, const __nv_bfloat16* __restrict__ w_in_t,
                                                      const float* __restrict__ b_in, const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                      float* __restrict__ h, __nv_bfloat16* __restrict__ n_bf,
                                                      const int* __restrict__ cur_len, const float* __restrict__ freq,
                                                      int* __restrict__ row_slot, int* __restrict__ row_pos, float2* __restrict__ cs) {
    pdl_prologue();
    __shared__ float xs[LDIM];
    __shared__ float red[2][8];
    const int r = blockIdx.x, i = threadIdx.x, warp = i >> 5, lane = i & 31;
    if (r >= n) return;
    if (i < LDIM) xs[i] = __bfloat162float(lat_in[(long long)(slot0 + r) * LDIM + i]);
    if (warp == 1) {                                           // row bookkeeping for the QKV epilogues: slot, position, RoPE table
        const int pos = cur_len[slot0 + r];
        if (lane == 0) { row_slot[r] = slot0 + r; row_pos[r] = pos; }
        const float rad = (float)pos * freq[lane];
        cs[r * 32 + lane] = make_float2(cosf(rad), sinf(rad));
    }
    __syncthreads();
    float v[4] = {0.f, 0.f, 0.f, 0.f};                         // w_in_t = input_linear.weight transposed to [32][1024] (coalesced)
#pragma unroll
    for (int k = 0; k < LDIM; k++) {
        const uint2 wv = __ldg(reinterpret_cast<const uint2*>(w_in_t + (long long)k * D_MODEL + i * 4));
        const float z = xs[k];
        v[0] = fmaf(__uint_as_float(wv.x << 16), z, v[0]); v[1] = fmaf(__uint_as_float(wv.x & 0xffff0000u), z, v[1]);
        v[2] = fmaf(__uint_as_float(wv.y << 16), z, v[2]); v[3] = fmaf(__uint_as_float(wv.y & 0xffff0000u), z, v[3]);
    }
    if (b_in) { const float4 b = *reinterpret_cast<const float4*>(b_in + i * 4); v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w; }
    *reinterpret_cast<float4*>(h + (long long)r * D_MODEL + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
    const float s1 = warp_sum(v[0] + v[1] + v[2] + v[3]);
    if (lane == 0) red[0][warp] = s1;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) tot += red[0][w];
    const float mean = tot / D_MODEL;
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) { v[j] -= mean; s2 += v[j] * v[j]; }
    s2 = warp_sum(s2);
    if (lane == 0) red[1][warp] = s2;
    __syncthreads();
    tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) tot += red[1][w];
    const float rs = 1.0f / sqrtf(tot / D_MODEL + 1e-5f);
    __nv_bfloat16 yb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { float y = v[j] * rs * lnw[i * 4 + j]; if (lnb) y += lnb[i * 4 + j]; yb[j] = __float2bfloat16_rn(y); }
    *reinterpret_cast<uint2*>(n_bf + (long long)r * D_MODEL + i * 4) = *reinterpret_cast<uint2*>(yb);
}

// ------------------------------------------------------------------------------------------------
// Mimi front end: z = latent*emb_std + emb_mean; e = Wq . f16(z) (1x1 conv, f16 operands; reference
// src/pocket_tts.cpp:472-478, models/mimi.h:77-83); depthwise x16 upsampler on one step with carried state
// (modules/conv.h:283-331): x[k][c] = e[c]*w[c][k] + e_prev[c]*w[c][16+k] (+bias). State kept = e_prev.
// ------------------------------------------------------------------------------------------------
// wq_t is the quantizer projection transposed to [32][512], wup_t the upsampler taps transposed to [32][512]: lane = channel, so
// every weight load and every store is a coalesced line.
__global__ void __launch_bounds__(512) mimi_front_kernel(int slot0, const float* __restrict__ lat_f32, const float* __restrict__ emb_std,
                                                         const float* __restrict__ emb_mean, const __half* __restrict__ wq_t,
                                                         const float* __restrict__ wup_t, const float* __restrict__ bup,
                                                         float* __restrict__ e_prev, float* __restrict__ x) {
    pdl_prologue();
    __shared__ float z[LDIM];
    const int slot = slot0 + blockIdx.x, c = threadIdx.x;
    if (c < LDIM) z[c] = __half2float(__float2half_rn(__fadd_rn(__fmul_rn(emb_std[c], lat_f32[slot * LDIM + c]), emb_mean[c])));
    __syncthreads();
    float e = 0.f;
#pragma unroll
    for (int i = 0; i < LDIM; i++) e = fmaf(__half2float(wq_t[i * M_DIM + c]), z[i], e);
    const float ep = e_prev[(long long)slot * M_DIM + c];
    e_prev[(long long)slot * M_DIM + c] = e;
    const float bias = bup ? bup[c] : 0.f;
    float* xo = x + (long long)slot * M_T * M_DIM + c;
#pragma unroll
    for (int k = 0; k < M_T; k++) {
        const float y = __fadd_rn(__fmul_rn(e, wup_t[k * M_DIM + c]), __fmul_rn(ep, wup_t[(16 + k) * M_DIM + c]));
        xo[(long long)k * M_DIM] = y + bias;
    }
}

// Stop rule + bookkeeping after the head (reference src/pocket_tts.cpp:457-467,487-489) and the Mimi front end in one launch. Per slot:
//   eos_step = first step with logit+4 > 0; stop when gen_step >= eos_step + frames_after_eos; hard cap max_gen_len.
// produced[r] = 1 when this step emits a frame. Hands the new latent to the next step (bf16 copy = the A operand of input_linear),
// advances the FlowLM position, then runs mimi_front_kernel's math on the latent this step produced.
__global__ void __launch_bounds__(512) step_front_kernel(int slot0, int n, const float* __restrict__ eos, const float* __restrict__ latent,
                                                         int* __restrict__ cur_len, int* __restrict__ gen_step, int* __restrict__ eos_step,
                                                         const int* __restrict__ max_gen, const int* __restrict__ fae, int* __restrict__ active,
                                                         __nv_bfloat16* __restrict__ lat_in_bf16, float* __restrict__ lat_f32, int* __restrict__ produced,
                                                         float* __restrict__ eos_out, const float* __restrict__ emb_std, const float* __restrict__ emb_mean,
                                                         const __half* __restrict__ wq_t, const float* __restrict__ wup_t, const float* __restrict__ bup,
                                                         float* __restrict__ e_prev, float* __restrict__ x) {
    pdl_prologue();
    __shared__ float z[LDIM];
    __shared__ int emit;
    const int r = blockIdx.x, slot = slot0 + r, c = threadIdx.x;
    if (r >= n) return;
    if (c == 0) {
        int e = 0;
        if (active[slot]) {
            const int g = gen_step[slot];
            int es = eos_step[slot];
            if (eos[r] > 0.f && es == -1) es = g;
            eos_step[slot] = es;
            cur_len[slot] += 1;                                   // increment_states (pocket_tts.cpp:96)
            if (es != -1 && g >= es + fae[slot]) { gen_step[slot] = max_gen[slot]; active[slot] = 0; }
            else {
                e = 1; gen_step[slot] = g + 1;
                if (g + 1 >= max_gen[slot]) active[slot] = 0;     // next receive would hit the cap (pocket_tts.cpp:450-453,495)
            }
        }
        emit = e; produced[r] = e; if (eos_out) eos_out[r] = eos[r];
    }
    __syncthreads();
    if (c < LDIM) {
        float v;
        if (emit) { v = latent[r * LDIM + c]; lat_f32[slot * LDIM + c] = v; lat_in_bf16[slot * LDIM + c] = __float2bfloat16_rn(v); }
        else v = lat_f32[slot * LDIM + c];
        z[c] = __half2float(__float2half_rn(__fadd_rn(__fmul_rn(emb_std[c], v), emb_mean[c])));
    }
    __syncthreads();
    float e = 0.f;
#pragma unroll
    for (int i = 0; i < LDIM; i++) e = fmaf(__half2float(wq_t[i * M_DIM + c]), z[i], e);
    const float ep = e_prev[(long long)slot * M_DIM + c];
    e_prev[(long long)slot * M_DIM + c] = e;
    const float bias = bup ? bup[c] : 0.f;
    float* xo = x + (long long)slot * M_T * M_DIM + c;
#pragma unroll
    for (int k = 0; k < M_T; k++) {
        const float y = __fadd_rn(__fmul_rn(e, wup_t[k * M_DIM + c]), __fmul_rn(ep, wup_t[(16 + k) * M_DIM + c]));
        xo[(long long)k * M_DIM] = y + bias;
    }
}

// End-of-step upkeep for the streaming convs: every conv input buffer is [slot][S + T][C] with the S carried
// rows first (reference modules/conv.h:60-76 keeps the last K-stride inputs; the transposed convs keep the last
// input row instead of the reference's partial output). Copies the last S rows to the front.
struct ShiftDesc { __half* buf; long long slot_stride; int S, T, C; };
struct ShiftAll { ShiftDesc d[8]; int n; };
__global__ void shift_states_kernel(ShiftAll sa, int slot0, int* __restrict__ mimi_off) {
    pdl_prologue();
    const ShiftDesc d = sa.d[blockIdx.y];
    const int slot = slot0 + blockIdx.x;
    __half* base = d.buf + (long long)slot * d.slot_stride;
    const int n8 = d.S * d.C / 8;                               // 16-byte pieces (every C here is a multiple of 64)
    // source rows [T, T+S) and destination rows [0, S) never overlap because T >= S for every conv here.
    const uint4* src = reinterpret_cast<const uint4*>(base + (long long)d.T * d.C);
    uint4* dst = reinterpret_cast<uint4*>(base);
    for (int i = threadIdx.x; i < n8; i += blockDim.x) dst[i] = src[i];
    if (blockIdx.y == 0 && threadIdx.x == 0) mimi_off[slot] += M_T;
}

// Sentence start (reference src/pocket_tts.cpp:416-444, models/mimi.h:71-75): zero the carried conv state rows
// and the upsampler state, reset the Mimi offset. The KV prefix restore is done with device copies by the host.
__global__ void reset_slot_kernel(ShiftAll sa, int slot, float* __restrict__ e_prev, int* __restrict__ mimi_off) {
    pdl_prologue();
    const ShiftDesc d = sa.d[blockIdx.y];
    __half* base = d.buf + (long long)slot * d.slot_stride;
    for (int i = threadIdx.x; i < d.S * d.C; i += blockDim.x) base[i] = __float2half_rn(0.f);
    if (blockIdx.y == 0) {
        for (int i = threadIdx.x; i < M_DIM; i += blockDim.x) e_prev[(long long)slot * M_DIM + i] = 0.f;
        if (threadIdx.x == 0) mimi_off[slot] = 0;
    }
}

}  // namespace ptts
