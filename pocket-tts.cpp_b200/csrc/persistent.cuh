// persistent.cuh — the whole FlowLM decode step + flow/LSD head for 1-2 utterances as ONE cooperative kernel.
//
// At batch 1-2 the step is ~75 dependent launches of 3-5 us each (0.35 ms per frame, measured) for 170 MB of weights that HBM
// delivers in 30 us: launch / prologue latency, not bytes, sets the frame rate of the streaming API (BASELINE config 2). Here one
// persistent grid (one CTA per SM, cooperative launch) walks the same phases with grid-wide barriers in between:
//   per layer: [LN1 + in_proj + RoPE + KV append] | [attention partials per (row, head, key split)] | [merge + out_proj + residual] |
//              [LN2 + linear1 + GELU] | [linear2 + residual]
//   head:      [out_norm + EOS + noise/input_proj + cond_embed] | [7 adaLN projections] | 6 x ([LN+modulate + mlp.0 + SiLU] | [mlp.2 * gate + res]) |
//              [final LN+modulate + linear + noise]
// Small per-row work (LayerNorms, the 32-wide input projections, the attention merge, RoPE tables) is recomputed by every CTA in shared
// memory instead of being exchanged; every matrix-vector product is spread over all warps of the grid (one warp = two output columns,
// 128-bit weight loads, shuffle reductions) exactly like gemv_small_kernel / gemv_ln_kernel, with the same rounding points (activations
// rounded to bf16 before every product, fp32 accumulation, ggml f16-table GELU). Reference: models/flow_lm.h:84-147,
// modules/transformer.h:55-199,253-278, modules/mlp.h:233-251.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace ptts {

constexpr int PF_RMAX = 2;            // utterances per launch
constexpr int PF_THREADS = 256;
constexpr int PF_MAX_KEYS = 2048;     // keys of one attention task (capacity / splits)

struct PfLin { const __nv_bfloat16* w; const float* b; };
struct PfParams {
    int slot0, R, n_splits;
    // per-slot state
    const __nv_bfloat16* lat_in; const int* cur_len; const int* active; const float* freq;
    __nv_bfloat16 *kc, *vc; long long kv_slot_stride, kv_layer_stride;
    const int *pfx_slot, *pfx_len;
    // FlowLM weights
    const __nv_bfloat16* input_linear_t; const float* input_linear_b;
    struct { PfLin in_proj, out_proj, lin1, lin2; const float *n1w, *n1b, *n2w, *n2b; } L[N_LAYERS];
    // head
    const float *onw, *onb; const __nv_bfloat16* w_eos; const float* b_eos;
    PfLin cond, ada, fin; int ada_out; const float* t_combined;
    struct { const float *lnw, *lnb; PfLin m0, m2; } rb[N_RES];
    const float *fnw, *fnb;
    const __nv_bfloat16* input_proj_t; const float* input_proj_b;
    // noise
    const float* injected; const unsigned long long* seed; const float* temp; const int* gen_step; const unsigned int* rng_id;
    // global scratch
    float *h, *q, *ws_ml, *ws_acc, *mod, *xh, *noise_f32, *latent, *eos;
    __nv_bfloat16 *ff_bf, *sy_bf, *h1_bf;
    unsigned int* barrier;   // [0] arrival counter of PfBarrier (never reset), [1] its value when the current launch started (written by the previous launch)
};

// y[n0 .. n0+NC) for every row: dot products of NC weight rows with the rows of xs (smem, already rounded to bf16), spread over the warps of
// the whole grid. NC = 2 columns per warp while that keeps every warp busy; matrices with fewer column pairs than warps (N = 1024, 512, 32)
// use one column per warp, which halves the weight bytes on the critical path of the phase (the slowest warp sets the barrier time).
template <int K, int NC, typename F>
__device__ __forceinline__ void pf_gemv_nc(const __nv_bfloat16* __restrict__ W, int N, const float* xs /*[PF_RMAX][K]*/, int R, F&& epi) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * PF_THREADS + threadIdx.x) >> 5, nw = (gridDim.x * PF_THREADS) >> 5;
    for (int unit = gw; unit < N / NC; unit += nw) {
        const int n0 = NC * unit;
        const __nv_bfloat16* w0 = W + (long long)n0 * K;
        float acc[PF_RMAX][NC];
#pragma unroll
        for (int r = 0; r < PF_RMAX; r++)
#pragma unroll
            for (int c = 0; c < NC; c++) acc[r][c] = 0.f;
#pragma unroll 4
        for (int k = lane * 8; k < K; k += 256) {
            uint4 wv[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) wv[c] = __ldg(reinterpret_cast<const uint4*>(w0 + (long long)c * K + k));
#pragma unroll
            for (int r = 0; r < PF_RMAX; r++) {
                if (r < R) {
                    const float4 xa = *reinterpret_cast<const float4*>(xs + r * K + k), xb = *reinterpret_cast<const float4*>(xs + r * K + k + 4);
                    const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const uint32_t a[4] = {wv[c].x, wv[c].y, wv[c].z, wv[c].w};
#pragma unroll
                        for (int j = 0; j < 4; j++) { acc[r][c] = fmaf(x[2 * j], __uint_as_float(a[j] << 16), acc[r][c]); acc[r][c] = fmaf(x[2 * j + 1], __uint_as_float(a[j] & 0xffff0000u), acc[r][c]); }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < PF_RMAX; r++) {
            if (r < R) {
                float v[2] = {0.f, 0.f};
#pragma unroll
                for (int c = 0; c < NC; c++) v[c] = warp_sum(acc[r][c]);
                if (lane == 0) epi(r, n0, v[0], v[1], NC);
            }
        }
    }
}
// epi(r, n0, v0, v1) is called once per column PAIR (RoPE rotates pairs): with one column per warp the pair is completed by a shuffle
template <int K, typename F>
__device__ __forceinline__ void pf_gemv(const __nv_bfloat16* __restrict__ W, int N, const float* xs, int R, F&& epi, bool pairs_needed = false) {
    const int nw = (gridDim.x * PF_THREADS) >> 5;
    if (N / 2 >= nw || pairs_needed) pf_gemv_nc<K, 2>(W, N, xs, R, [&](int r, int n0, float v0, float v1, int) { epi(r, n0, v0, v1, 2); });
    else pf_gemv_nc<K, 1>(W, N, xs, R, [&](int r, int n0, float v0, float, int) { epi(r, n0, v0, 0.f, 1); });
}

// L2 prefetch of the weight rows this warp will read in the NEXT phase (issued before the grid barrier: hides the cold-DRAM latency)
__device__ __forceinline__ void pf_prefetch(const __nv_bfloat16* W, int N, int K) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * PF_THREADS + threadIdx.x) >> 5, nw = (gridDim.x * PF_THREADS) >> 5;
    for (int pair = gw; pair < N / 2; pair += nw) {
        const char* p = reinterpret_cast<const char*>(W + (long long)2 * pair * K);
        for (int off = lane * 128; off < 2 * K * 2; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
    }
}

// LayerNorm (ggml_norm semantics, see layernorm_kernel) of R rows of C floats from global memory into xs (rounded to bf16), whole CTA
template <int C>
__device__ __forceinline__ void pf_layernorm(const float* __restrict__ x, long long ld, int R, float eps, const float* w, const float* b,
                                             const float* shift, const float* scale, long long mod_ld, float* xs, float* red) {
    constexpr int PER = C / PF_THREADS;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    for (int r = 0; r < R; r++) {
        float v[PER]; float s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; i++) { v[i] = x[(long long)r * ld + tid * PER + i]; s += v[i]; }
        s = warp_sum(s);
        if (lane == 0) red[wid] = s;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) tot += red[i];
        const float mean = tot / C;
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < PER; i++) { v[i] -= mean; s2 += v[i] * v[i]; }
        s2 = warp_sum(s2);
        __syncthreads();
        if (lane == 0) red[wid] = s2;
        __syncthreads();
        tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) tot += red[i];
        const float rs = 1.0f / sqrtf(tot / C + eps);
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int c = tid * PER + i;
            float y = v[i] * rs;
            if (w) y *= w[c];
            if (b) y += b[c];
            if (scale) y = y * (scale[(long long)r * mod_ld + c] + 1.f) + shift[(long long)r * mod_ld + c];
            xs[r * C + c] = __bfloat162float(__float2bfloat16_rn(y));
        }
        __syncthreads();
    }
}

// Grid-wide barrier for a cooperative launch (all CTAs resident): one arrival counter, monotonically increasing; thread 0 of every CTA
// arrives with a release fence and spins on an acquire load until the phase's target is reached. ~1 us on 148 CTAs, about a third of
// cooperative_groups' grid.sync() (which also has to work for multi-device grids).
struct PfBarrier {
    unsigned int* counter; unsigned int target;
    __device__ __forceinline__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            __threadfence();
            atomicAdd(counter, 1u);
            unsigned int v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
        }
        __syncthreads();
    }
};

__global__ void __launch_bounds__(PF_THREADS, 2) flow_persistent_kernel(const PfParams p) {
    PfBarrier grid; grid.counter = p.barrier; grid.target = p.barrier[1];          // the count all earlier launches left behind (stable: written after their last barrier)
    extern __shared__ __align__(16) float pf_smem[];
    float* xs = pf_smem;                                   // [PF_RMAX][4096] GEMV input rows
    float* sc = xs + PF_RMAX * D_FF;                       // [PF_MAX_KEYS] attention scores of one task | scratch
    float* red = sc + PF_MAX_KEYS;                         // [16]
    float2* cs = reinterpret_cast<float2*>(red + 16);      // [PF_RMAX][32] RoPE table
    int* rs_slot = reinterpret_cast<int*>(cs + PF_RMAX * 32);   // [PF_RMAX] row -> slot (-1 = dead row), [PF_RMAX] position
    int* rs_pos = rs_slot + PF_RMAX;
    float* zs = reinterpret_cast<float*>(rs_pos + PF_RMAX);     // [PF_RMAX][32] noise (bf16-rounded)
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int R = p.R;

    // ---- phase 0 (every CTA): row bookkeeping, RoPE table, h = input_linear(bf16(previous latent)), norm1 of layer 0 ----
    if (tid < R) { const int slot = p.slot0 + tid; rs_slot[tid] = p.active[slot] ? slot : -1; rs_pos[tid] = p.cur_len[slot]; }
    if (tid < R * 32) { const int r = tid >> 5; const float rad = (float)p.cur_len[p.slot0 + r] * p.freq[tid & 31]; cs[tid] = make_float2(cosf(rad), sinf(rad)); }
    if (tid < R * LDIM) sc[tid] = __bfloat162float(p.lat_in[(long long)(p.slot0 + tid / LDIM) * LDIM + tid % LDIM]);
    __syncthreads();
    float* h0 = xs + PF_RMAX * D_MODEL;                    // f32 staging of h inside xs' upper half (xs holds only 1024-wide rows until linear2)
    for (int r = 0; r < R; r++) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < LDIM; k++) {
            const uint2 wv = __ldg(reinterpret_cast<const uint2*>(p.input_linear_t + (long long)k * D_MODEL + tid * 4));
            const float z = sc[r * LDIM + k];
            v[0] = fmaf(__uint_as_float(wv.x << 16), z, v[0]); v[1] = fmaf(__uint_as_float(wv.x & 0xffff0000u), z, v[1]);
            v[2] = fmaf(__uint_as_float(wv.y << 16), z, v[2]); v[3] = fmaf(__uint_as_float(wv.y & 0xffff0000u), z, v[3]);
        }
        if (p.input_linear_b) { const float4 b = *reinterpret_cast<const float4*>(p.input_linear_b + tid * 4); v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w; }
        *reinterpret_cast<float4*>(h0 + r * D_MODEL + tid * 4) = make_float4(v[0], v[1], v[2], v[3]);
        if (blockIdx.x == 0) *reinterpret_cast<float4*>(p.h + (long long)r * D_MODEL + tid * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    pf_layernorm<D_MODEL>(h0, D_MODEL, R, 1e-5f, p.L[0].n1w, p.L[0].n1b, nullptr, nullptr, 0, xs, red);

    for (int l = 0; l < N_LAYERS; l++) {
        const auto& L = p.L[l];
        __nv_bfloat16* kc = p.kc + l * p.kv_layer_stride;
        __nv_bfloat16* vc = p.vc + l * p.kv_layer_stride;
        // ---- P1: in_proj + RoPE + KV append (reference transformer.h:64-153, rope.h:183-272) ----
        if (l > 0) pf_layernorm<D_MODEL>(p.h, D_MODEL, R, 1e-5f, L.n1w, L.n1b, nullptr, nullptr, 0, xs, red);
        {
            Epi e; e.mode = EPI_FLOW_QKV; e.row_slot = rs_slot; e.row_pos = rs_pos; e.cs = cs; e.kv_f32 = 0; e.kv_slot_stride = p.kv_slot_stride;
            e.q_out_f32 = p.q; e.kcache = kc; e.vcache = vc; e.bias = L.in_proj.b;
            pf_gemv<D_MODEL>(L.in_proj.w, 3 * D_MODEL, xs, R, [&](int r, int n0, float v0, float v1, int) { const float v[2] = {v0, v1}; epi_apply<2>(e, r, n0, v, 3 * D_MODEL); }, true);
        }
        pf_prefetch(L.out_proj.w, D_MODEL, D_MODEL);
        grid.sync();
        // ---- P2: attention partials, one CTA per (row, head, key split) (reference transformer.h:157-199, torch.h:128-150) ----
        const int S = p.n_splits;
        for (int task = blockIdx.x; task < R * N_HEADS * S; task += gridDim.x) {
            const int r = task / (N_HEADS * S), hd = (task / S) % N_HEADS, sp = task % S;
            const int slot = rs_slot[r];
            float* wml = p.ws_ml + (long long)task * 2;
            float* wacc = p.ws_acc + (long long)task * D_HEAD;
            if (slot < 0) continue;
            const int len = rs_pos[r] + 1;
            const int P = p.pfx_len ? min(p.pfx_len[slot], len) : 0;
            const int pslot = p.pfx_len ? p.pfx_slot[slot] : slot;
            int chunk = (len + S - 1) / S;
            const int k0 = sp * chunk, k1 = min(len, k0 + chunk), nk = max(0, k1 - k0);
            __syncthreads();
            float* qs = zs;                                    // 64 floats of scratch: q of this head (zs is only used by the head phases)
            if (tid < D_HEAD) qs[tid] = p.q[(long long)r * D_MODEL + hd * D_HEAD + tid] * 0.125f;
            __syncthreads();
            float mx = -INFINITY;
            for (int j = tid; j < nk; j += PF_THREADS) {
                const int key = k0 + j;
                const __nv_bfloat16* kr = kc + (long long)(key < P ? pslot : slot) * p.kv_slot_stride + (long long)key * D_MODEL + hd * D_HEAD;
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < D_HEAD; d += 8) {
                    const uint4 kv = *reinterpret_cast<const uint4*>(kr + d);
                    const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
                    for (int t = 0; t < 4; t++) { a = fmaf(__uint_as_float(w[t] << 16), qs[d + 2 * t], a); a = fmaf(__uint_as_float(w[t] & 0xffff0000u), qs[d + 2 * t + 1], a); }
                }
                sc[j] = a; mx = fmaxf(mx, a);
            }
            mx = warp_max(mx);
            if (lane == 0) red[wid] = mx;
            __syncthreads();
            mx = red[0];
#pragma unroll
            for (int i = 1; i < 8; i++) mx = fmaxf(mx, red[i]);
            __syncthreads();
            float sum = 0.f;
            for (int j = tid; j < nk; j += PF_THREADS) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
            sum = warp_sum(sum);
            if (lane == 0) red[wid] = sum;
            __syncthreads();
            sum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) sum += red[i];
            // P V: warp w takes keys w, w+8, ...; lane owns dims 2*lane, 2*lane+1
            float a0 = 0.f, a1 = 0.f;
            for (int j = wid; j < nk; j += 8) {
                const int key = k0 + j;
                const __nv_bfloat16* vr = vc + (long long)(key < P ? pslot : slot) * p.kv_slot_stride + (long long)key * D_MODEL + hd * D_HEAD + 2 * lane;
                const uint32_t vv = *reinterpret_cast<const uint32_t*>(vr);
                const float pj = sc[j];
                a0 = fmaf(pj, __uint_as_float(vv << 16), a0); a1 = fmaf(pj, __uint_as_float(vv & 0xffff0000u), a1);
            }
            __syncthreads();                                   // sc is reused as the cross-warp buffer
            sc[wid * D_HEAD + 2 * lane] = a0; sc[wid * D_HEAD + 2 * lane + 1] = a1;
            __syncthreads();
            if (tid < D_HEAD) {
                float o = 0.f;
#pragma unroll
                for (int w = 0; w < 8; w++) o += sc[w * D_HEAD + tid];
                wacc[tid] = o;
            }
            if (tid == 0) { wml[0] = nk > 0 ? mx : -INFINITY; wml[1] = nk > 0 ? sum : 0.f; }
        }
        pf_prefetch(L.lin1.w, D_FF, D_MODEL);
        grid.sync();
        // ---- P3: merge the key splits (fixed order) -> bf16 attention output; out_proj + residual ----
        // split weights exp(m_sp - M) and the sum L once per (row, head) (R x 16 threads), then 4 outputs per thread with independent loads: every CTA
        // repeats this merge, and with the weights recomputed per output it was the longest phase of a layer (8.2 us against 4.8-5.8 for the others)
        float* mw = sc;                                        // [R * 16][16] weights | [R * 16] L | [R * 16 * S] (m, l) pairs   (sc is free between the attention phases)
        float* mL = sc + PF_RMAX * N_HEADS * 16;
        float2* mp = reinterpret_cast<float2*>(mL + PF_RMAX * N_HEADS);
        for (int i = tid; i < R * N_HEADS * S; i += PF_THREADS) mp[i] = *reinterpret_cast<const float2*>(p.ws_ml + (long long)i * 2);   // one round trip for all (m, l)
        __syncthreads();
        if (tid < R * N_HEADS) {
            const int r = tid / N_HEADS;
            float M = -INFINITY, Lsum = 0.f;
            if (rs_slot[r] >= 0) {
                for (int sp = 0; sp < S; sp++) M = fmaxf(M, mp[tid * S + sp].x);
                for (int sp = 0; sp < S; sp++) {
                    const float ms = mp[tid * S + sp].x;
                    const float w = ms == -INFINITY ? 0.f : expf(ms - M);
                    Lsum = fmaf(mp[tid * S + sp].y, w, Lsum);
                    mw[tid * 16 + sp] = w;
                }
            }
            mL[tid] = Lsum;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < PF_RMAX * D_MODEL / PF_THREADS; it++) {    // every load below is independent of the others: one round trip for all partial outputs
            const int i = tid + it * PF_THREADS;
            if (i < R * D_MODEL) {
                const int r = i / D_MODEL, c = i % D_MODEL, hd = c / D_HEAD, d = c % D_HEAD;
                float o = 0.f;
                if (rs_slot[r] >= 0) {
                    const int rh = r * N_HEADS + hd;
                    const float* ap = p.ws_acc + (long long)rh * S * D_HEAD + d;
                    float a[16];
#pragma unroll
                    for (int sp = 0; sp < 16; sp++) a[sp] = sp < S ? ap[sp * D_HEAD] : 0.f;
#pragma unroll
                    for (int sp = 0; sp < 16; sp++) if (sp < S) o = fmaf(a[sp], mw[rh * 16 + sp], o);
                    o /= mL[rh];
                }
                xs[r * D_MODEL + c] = __bfloat162float(__float2bfloat16_rn(o));
            }
        }
        __syncthreads();
        pf_gemv<D_MODEL>(L.out_proj.w, D_MODEL, xs, R, [&](int r, int n0, float v0, float v1, int nc) {
            float* hp = p.h + (long long)r * D_MODEL + n0;
            hp[0] += v0 + (L.out_proj.b ? L.out_proj.b[n0] : 0.f);
            if (nc == 2) hp[1] += v1 + (L.out_proj.b ? L.out_proj.b[n0 + 1] : 0.f);
        });
        pf_prefetch(L.lin2.w, D_MODEL, D_FF);
        grid.sync();
        // ---- P4: norm2 + linear1 + GELU (reference transformer.h:266-272) ----
        pf_layernorm<D_MODEL>(p.h, D_MODEL, R, 1e-5f, L.n2w, L.n2b, nullptr, nullptr, 0, xs, red);
        pf_gemv<D_MODEL>(L.lin1.w, D_FF, xs, R, [&](int r, int n0, float v0, float v1, int nc) {
            __nv_bfloat16* fp = p.ff_bf + (long long)r * D_FF + n0;
            fp[0] = __float2bfloat16_rn(gelu_ggml(v0 + (L.lin1.b ? L.lin1.b[n0] : 0.f)));
            if (nc == 2) fp[1] = __float2bfloat16_rn(gelu_ggml(v1 + (L.lin1.b ? L.lin1.b[n0 + 1] : 0.f)));
        });
        if (l + 1 < N_LAYERS) pf_prefetch(p.L[l + 1].in_proj.w, 3 * D_MODEL, D_MODEL);
        grid.sync();
        // ---- P5: linear2 + residual ----
        for (int i = tid; i < R * D_FF; i += PF_THREADS) xs[i] = __bfloat162float(p.ff_bf[i]);
        __syncthreads();
        pf_gemv<D_FF>(L.lin2.w, D_MODEL, xs, R, [&](int r, int n0, float v0, float v1, int nc) {
            float* hp = p.h + (long long)r * D_MODEL + n0;
            hp[0] += v0 + (L.lin2.b ? L.lin2.b[n0] : 0.f);
            if (nc == 2) hp[1] += v1 + (L.lin2.b ? L.lin2.b[n0 + 1] : 0.f);
        });
        if (l + 1 == N_LAYERS) pf_prefetch(p.cond.w, D_FLOW, D_MODEL);
        grid.sync();
    }

    // ---- head H1: c = out_norm(h) (bf16), EOS logit, noise + input_proj, cond_embed (+ t_combined, SiLU) (flow_lm.h:114-140, mlp.h:233-245) ----
    pf_layernorm<D_MODEL>(p.h, D_MODEL, R, 1e-5f, p.onw, p.onb, nullptr, nullptr, 0, xs, red);
    const int eos_blk = min(1, (int)gridDim.x - 1), xh_blk = min(2, (int)gridDim.x - 1);   // block 0's extras on other CTAs: the slowest CTA sets the barrier time
    if ((int)blockIdx.x == eos_blk) {
        for (int r = 0; r < R; r++) {
            float d = 0.f;
            for (int c = tid; c < D_MODEL; c += PF_THREADS) d = fmaf(xs[r * D_MODEL + c], __bfloat162float(p.w_eos[c]), d);
            d = warp_sum(d);
            __syncthreads();
            if (lane == 0) red[wid] = d;
            __syncthreads();
            if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; i++) t += red[i]; p.eos[r] = t + (p.b_eos ? p.b_eos[0] : 0.f) + 4.0f; }
        }
    }
    if (tid < R * LDIM) {                                      // noise: injected or Philox keyed by (seed, sentence stream id, step), see noise_inproj_kernel
        const int r = tid / LDIM, i = tid % LDIM, slot = p.slot0 + r;
        float z;
        if (p.injected) z = p.injected[r * LDIM + i];
        else {
            const float sd = sqrtf(p.temp[slot]);
            if (sd == 0.f) z = 0.f;
            else {
                uint32_t o[4];
                const unsigned long long seed = *p.seed;
                philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), p.rng_id[slot], (uint32_t)p.gen_step[slot], (uint32_t)(i >> 1), 0x5054545Au, o);
                const float u1 = ((float)(o[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(o[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float rad = sqrtf(-2.0f * logf(u1));
                float sn, cn; sincosf(6.28318530717958647692f * u2, &sn, &cn);
                z = ((i & 1) ? rad * sn : rad * cn) * sd;
            }
        }
        if (blockIdx.x == 0) p.noise_f32[r * LDIM + i] = z;
        zs[r * LDIM + i] = __bfloat162float(__float2bfloat16_rn(z));
    }
    __syncthreads();
    if ((int)blockIdx.x == xh_blk) {                           // xh = input_proj(bf16(noise)) (32 -> 512)
        for (int i = tid; i < R * D_FLOW; i += PF_THREADS) {
            const int r = i / D_FLOW, c = i % D_FLOW;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < LDIM; k++) a = fmaf(__bfloat162float(p.input_proj_t[(long long)k * D_FLOW + c]), zs[r * LDIM + k], a);
            p.xh[(long long)r * D_FLOW + c] = a + (p.input_proj_b ? p.input_proj_b[c] : 0.f);
        }
    }
    pf_gemv<D_MODEL>(p.cond.w, D_FLOW, xs, R, [&](int r, int n0, float v0, float v1, int nc) {
        __nv_bfloat16* sp = p.sy_bf + (long long)r * D_FLOW + n0;
        sp[0] = __float2bfloat16_rn(silu_f(v0 + (p.cond.b ? p.cond.b[n0] : 0.f) + p.t_combined[n0]));
        if (nc == 2) sp[1] = __float2bfloat16_rn(silu_f(v1 + (p.cond.b ? p.cond.b[n0 + 1] : 0.f) + p.t_combined[n0 + 1]));
    });
    pf_prefetch(p.ada.w, p.ada_out, D_FLOW);
    grid.sync();
    // ---- H2: the seven adaLN projections of silu(y) ----
    for (int i = tid; i < R * D_FLOW; i += PF_THREADS) xs[i] = __bfloat162float(p.sy_bf[i]);
    __syncthreads();
    pf_gemv<D_FLOW>(p.ada.w, p.ada_out, xs, R, [&](int r, int n0, float v0, float v1, int nc) {
        float* mp = p.mod + (long long)r * p.ada_out + n0;
        mp[0] = v0 + (p.ada.b ? p.ada.b[n0] : 0.f);
        if (nc == 2) mp[1] = v1 + (p.ada.b ? p.ada.b[n0 + 1] : 0.f);
    });
    pf_prefetch(p.rb[0].m0.w, D_FLOW, D_FLOW);
    grid.sync();
    // ---- H3/H4 x 6: x += gate * mlp2(silu(mlp0(LN(x) (1 + scale) + shift))) (mlp.h:124-140) ----
    for (int b = 0; b < N_RES; b++) {
        const float* m = p.mod + b * 3 * D_FLOW;
        pf_layernorm<D_FLOW>(p.xh, D_FLOW, R, 1e-6f, p.rb[b].lnw, p.rb[b].lnb, m, m + D_FLOW, p.ada_out, xs, red);
        pf_gemv<D_FLOW>(p.rb[b].m0.w, D_FLOW, xs, R, [&](int r, int n0, float v0, float v1, int nc) {
            __nv_bfloat16* hp = p.h1_bf + (long long)r * D_FLOW + n0;
            hp[0] = __float2bfloat16_rn(silu_f(v0 + (p.rb[b].m0.b ? p.rb[b].m0.b[n0] : 0.f)));
            if (nc == 2) hp[1] = __float2bfloat16_rn(silu_f(v1 + (p.rb[b].m0.b ? p.rb[b].m0.b[n0 + 1] : 0.f)));
        });
        pf_prefetch(p.rb[b].m2.w, D_FLOW, D_FLOW);
        grid.sync();
        for (int i = tid; i < R * D_FLOW; i += PF_THREADS) xs[i] = __bfloat162float(p.h1_bf[i]);
        __syncthreads();
        pf_gemv<D_FLOW>(p.rb[b].m2.w, D_FLOW, xs, R, [&](int r, int n0, float v0, float v1, int nc) {
            const float* g = m + 2 * D_FLOW + (long long)r * p.ada_out;
            float* xp = p.xh + (long long)r * D_FLOW + n0;
            xp[0] += (v0 + (p.rb[b].m2.b ? p.rb[b].m2.b[n0] : 0.f)) * g[n0];
            if (nc == 2) xp[1] += (v1 + (p.rb[b].m2.b ? p.rb[b].m2.b[n0 + 1] : 0.f)) * g[n0 + 1];
        });
        if (b + 1 < N_RES) pf_prefetch(p.rb[b + 1].m0.w, D_FLOW, D_FLOW);
        grid.sync();
    }
    // ---- H5: latent = noise + linear(LN(x) (1 + scale) + shift) (mlp.h:156-170, flow_lm.h:141) ----
    {
        const float* m = p.mod + N_RES * 3 * D_FLOW;
        pf_layernorm<D_FLOW>(p.xh, D_FLOW, R, 1e-6f, p.fnw, p.fnb, m, m + D_FLOW, p.ada_out, xs, red);
        pf_gemv<D_FLOW>(p.fin.w, LDIM, xs, R, [&](int r, int n0, float v0, float v1, int nc) {
            float* lp = p.latent + (long long)r * LDIM + n0;
            lp[0] = v0 + (p.fin.b ? p.fin.b[n0] : 0.f) + p.noise_f32[r * LDIM + n0];
            if (nc == 2) lp[1] = v1 + (p.fin.b ? p.fin.b[n0 + 1] : 0.f) + p.noise_f32[r * LDIM + n0 + 1];
        });
    }
    grid.sync();                                               // nobody is still waiting on an earlier target when the base moves
    if (blockIdx.x == 0 && threadIdx.x == 0) p.barrier[1] = grid.target;
}

constexpr size_t PF_SMEM_BYTES = (size_t)(PF_RMAX * D_FF + PF_MAX_KEYS + 16) * 4 + PF_RMAX * 32 * 8 + PF_RMAX * 2 * 4 + PF_RMAX * 64 * 4;

}  // namespace ptts
