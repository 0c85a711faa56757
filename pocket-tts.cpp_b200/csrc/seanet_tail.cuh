// seanet_tail.cuh — the last SEANet residual block and the output conv as one streaming kernel (reference modules/seanet.h:14-27 the
// resnet block, :205-208 its place in the decoder; modules/conv.h:60-131 streaming conv1d): per output sample t of a frame
//     v  = conv_k3(a1)[t]          a1 = f16(elu(y)) with two carried rows in front, 64 -> 32 channels
//     z  = conv_1x1(f16(elu(v)))[t] + y[t]                                          32 -> 64 channels, y = the transposed conv's f32 output (skip)
//     a3 = f16(elu(z))
//     pcm[t] = sum_tap a3[t - 2 + tap] . w11[tap] + b11                             64 -> 1 channel, k = 3
// The unfused path runs three GEMM launches + conv_n1_kernel here and moves 0.7 GB per step for 491 520 rows x 384 B of real input
// (batch 256): a padded 32-channel intermediate, the f16 a3 tensor written and re-read, and per-launch epilogue latency on 64-wide rows.
//
// One WARP owns 16 consecutive samples of one utterance from input to output:
//   * input = the 18-row a1 window (one contiguous 2304-byte run of the channel-last buffer, rows overlap between tiles) and the 16 y rows
//     (4096 B), brought in by cp.async through a private 3-stage ring: no block-level synchronisation anywhere in the steady state;
//   * conv_k3 is a 16 x 192 x 32 GEMM on mma.sync m16n8k16 (f16 -> f32): A fragments are 128-bit loads from the window (k permuted
//     inside 32-wide blocks, same permutation on the weights), weights (12 KB) live in shared memory;
//   * its accumulator fragments, after bias / ELU / f16 rounding, ARE the A fragments of the 1x1 conv (flash-attention style reuse),
//     whose weights stay in registers; bias + skip + ELU + f16 rounding happen on the accumulators;
//   * the k=3 output conv is split by tap: every row contributes d_tap[r] = a3[r] . w11[tap], a 16 x 64 x 3 product that again takes
//     the previous accumulators (bias + skip + ELU, rounded to f16) as its A fragments; pcm[t] = d0[t-2] + d1[t-1] + d2[t] + b is assembled by pcm_combine_kernel, so neither the a3 rows nor a halo between
//     tiles exist. The carried state of the output conv becomes the last two d rows (same zero initial state).
// Rounding points are the unfused path's (f16 a1 / a2 / a3, f32 accumulation, f32 skip); biases are the accumulators' initial values and the
// output conv sums in mma order, so results differ from the unfused launches in the last f32 bits only (measured: 94 dB).
#pragma once
#include "common.cuh"

namespace ptts {

constexpr int ST_THREADS = 192, ST_WARPS = 6, ST_STAGES = 2;   // 111 KB of shared memory: two CTAs per SM, or one beside a FlowLM GEMM CTA
constexpr int ST_ROWS = 16;                                   // samples per work item
constexpr int ST_C = 64, ST_H = 32, ST_K3 = 3 * ST_C;         // channels, hidden channels, conv_k3 reduction length
// shared-memory row pitches: every fragment address is (per-lane base) + (compile-time offset), and the two rows a quarter-warp reads with one
// 128-bit load (or the four rows a half-warp reads with one 64-bit load) fall into different bank groups
constexpr int ST_A1_PITCH = ST_C * 2 + 64;                    // 192 B per a1 row  (rows 16 banks apart)
constexpr int ST_Y_PITCH = ST_C * 4 + 32;                     // 288 B per y row   (rows 8 banks apart)
constexpr int ST_W_PITCH = ST_K3 * 2 + 64;                    // 448 B per conv_k3 weight row
constexpr int ST_A1_BYTES = (ST_ROWS + 2) * ST_A1_PITCH;      // 3456
constexpr int ST_Y_BYTES = ST_ROWS * ST_Y_PITCH;              // 4608
constexpr int ST_STAGE_BYTES = ST_A1_BYTES + ST_Y_BYTES;      // 8064
constexpr int ST_W_BYTES = ST_H * ST_W_PITCH;                 // 14336
constexpr size_t ST_SMEM_BYTES = (size_t)ST_W_BYTES + (size_t)ST_WARPS * ST_STAGES * ST_STAGE_BYTES;
constexpr int ST_DROW = 4;                                    // floats per row of the d buffer (d0, d1, d2, pad)

struct StParams {
    const __half* a1; long long a1_slot_stride;               // [slot][2 + T][64] f16: elu(y), two carried rows in front
    const float* y; long long y_slot_stride;                  // [slot][T][64] f32
    float* d; long long d_slot_stride;                        // [slot][2 + T][4] f32: per-row tap products, two carried rows in front
    const __half* w3; const float* b3;                        // conv_k3 [32][192] (tap-major k), bias [32]
    const __half* w1; int w1_ld; const float* b1;             // 1x1 conv [64][w1_ld] (first 32 k real), bias [64]
    const __half* w11;                                        // output conv [3][64]
    int slot0, n_slots, T;                                    // T % 16 == 0
    int ipw;                                                  // work items per warp; CTA b owns items [b, b + 1) * ST_WARPS * ipw (short-lived CTAs: the
                                                              // high-priority FlowLM stream gets SM slots between them). 0 = persistent grid-stride loop
};

__device__ __forceinline__ void st_cp16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void st_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void st_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t st_pack_elu(float x, float y) {          // f16x2 of (elu(x), elu(y))
    const __half2 h = __floats2half2_rn(elu_f(x), elu_f(y));
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(ST_THREADS, 2) seanet_tail_kernel(const StParams p) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const uint32_t w3s = (uint32_t)__cvta_generic_to_shared(st_smem);
    const uint32_t ring = w3s + ST_W_BYTES + warp * (ST_STAGES * ST_STAGE_BYTES);
    // conv_k3 weights -> shared memory (padded rows)
    for (int i = tid; i < ST_H * (ST_K3 / 8); i += ST_THREADS) {
        const int n = i / (ST_K3 / 8), c = i % (ST_K3 / 8);
        *reinterpret_cast<uint4*>(st_smem + n * ST_W_PITCH + (c << 4)) = __ldg(reinterpret_cast<const uint4*>(p.w3 + n * ST_K3 + c * 8));
    }
    const uint32_t w3l = w3s + g * ST_W_PITCH + t * 16;                         // + 8 j pitch + 64 kb
    const uint32_t a1l = g * ST_A1_PITCH + t * 16, yl = ST_A1_BYTES + g * ST_Y_PITCH + t * 8;   // lane offsets inside a stage
    const uint32_t cpa = (lane >> 3) * ST_A1_PITCH + (lane & 7) * 16, cpy = ST_A1_BYTES + (lane >> 4) * ST_Y_PITCH + (lane & 15) * 16;   // cp.async targets
    // 1x1 conv weights as B fragments in registers: tile j (columns 8 j + g), k-step ks: k = 16 ks + 2 t + {0,1} and + 8
    uint32_t w1r[8][2][2];
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
            const __half* w = p.w1 + (long long)(8 * j + g) * p.w1_ld + 16 * ks + 2 * t;
            w1r[j][ks][0] = __ldg(reinterpret_cast<const uint32_t*>(w)); w1r[j][ks][1] = __ldg(reinterpret_cast<const uint32_t*>(w + 8));
        }
    // output conv as a 64 x 8 B operand (column = tap, 3 real): k-step ks, k = 16 ks + 2 t + {0,1} and + 8; bias of this lane's columns 8 j + 2 t, + 1
    uint32_t w11r[4][2]; float2 b1r[8];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
        w11r[ks][0] = g < 3 ? __ldg(reinterpret_cast<const uint32_t*>(p.w11 + g * ST_C + 16 * ks + 2 * t)) : 0u;
        w11r[ks][1] = g < 3 ? __ldg(reinterpret_cast<const uint32_t*>(p.w11 + g * ST_C + 16 * ks + 2 * t + 8)) : 0u;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) b1r[j] = p.b1 ? make_float2(__ldg(p.b1 + 8 * j + 2 * t), __ldg(p.b1 + 8 * j + 2 * t + 1)) : make_float2(0.f, 0.f);
    float2 b3r[4];
#pragma unroll
    for (int j = 0; j < 4; j++) b3r[j] = p.b3 ? make_float2(__ldg(p.b3 + 8 * j + 2 * t), __ldg(p.b3 + 8 * j + 2 * t + 1)) : make_float2(0.f, 0.f);
    __syncthreads();

    const int tiles_per_slot = p.T / ST_ROWS;
    const int all_items = p.n_slots * tiles_per_slot;         // < 2^31: 120 items per slot
    const int chunk = ST_WARPS * p.ipw;
    const int n_items = p.ipw > 0 ? min(all_items, (int)(blockIdx.x + 1) * chunk) : all_items;
    const int gw = (p.ipw > 0 ? blockIdx.x * chunk : blockIdx.x * ST_WARPS) + warp, nw = p.ipw > 0 ? ST_WARPS : gridDim.x * ST_WARPS;
    auto issue = [&](int item, int stage) {             // the item's a1 window and y rows -> ring stage (swizzled 16-byte chunks)
        if (item < n_items) {
            const int s = item / tiles_per_slot, t0 = (item - s * tiles_per_slot) * ST_ROWS;
            const char* a = reinterpret_cast<const char*>(p.a1 + (long long)(p.slot0 + s) * p.a1_slot_stride + (long long)t0 * ST_C);
            const char* yv = reinterpret_cast<const char*>(p.y + (long long)(p.slot0 + s) * p.y_slot_stride + (long long)t0 * ST_C);
            const uint32_t base = ring + stage * ST_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 5; i++)                         // window chunk lane + 32 i (8 per a1 row, 18 rows)
                if (i < 4 || lane < 16) st_cp16(base + cpa + i * 4 * ST_A1_PITCH, a + (lane + 32 * i) * 16);
#pragma unroll
            for (int i = 0; i < 8; i++)                         // y chunk lane + 32 i (16 per row)
                st_cp16(base + cpy + i * 2 * ST_Y_PITCH, yv + (lane + 32 * i) * 16);
        }
        st_commit();
    };
#pragma unroll
    for (int s = 0; s < ST_STAGES - 1; s++) issue(gw + s * nw, s);

    int stage = 0;
    for (int item = gw; item < n_items; item += nw) {
        issue(item + (ST_STAGES - 1) * nw, (stage + ST_STAGES - 1) % ST_STAGES);
        st_wait<ST_STAGES - 1>();
        __syncwarp();
        const uint32_t sb = ring + stage * ST_STAGE_BYTES, a1s = sb + a1l, ys = sb + yl;
        // ---- v = conv_k3(a1): 16 x 192 x 32 ----
        float acc3[4][4];
#pragma unroll
        for (int j = 0; j < 4; j++) { acc3[j][0] = acc3[j][2] = b3r[j].x; acc3[j][1] = acc3[j][3] = b3r[j].y; }   // bias = initial accumulator
#pragma unroll
        for (int kb = 0; kb < ST_K3 / 32; kb++) {
            // window element 64 r + 32 kb + 8 t of tile row r = a1 row r + (kb >> 1), chunk 4 (kb & 1) + t
            uint4 al, ah;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(al.x), "=r"(al.y), "=r"(al.z), "=r"(al.w) : "r"(a1s + (kb >> 1) * ST_A1_PITCH + (kb & 1) * 64));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(ah.x), "=r"(ah.y), "=r"(ah.z), "=r"(ah.w) : "r"(a1s + (8 + (kb >> 1)) * ST_A1_PITCH + (kb & 1) * 64));
            const uint32_t a0[4] = {al.x, ah.x, al.y, ah.y}, a1[4] = {al.z, ah.z, al.w, ah.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint4 w;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(w3l + 8 * j * ST_W_PITCH + 64 * kb));
                st_mma(acc3[j], a0, w.x, w.y);
                st_mma(acc3[j], a1, w.z, w.w);
            }
        }
        // ---- z = conv_1x1(f16(elu(v + b3))) : the accumulator fragments of column tiles 2 ks, 2 ks + 1 are the A fragment of k-step ks ----
        uint32_t af[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
            af[ks][0] = st_pack_elu(acc3[2 * ks][0], acc3[2 * ks][1]);
            af[ks][1] = st_pack_elu(acc3[2 * ks][2], acc3[2 * ks][3]);
            af[ks][2] = st_pack_elu(acc3[2 * ks + 1][0], acc3[2 * ks + 1][1]);
            af[ks][3] = st_pack_elu(acc3[2 * ks + 1][2], acc3[2 * ks + 1][3]);
        }
        // a3 = f16(elu(z)) fragments of column tiles 2 ks, 2 ks + 1 are in turn the A fragment of k-step ks of the output conv's tap products
        uint32_t a3f[4][4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float acc1[4] = {b1r[j].x, b1r[j].y, b1r[j].x, b1r[j].y};
            st_mma(acc1, af[0], w1r[j][0][0], w1r[j][0][1]);
            st_mma(acc1, af[1], w1r[j][1][0], w1r[j][1][1]);
            float2 ylo, yhi;                                             // skip rows g, g + 8, columns 8 j + 2 t, + 1
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(ylo.x), "=f"(ylo.y) : "r"(ys + 32 * j));
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(yhi.x), "=f"(yhi.y) : "r"(ys + 8 * ST_Y_PITCH + 32 * j));
            a3f[j >> 1][(j & 1) * 2] = st_pack_elu(__fadd_rn(acc1[0], ylo.x), __fadd_rn(acc1[1], ylo.y));
            a3f[j >> 1][(j & 1) * 2 + 1] = st_pack_elu(__fadd_rn(acc1[2], yhi.x), __fadd_rn(acc1[3], yhi.y));
        }
        float dacc[4] = {0.f, 0.f, 0.f, 0.f};                           // columns = taps: (row g: d0, d1 | row g + 8: d0, d1) in lane t = 0, d2 in lane t = 1
#pragma unroll
        for (int ks = 0; ks < 4; ks++) st_mma(dacc, a3f[ks], w11r[ks][0], w11r[ks][1]);
        if (t < 2) {
            const int s = item / tiles_per_slot, t0 = (item - s * tiles_per_slot) * ST_ROWS;
            float* dp = p.d + (long long)(p.slot0 + s) * p.d_slot_stride + (long long)(2 + t0 + g) * ST_DROW;
            if (t == 0) { *reinterpret_cast<float2*>(dp) = make_float2(dacc[0], dacc[1]); *reinterpret_cast<float2*>(dp + 8 * ST_DROW) = make_float2(dacc[2], dacc[3]); }
            else { dp[2] = dacc[0]; dp[8 * ST_DROW + 2] = dacc[2]; }
        }
        __syncwarp();                                            // the stage is refilled by the next iteration's issue
        stage = (stage + 1) % ST_STAGES;
    }
    st_wait<0>();
}

// pcm[t] = d0[t - 2] + d1[t - 1] + d2[t] + bias over the d rows of seanet_tail_kernel (rows 0, 1 of a slot = carried from the previous frame)
__global__ void __launch_bounds__(256) pcm_combine_kernel(const float* __restrict__ d, long long d_slot_stride, int slot0, int n_slots, int T,
                                                          const float* __restrict__ bias, float* __restrict__ pcm) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_slots * T) return;
    const int s = (int)(i / T), t = (int)(i % T);
    const float* r = d + (long long)(slot0 + s) * d_slot_stride + (long long)t * ST_DROW;
    pcm[(long long)(slot0 + s) * T + t] = (r[0] + r[ST_DROW + 1]) + r[2 * ST_DROW + 2] + (bias ? bias[0] : 0.f);
}

}  // namespace ptts
