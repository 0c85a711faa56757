// seanet_res.cuh — SEANet resnet block 6 (128 channels, 64 hidden; reference modules/seanet.h:14-27, :201-203 its place in the decoder) as one
// streaming kernel, the 128-channel sibling of seanet_tail.cuh: per sample t of the 480-row stage
//     v = conv_k3(a1)[t]                         a1 = f16(elu(y)) with two carried rows in front, 128 -> 64 channels
//     z = conv_1x1(f16(elu(v)))[t] + y[t]        64 -> 128 channels, y = the transposed conv's f32 output (skip)
//     out[t] = f16(elu(z))                       the next transposed conv's input row
// instead of two GEMM launches with a 64-channel f16 intermediate in HBM. One WARP owns 16 consecutive samples: the 18-row a1 window through a
// 2-stage cp.async ring, the 16 y rows through a single buffer filled while conv_k3 runs, conv_k3 as a 16 x 384 x 64 mma.sync GEMM (weights in
// padded shared memory, k permuted inside 32-wide blocks on both operands), its accumulators (bias, ELU, f16) reused as the A fragments of
// the 1x1 conv, whose weights sit in shared memory in fragment order (two 128-bit loads give a lane its B fragments of all four k-steps of a
// column tile), bias + skip + ELU + f16 on the accumulators, the 16 x 128 result staged per warp and written with 16-byte stores.
// Rounding points are the unfused path's; biases are the accumulators' initial values (differences in the last f32 bits only).
// MEASURED: not a win (kept as an opt-in experiment, PTTS_B200_FUSED_RES6=1): Mimi decode 0.484 -> 0.494 ms at batch 256. The stage's
// intermediates fit in L2, so the two tcgen05 GEMMs it replaces were never HBM bound.
#pragma once
#include "common.cuh"
#include "seanet_tail.cuh"

namespace ptts {

constexpr int SR_THREADS = 192, SR_WARPS = 6;
constexpr int SR_ROWS = 16, SR_C = 128, SR_H = 64, SR_K3 = 3 * SR_C;
constexpr int SR_A_PITCH = SR_C * 2 + 64;                     // 320 B per a1 row (rows 16 banks apart)
constexpr int SR_Y_PITCH = SR_C * 4 + 32;                     // 544 B per y row (rows 8 banks apart)
constexpr int SR_O_PITCH = SR_C * 2 + 16;                     // 272 B per staged output row (rows 4 banks apart)
constexpr int SR_W3_PITCH = SR_K3 * 2 + 64;                   // 832 B per conv_k3 weight row
constexpr int SR_W1_PITCH = SR_H * 2 + 16;                    // 144 B per 1x1 weight row (fragment order)
constexpr int SR_A_BYTES = (SR_ROWS + 2) * SR_A_PITCH;        // 5760
constexpr int SR_Y_BYTES = SR_ROWS * SR_Y_PITCH;              // 8704
constexpr int SR_O_BYTES = SR_ROWS * SR_O_PITCH;              // 4352
constexpr int SR_WARP_BYTES = 2 * SR_A_BYTES + SR_Y_BYTES + SR_O_BYTES;   // 24576
constexpr int SR_W3_BYTES = SR_H * SR_W3_PITCH;               // 53248
constexpr int SR_W1_BYTES = SR_C * SR_W1_PITCH;               // 18432
constexpr size_t SR_SMEM_BYTES = (size_t)SR_W3_BYTES + SR_W1_BYTES + SR_C * 4 + (size_t)SR_WARPS * SR_WARP_BYTES;

struct SrParams {
    const __half* a1; long long a1_slot_stride;               // [slot][2 + T][128] f16
    const float* y; long long y_slot_stride;                  // [slot][T][128] f32
    __half* out; long long out_slot_stride; int out_row0;     // [slot][out_row0 + T][128] f16
    const __half* w3; const float* b3;                        // conv_k3 [64][384] (tap-major k), bias [64]
    const __half* w1; int w1_ld; const float* b1;             // 1x1 conv [128][w1_ld >= 64], bias [128]
    int slot0, n_slots, T;                                    // T % 16 == 0
};

__global__ void __launch_bounds__(SR_THREADS, 1) seanet_res_kernel(const SrParams p) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char sr_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const uint32_t w3s = (uint32_t)__cvta_generic_to_shared(sr_smem);
    const uint32_t w1s = w3s + SR_W3_BYTES, b1s = w1s + SR_W1_BYTES;
    const uint32_t wbase = b1s + SR_C * 4 + warp * SR_WARP_BYTES;       // this warp's [a1 stage 0 | a1 stage 1 | y | out]
    for (int i = tid; i < SR_H * (SR_K3 / 8); i += SR_THREADS) {        // conv_k3 weights, padded rows
        const int n = i / (SR_K3 / 8), c = i % (SR_K3 / 8);
        *reinterpret_cast<uint4*>(sr_smem + n * SR_W3_PITCH + (c << 4)) = __ldg(reinterpret_cast<const uint4*>(p.w3 + n * SR_K3 + c * 8));
    }
    for (int i = tid; i < SR_C * 16; i += SR_THREADS) {                 // 1x1 weights in fragment order: row n, lane group tt, k-step ks -> k = 16 ks + 2 tt + {0, 1, 8, 9}
        const int n = i >> 4, tt = (i >> 2) & 3, ks = i & 3;
        const __half* w = p.w1 + (long long)n * p.w1_ld + 16 * ks + 2 * tt;
        uint2 v; v.x = __ldg(reinterpret_cast<const uint32_t*>(w)); v.y = __ldg(reinterpret_cast<const uint32_t*>(w + 8));
        *reinterpret_cast<uint2*>(sr_smem + SR_W3_BYTES + n * SR_W1_PITCH + tt * 32 + ks * 8) = v;
    }
    for (int i = tid; i < SR_C; i += SR_THREADS) reinterpret_cast<float*>(sr_smem + SR_W3_BYTES + SR_W1_BYTES)[i] = p.b1 ? __ldg(p.b1 + i) : 0.f;
    float2 b3r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) b3r[j] = p.b3 ? make_float2(__ldg(p.b3 + 8 * j + 2 * t), __ldg(p.b3 + 8 * j + 2 * t + 1)) : make_float2(0.f, 0.f);
    __syncthreads();

    const uint32_t w3l = w3s + g * SR_W3_PITCH + t * 16, w1l = w1s + g * SR_W1_PITCH + t * 32, b1l = b1s + t * 8;
    const uint32_t al = g * SR_A_PITCH + t * 16;
    const uint32_t ysb = wbase + 2 * SR_A_BYTES, osb = ysb + SR_Y_BYTES;
    const uint32_t yl = ysb + g * SR_Y_PITCH + t * 8, ol = osb + g * SR_O_PITCH + t * 4;
    const int tiles_per_slot = p.T / SR_ROWS;
    const int n_items = p.n_slots * tiles_per_slot;
    const int gw = blockIdx.x * SR_WARPS + warp, nw = gridDim.x * SR_WARPS;
    auto issue_a = [&](int item, int stage) {                  // 18 a1 rows x 256 B (16 chunks per row) -> ring stage
        if (item < n_items) {
            const int s = item / tiles_per_slot, t0 = (item - s * tiles_per_slot) * SR_ROWS;
            const char* a = reinterpret_cast<const char*>(p.a1 + (long long)(p.slot0 + s) * p.a1_slot_stride + (long long)t0 * SR_C);
            const uint32_t base = wbase + stage * SR_A_BYTES + (lane >> 4) * SR_A_PITCH + (lane & 15) * 16;
#pragma unroll
            for (int i = 0; i < 9; i++) st_cp16(base + i * 2 * SR_A_PITCH, a + (lane + 32 * i) * 16);
        }
        st_commit();
    };
    auto issue_y = [&](int item) {                             // 16 y rows x 512 B (32 chunks per row)
        const int s = item / tiles_per_slot, t0 = (item - s * tiles_per_slot) * SR_ROWS;
        const char* yv = reinterpret_cast<const char*>(p.y + (long long)(p.slot0 + s) * p.y_slot_stride + (long long)t0 * SR_C);
#pragma unroll
        for (int i = 0; i < 16; i++) st_cp16(ysb + i * SR_Y_PITCH + lane * 16, yv + (lane + 32 * i) * 16);
        st_commit();
    };
    issue_a(gw, 0);
    int stage = 0;
    for (int item = gw; item < n_items; item += nw) {
        issue_y(item);
        issue_a(item + nw, stage ^ 1);
        st_wait<2>();                                          // this item's a1 window has landed (its y rows and the next window may be in flight)
        __syncwarp();
        const uint32_t a1s = wbase + stage * SR_A_BYTES + al;
        // ---- v = conv_k3(a1): 16 x 384 x 64, bias = initial accumulator ----
        float acc3[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) { acc3[j][0] = acc3[j][2] = b3r[j].x; acc3[j][1] = acc3[j][3] = b3r[j].y; }
#pragma unroll
        for (int kb = 0; kb < SR_K3 / 32; kb++) {
            // window element 128 r + 32 kb + 8 t of tile row r = a1 row r + (kb >> 2), chunk 4 (kb & 3) + t
            uint4 a_lo, a_hi;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a_lo.x), "=r"(a_lo.y), "=r"(a_lo.z), "=r"(a_lo.w) : "r"(a1s + (kb >> 2) * SR_A_PITCH + (kb & 3) * 64));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a_hi.x), "=r"(a_hi.y), "=r"(a_hi.z), "=r"(a_hi.w) : "r"(a1s + (8 + (kb >> 2)) * SR_A_PITCH + (kb & 3) * 64));
            const uint32_t a0[4] = {a_lo.x, a_hi.x, a_lo.y, a_hi.y}, a1[4] = {a_lo.z, a_hi.z, a_lo.w, a_hi.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint4 w;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(w3l + 8 * j * SR_W3_PITCH + 64 * kb));
                st_mma(acc3[j], a0, w.x, w.y);
                st_mma(acc3[j], a1, w.z, w.w);
            }
        }
        // ---- f16(elu(v)) fragments of column tiles 2 ks, 2 ks + 1 = A fragment of k-step ks of the 1x1 conv ----
        uint32_t af[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            af[ks][0] = st_pack_elu(acc3[2 * ks][0], acc3[2 * ks][1]);
            af[ks][1] = st_pack_elu(acc3[2 * ks][2], acc3[2 * ks][3]);
            af[ks][2] = st_pack_elu(acc3[2 * ks + 1][0], acc3[2 * ks + 1][1]);
            af[ks][3] = st_pack_elu(acc3[2 * ks + 1][2], acc3[2 * ks + 1][3]);
        }
        st_wait<1>();                                          // the y rows (only the next a1 window may still be in flight)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; j++) {
            float2 bj;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(bj.x), "=f"(bj.y) : "r"(b1l + 32 * j));
            float acc1[4] = {bj.x, bj.y, bj.x, bj.y};
            uint4 wa, wb;                                      // B fragments of k-steps 0, 1 | 2, 3 of column tile j
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(wa.x), "=r"(wa.y), "=r"(wa.z), "=r"(wa.w) : "r"(w1l + 8 * j * SR_W1_PITCH));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(wb.x), "=r"(wb.y), "=r"(wb.z), "=r"(wb.w) : "r"(w1l + 8 * j * SR_W1_PITCH + 16));
            st_mma(acc1, af[0], wa.x, wa.y); st_mma(acc1, af[1], wa.z, wa.w);
            st_mma(acc1, af[2], wb.x, wb.y); st_mma(acc1, af[3], wb.z, wb.w);
            float2 ylo, yhi;                                   // skip rows g, g + 8, columns 8 j + 2 t, + 1
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(ylo.x), "=f"(ylo.y) : "r"(yl + 32 * j));
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(yhi.x), "=f"(yhi.y) : "r"(yl + 8 * SR_Y_PITCH + 32 * j));
            const uint32_t o_lo = st_pack_elu(__fadd_rn(acc1[0], ylo.x), __fadd_rn(acc1[1], ylo.y));
            const uint32_t o_hi = st_pack_elu(__fadd_rn(acc1[2], yhi.x), __fadd_rn(acc1[3], yhi.y));
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ol + 16 * j), "r"(o_lo) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ol + 8 * SR_O_PITCH + 16 * j), "r"(o_hi) : "memory");
        }
        __syncwarp();
        {   // 16 rows x 256 B -> global, 16-byte stores (16 chunks per row)
            const int s = item / tiles_per_slot, t0 = (item - s * tiles_per_slot) * SR_ROWS;
            char* dst = reinterpret_cast<char*>(p.out + (long long)(p.slot0 + s) * p.out_slot_stride + (long long)(p.out_row0 + t0) * SR_C);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int q = lane + 32 * i, r = q >> 4, c = q & 15;
                uint4 v;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(osb + r * SR_O_PITCH + c * 16));
                *reinterpret_cast<uint4*>(dst + r * (SR_C * 2) + c * 16) = v;
            }
        }
        __syncwarp();                                          // y / out / this a1 stage are refilled by the next iterations
        stage ^= 1;
    }
    st_wait<0>();
}

}  // namespace ptts
