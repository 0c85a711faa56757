// =====================================================================================
// ptts_oracle.cpp — CPU ORACLE for the Pocket-TTS per-frame generation path.
//
// TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked, imported or executed by the
// product (pocket-tts.cpp_b200/). Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may call it, and only as the checker / CPU baseline.
//
// PARITY UNPINNED: the reference (Codes4Fun/pocket-tts.cpp) ships no tests, golden vectors
// or fixtures for this path, and it cannot be compiled here (ggml, SentencePiece C++, SDL2,
// FFmpeg are absent and unpinned: cmake/FindGGML.cmake:11-34). This file is a restatement of
// the reference's ggml *graph* (file:line cited per function) with ggml's CPU-backend
// arithmetic (un-vendored, unpinned github.com/ggml-org/ggml, late-2025 API level) restated
// from its published algorithm: mul_mat rounds the activation operand to the weight's
// vec_dot_type (BF16 weights -> bf16 activations, F16 -> f16), accumulates in f32;
// ggml_norm uses double sums and a biased variance; ggml_gelu goes through an f16 lookup
// table; ggml_conv_1d = im2col(F16) + f16 mul_mat; ggml_conv_transpose_1d is f32.
// It is cross-checked by an independent PyTorch restatement (tests/torch_second_opinion.py).
//
// Layout: activations are row-major [T][C] (time-major, channel-last) == ggml [C, T].
// Batch is an outer loop (one oracle_stream per utterance), as in the reference (batch 1).
// =====================================================================================
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------
// rounding helpers (ggml: GGML_FP32_TO_BF16 round-nearest-even, GGML_FP32_TO_FP16 via F16C)
// ---------------------------------------------------------------------------------
inline float bf16r(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    if ((u & 0x7fffffff) > 0x7f800000) { u |= 0x00400000; u &= 0xffff0000; }   // quiet NaN
    else { u += 0x7fff + ((u >> 16) & 1); u &= 0xffff0000; }
    float y; memcpy(&y, &u, 4); return y;
}
inline float f16r(float x) { return (float)(_Float16)x; }

enum ActRound { ACT_F32 = 0, ACT_BF16 = 1 };

struct Lin {             // torch Linear [out,in]; weights hold file-dtype-exact values (src/loader.h:205)
    const float* w = nullptr; const float* b = nullptr; int out = 0, in = 0;
};

int g_threads = 1;

// y[T][out] = x[T][in] . W^T (+b). ggml_mul_mat semantics (src/torch.h:79-87): activations are
// rounded to the weight type's vec_dot_type first, products accumulated in f32.
void linear(const Lin& L, const float* x, int T, float* y, ActRound ar) {
    std::vector<float> xr;
    const float* xs = x;
    if (ar == ACT_BF16) {
        xr.resize((size_t)T * L.in);
        for (size_t i = 0; i < xr.size(); i++) xr[i] = bf16r(x[i]);
        xs = xr.data();
    }
    const int in = L.in, out = L.out;
#pragma omp parallel for schedule(static) num_threads(g_threads) if ((long)out * in * T > 65536)
    for (int o = 0; o < out; o++) {
        const float* w = L.w + (size_t)o * in;
        for (int t = 0; t < T; t++) {
            const float* xv = xs + (size_t)t * in;
            float acc = 0.f;
#pragma omp simd reduction(+ : acc)
            for (int i = 0; i < in; i++) acc += w[i] * xv[i];
            y[(size_t)t * out + o] = acc + (L.b ? L.b[o] : 0.f);
        }
    }
}

// ggml_norm + mul + add (src/torch.h:49-60; src/pocket_tts/modules/mlp.h:52-64): double sums, biased variance.
void layer_norm(const float* x, int T, int C, float eps, const float* w, const float* b, float* y) {
    for (int t = 0; t < T; t++) {
        const float* xv = x + (size_t)t * C; float* yv = y + (size_t)t * C;
        double sum = 0; for (int i = 0; i < C; i++) sum += (double)xv[i];
        float mean = (float)(sum / C);
        double sum2 = 0;
        for (int i = 0; i < C; i++) { float v = xv[i] - mean; yv[i] = v; sum2 += (double)(v * v); }
        float var = (float)(sum2 / C);
        float scale = 1.0f / sqrtf(var + eps);
        for (int i = 0; i < C; i++) {
            float v = yv[i] * scale;
            if (w) v = v * w[i];
            if (b) v = v + b[i];
            yv[i] = v;
        }
    }
}

// ggml_gelu on the CPU backend: tanh form evaluated through an f16 table (input and output rounded to f16).
inline float gelu_ggml(float x) {
    if (x <= -10.0f) return 0.0f;
    if (x >= 10.0f) return x;
    float xf = f16r(x);
    const float GELU_COEF_A = 0.044715f, SQRT_2_OVER_PI = 0.79788456080286535587989211986876f;
    float g = 0.5f * xf * (1.0f + tanhf(SQRT_2_OVER_PI * xf * (1.0f + GELU_COEF_A * xf * xf)));
    return f16r(g);
}
inline float silu(float x) { return x / (1.0f + expf(-x)); }
inline float elu(float x) { return x > 0.f ? x : expm1f(x); }

// ggml_soft_max_ext(a, mask, scale, 0): softmax(a*scale + mask) with a double sum.
void softmax_row(float* s, int n, float scale, const float* mask) {
    float mx = -INFINITY;
    for (int i = 0; i < n; i++) { s[i] = s[i] * scale + (mask ? mask[i] : 0.f); mx = std::max(mx, s[i]); }
    double sum = 0;
    for (int i = 0; i < n; i++) { float e = expf(s[i] - mx); s[i] = e; sum += (double)e; }
    float inv = (float)(1.0 / sum);
    for (int i = 0; i < n; i++) s[i] *= inv;
}

// ---------------------------------------------------------------------------------
// model containers
// ---------------------------------------------------------------------------------
struct Named { std::vector<float> data; std::vector<int64_t> shape; };

struct FlowLayer { Lin in_proj, out_proj, lin1, lin2; const float *n1w, *n1b, *n2w, *n2b; };
struct ResBlock { const float *lnw, *lnb; Lin mlp0, mlp2, ada; };
struct MimiLayer { Lin in_proj, out_proj, lin1, lin2; const float *n1w, *n1b, *n2w, *n2b, *ls1, *ls2; };
struct Conv { int cin, cout, k; std::vector<float> w; const float* b; };        // w repacked [cout][k][cin], f16-exact
struct ConvT { int cin, cout, k, s; std::vector<float> w; const float* b; };    // w repacked [k][cout][cin], f32

}  // namespace

struct oracle_ctx {
    int file_bf16 = 1;
    ActRound ar = ACT_BF16;
    std::map<std::string, Named> T;
    bool finalized = false;
    // FlowLM
    const float* embed = nullptr; int n_embed = 0;
    const float *emb_std, *emb_mean, *bos_emb;
    Lin input_linear, out_eos; const float *onw, *onb;
    FlowLayer fl[6];
    // head
    Lin input_proj, cond_embed, final_lin, final_ada; ResBlock rb[6];
    const float *fnw = nullptr, *fnb = nullptr;
    std::vector<float> t_combined;   // (TE1(1)+TE0(0))/2, constant (src/pocket_tts/models/flow_lm.h:137-138)
    // Mimi
    std::vector<float> quant_w;      // [512][32] f16-exact
    std::vector<float> up_w;         // [512][32] f32   (torch [512,1,32])
    const float* up_b = nullptr;
    MimiLayer ml[2];
    Conv c0, r3a, r3b, r6a, r6b, r9a, r9b, c11;
    ConvT t2, t5, t8;
    std::vector<float> pattern;      // Mimi bias pattern [16][734] (src/torch.h:168-201)
    float rope_freq_flow[32], rope_freq_mimi[32];
    bool taps = false;
};

struct oracle_stream {
    oracle_ctx* ctx;
    int cap;                                   // FlowLM KV capacity (reference: 1000, src/pocket_tts.cpp:367)
    // FlowLM KV, fp32 [layer][pos][1024] (src/pocket_tts/modules/transformer.h:21-33)
    std::vector<float> K[6], V[6], cK[6], cV[6];
    int current_end = 0, cond_end = 0;
    // Mimi state
    std::vector<float> up_state;               // [32][512]
    std::vector<float> mK[2], mV[2];           // bf16-exact [250][512]
    int mimi_offset = 0;
    std::vector<float> s0, s3, s6, s9, s11;    // causal conv states [(K-1)][Cin]
    std::vector<float> st2, st5, st8;          // convT states: full previous pre-bias output [(T-1)s+K][Cout]
    // sentence state (src/pocket_tts.cpp:333-349)
    int frames_after_eos = 0, max_gen_len = 0, generation_step = 0, eos_step = -1;
    float backbone_input[32];
    std::map<std::string, std::vector<float>> tap;
};

namespace {

const float* need(oracle_ctx* c, const std::string& name, std::vector<int64_t> shape, bool optional = false) {
    auto it = c->T.find(name);
    if (it == c->T.end()) {
        if (optional) return nullptr;
        fprintf(stderr, "oracle: missing tensor %s\n", name.c_str()); exit(1);
    }
    if (!shape.empty() && it->second.shape != shape) {
        fprintf(stderr, "oracle: bad shape for %s\n", name.c_str()); exit(1);
    }
    return it->second.data.data();
}

Lin get_lin(oracle_ctx* c, const std::string& p, int out, int in) {
    Lin L; L.out = out; L.in = in;
    L.w = need(c, p + ".weight", {out, in});
    L.b = need(c, p + ".bias", {out}, true);
    return L;
}

Conv get_conv(oracle_ctx* c, const std::string& p, int cout, int cin, int k) {
    // ggml_conv_1d weights are converted to F16 at load (src/loader.h:209)
    const float* w = need(c, p + ".conv.weight", {cout, cin, k});
    Conv cv; cv.cin = cin; cv.cout = cout; cv.k = k; cv.w.resize((size_t)cout * k * cin);
    for (int co = 0; co < cout; co++) for (int ci = 0; ci < cin; ci++) for (int kk = 0; kk < k; kk++)
        cv.w[((size_t)co * k + kk) * cin + ci] = f16r(w[((size_t)co * cin + ci) * k + kk]);
    cv.b = need(c, p + ".conv.bias", {cout}, true);
    return cv;
}

ConvT get_convt(oracle_ctx* c, const std::string& p, int cin, int cout, int k, int s) {
    // ggml_conv_transpose_1d weights are F32 at load (src/loader.h:210)
    const float* w = need(c, p + ".convtr.weight", {cin, cout, k});
    ConvT cv; cv.cin = cin; cv.cout = cout; cv.k = k; cv.s = s; cv.w.resize((size_t)k * cout * cin);
    for (int ci = 0; ci < cin; ci++) for (int co = 0; co < cout; co++) for (int kk = 0; kk < k; kk++)
        cv.w[((size_t)kk * cout + co) * cin + ci] = w[((size_t)ci * cout + co) * k + kk];
    cv.b = need(c, p + ".convtr.bias", {cout}, true);
    return cv;
}

// src/pocket_tts/modules/mlp.h:92-106 + :18-37 (variance-"RMSNorm": unbiased var, numerator NOT mean-subtracted)
void timestep_embed(oracle_ctx* c, int idx, float t, float* out512) {
    std::string p = "flow_lm.flow_net.time_embed." + std::to_string(idx);
    Lin m0 = get_lin(c, p + ".mlp.0", 512, 256), m2 = get_lin(c, p + ".mlp.2", 512, 512);
    const float* alpha = need(c, p + ".mlp.3.alpha", {512});
    const float* freqs = need(c, p + ".freqs", {128});
    float e[256];
    for (int i = 0; i < 128; i++) { float a = freqs[i] * t; e[i] = cosf(a); e[128 + i] = sinf(a); }
    float h[512], u[512];
    linear(m0, e, 1, h, c->ar);
    for (int i = 0; i < 512; i++) h[i] = silu(h[i]);
    linear(m2, h, 1, u, c->ar);
    // ggml_mean (f32 row sum / n), then sum of squared deviations / (n-1)
    double s = 0; for (int i = 0; i < 512; i++) s += (double)u[i];
    float mean = (float)(s / 512.0);
    double ss = 0; for (int i = 0; i < 512; i++) { float d = u[i] - mean; ss += (double)(d * d); }
    float var = (float)ss * (1.f / 511.f) + 1e-5f;
    float sd = sqrtf(var);
    for (int i = 0; i < 512; i++) out512[i] = alpha[i] * (u[i] / sd);
}

// src/torch.h:168-201 create_bias_pattern(capacity=250, t=16, hi=0, lo=-inf)
void make_pattern(oracle_ctx* c) {
    const int capacity = 250, t = 16;
    const int start = capacity * 2 - t, width = start + capacity;
    c->pattern.assign((size_t)t * width, 0.f);
    for (int j = 0; j < t; j++) {
        int toff = j * width;
        int right = start + 1 + j;
        for (int i = 0; i < right; i++) c->pattern[toff + i] = 0.f;
        for (int i = right; i < width; i++) c->pattern[toff + i] = -INFINITY;
        int b = t - j - 1;
        toff += capacity - 1;
        for (int i = 0; i < b; i++) c->pattern[toff - i] = -INFINITY;
    }
}
// src/torch.h:203-221 bias_pattern_index: returns pointer to row j of the [250] x 16 view
const float* pattern_row(oracle_ctx* c, int offset, int j) {
    const int capacity = 250, t = 16, start = capacity * 2 - t, width = start + capacity;
    int off = (offset <= capacity) ? start - offset : capacity - (offset % capacity);
    return c->pattern.data() + (size_t)j * width + off;
}

// ---------------------------------------------------------------------------------
// FlowLM transformer over T rows (src/pocket_tts/modules/transformer.h:55-199,253-278,363-374)
// ---------------------------------------------------------------------------------
void flowlm_forward(oracle_stream* s, std::vector<float>& x, int T) {
    oracle_ctx* c = s->ctx;
    const int D = 1024, H = 16, Dh = 64;
    const int p0 = s->current_end;
    if (p0 + T > s->cap) { fprintf(stderr, "oracle: FlowLM KV overflow (%d+%d>%d)\n", p0, T, s->cap); exit(1); }
    std::vector<float> n((size_t)T * D), qkv((size_t)T * 3 * D), q((size_t)T * D), att((size_t)T * D), upd((size_t)T * D),
        hid((size_t)T * 4 * D);
    for (int l = 0; l < 6; l++) {
        const FlowLayer& L = c->fl[l];
        layer_norm(x.data(), T, D, 1e-5f, L.n1w, L.n1b, n.data());
        linear(L.in_proj, n.data(), T, qkv.data(), c->ar);
        // RoPE (src/pocket_tts/modules/rope.h:22-43,183-272): pairs (2i,2i+1) rotated, output de-interleaved [re | im]
        float* Kc = s->K[l].data(); float* Vc = s->V[l].data();
        for (int t = 0; t < T; t++) {
            const float ts = (float)t + (float)p0;                       // arange + offset, f32
            for (int h = 0; h < H; h++) {
                const float* qs = &qkv[(size_t)t * 3 * D + h * Dh];
                const float* ks = qs + D; const float* vs = qs + 2 * D;
                float* qd = &q[(size_t)t * D + h * Dh];
                float* kd = &Kc[(size_t)(p0 + t) * D + h * Dh];
                float* vd = &Vc[(size_t)(p0 + t) * D + h * Dh];
                for (int i = 0; i < 32; i++) {
                    float rad = ts * c->rope_freq_flow[i];
                    float cr = cosf(rad), sr = sinf(rad);
                    float a = qs[2 * i], b = qs[2 * i + 1];
                    qd[i] = a * cr - b * sr; qd[32 + i] = a * sr + b * cr;
                    a = ks[2 * i]; b = ks[2 * i + 1];
                    kd[i] = a * cr - b * sr; kd[32 + i] = a * sr + b * cr;
                }
                memcpy(vd, vs, Dh * sizeof(float));
            }
        }
        // SDPA (src/torch.h:128-150): scale 1/8; mask only when T != 1: key x visible iff y + shift >= x
        const int Lk = p0 + T;
#pragma omp parallel for schedule(static) num_threads(g_threads) collapse(2) if ((long)T * Lk > 2048)
        for (int t = 0; t < T; t++) for (int h = 0; h < H; h++) {
            std::vector<float> sc(Lk), mask;
            const float* qv = &q[(size_t)t * D + h * Dh];
            for (int j = 0; j < Lk; j++) {
                const float* kv = &Kc[(size_t)j * D + h * Dh];
                float acc = 0.f;
                for (int d = 0; d < Dh; d++) acc += kv[d] * qv[d];
                sc[j] = acc;
            }
            if (T != 1) { mask.resize(Lk); for (int j = 0; j < Lk; j++) mask[j] = (t + p0 >= j) ? 0.f : -INFINITY; }
            softmax_row(sc.data(), Lk, 0.125f, T != 1 ? mask.data() : nullptr);
            float* o = &att[(size_t)t * D + h * Dh];
            for (int d = 0; d < Dh; d++) {
                float acc = 0.f;
                for (int j = 0; j < Lk; j++) acc += Vc[(size_t)j * D + h * Dh + d] * sc[j];
                o[d] = acc;
            }
        }
        linear(L.out_proj, att.data(), T, upd.data(), c->ar);
        for (size_t i = 0; i < (size_t)T * D; i++) x[i] = x[i] + upd[i];
        layer_norm(x.data(), T, D, 1e-5f, L.n2w, L.n2b, n.data());
        linear(L.lin1, n.data(), T, hid.data(), c->ar);
        for (size_t i = 0; i < (size_t)T * 4 * D; i++) hid[i] = gelu_ggml(hid[i]);
        linear(L.lin2, hid.data(), T, upd.data(), c->ar);
        for (size_t i = 0; i < (size_t)T * D; i++) x[i] = x[i] + upd[i];
        if (c->taps) s->tap["flow.layer" + std::to_string(l)] = x;
    }
    s->current_end += T;   // increment_states (src/pocket_tts.cpp:96)
}

// out_norm + EOS logit + 1-step LSD head (src/pocket_tts/models/flow_lm.h:114-142, modules/mlp.h:233-251)
void flow_head(oracle_ctx* c, const float* h_last, const float* noise, float* latent, float* eos_logit) {
    float cvec[1024];
    layer_norm(h_last, 1, 1024, 1e-5f, c->onw, c->onb, cvec);
    float e; linear(c->out_eos, cvec, 1, &e, c->ar);
    *eos_logit = e - (-4.0f);                                        // ggml_sub(out_eos, eos_threshold), ggml_sum over 1 elt
    float x[512], y[512], cond[512];
    linear(c->input_proj, noise, 1, x, c->ar);
    linear(c->cond_embed, cvec, 1, cond, c->ar);
    for (int i = 0; i < 512; i++) y[i] = c->t_combined[i] + cond[i];
    float sy[512]; for (int i = 0; i < 512; i++) sy[i] = silu(y[i]);
    float mod[1536], hn[512], h1[512], h2[512];
    for (int r = 0; r < 6; r++) {
        const ResBlock& R = c->rb[r];
        linear(R.ada, sy, 1, mod, c->ar);                            // [shift | scale | gate] (torch_chunk_3)
        layer_norm(x, 1, 512, 1e-6f, R.lnw, R.lnb, hn);
        for (int i = 0; i < 512; i++) hn[i] = hn[i] * (mod[512 + i] + 1.f) + mod[i];   // mlp_modulate
        linear(R.mlp0, hn, 1, h1, c->ar);
        for (int i = 0; i < 512; i++) h1[i] = silu(h1[i]);
        linear(R.mlp2, h1, 1, h2, c->ar);
        for (int i = 0; i < 512; i++) x[i] = x[i] + mod[1024 + i] * h2[i];
    }
    float mod2[1024];
    linear(c->final_ada, sy, 1, mod2, c->ar);                        // [shift | scale] (torch_chunk_2)
    layer_norm(x, 1, 512, 1e-6f, c->fnw, c->fnb, hn);
    for (int i = 0; i < 512; i++) hn[i] = hn[i] * (mod2[512 + i] + 1.f) + mod2[i];
    float v[32];
    linear(c->final_lin, hn, 1, v, c->ar);
    for (int i = 0; i < 32; i++) latent[i] = noise[i] + v[i];
}

// ggml_conv_1d with carried causal state (src/pocket_tts/modules/conv.h:48-87): xin = [state | x], state <- last K-1 rows
void conv1d_stream(const Conv& cv, std::vector<float>* state, const float* x, int T, float* y) {
    const int K = cv.k, Cin = cv.cin, Cout = cv.cout, TP = K - 1;
    std::vector<float> xin((size_t)(TP + T) * Cin);
    if (TP) memcpy(xin.data(), state->data(), (size_t)TP * Cin * sizeof(float));
    memcpy(xin.data() + (size_t)TP * Cin, x, (size_t)T * Cin * sizeof(float));
    if (TP) memcpy(state->data(), xin.data() + (size_t)T * Cin, (size_t)TP * Cin * sizeof(float));
    for (auto& v : xin) v = f16r(v);                                 // im2col emits F16
    const int win = K * Cin;
#pragma omp parallel for schedule(static) num_threads(g_threads) if ((long)T * Cout * win > 65536)
    for (int t = 0; t < T; t++) {
        const float* xv = xin.data() + (size_t)t * Cin;
        for (int co = 0; co < Cout; co++) {
            const float* w = cv.w.data() + (size_t)co * win;
            float acc = 0.f;
#pragma omp simd reduction(+ : acc)
            for (int i = 0; i < win; i++) acc += w[i] * xv[i];
            y[(size_t)t * Cout + co] = acc + (cv.b ? cv.b[co] : 0.f);
        }
    }
}

// ggml_conv_transpose_1d + overlap-add with the carried previous output (src/pocket_tts/modules/conv.h:272-332)
void convt_stream(const ConvT& cv, std::vector<float>& state, const float* x, int T, float* out) {
    const int K = cv.k, S = cv.s, Cin = cv.cin, Cout = cv.cout, PT = K - S;
    const int Lfull = (T - 1) * S + K;
    std::vector<float> y((size_t)Lfull * Cout, 0.f);
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int co = 0; co < Cout; co++) {
        for (int t = 0; t < T; t++) {
            const float* xv = x + (size_t)t * Cin;
            for (int k = 0; k < K; k++) {
                const float* w = cv.w.data() + ((size_t)k * Cout + co) * Cin;
                float acc = 0.f;
#pragma omp simd reduction(+ : acc)
                for (int ci = 0; ci < Cin; ci++) acc += w[ci] * xv[ci];
                y[(size_t)(t * S + k) * Cout + co] += acc;
            }
        }
    }
    // y[..., :PT] += prev_y[..., -PT:] ; prev_y = y (pre-bias) ; out = y[..., :-PT] + bias
    for (int i = 0; i < PT; i++) for (int co = 0; co < Cout; co++)
        y[(size_t)i * Cout + co] += state[(size_t)(Lfull - PT + i) * Cout + co];
    state = y;
    for (int i = 0; i < T * S; i++) for (int co = 0; co < Cout; co++)
        out[(size_t)i * Cout + co] = y[(size_t)i * Cout + co] + (cv.b ? cv.b[co] : 0.f);
}

void resblock(const Conv& a, const Conv& b, std::vector<float>& state, std::vector<float>& x, int T) {
    // src/pocket_tts/modules/seanet.h:14-27
    const int C = a.cin;
    std::vector<float> v((size_t)T * C), h((size_t)T * a.cout), o((size_t)T * C);
    for (size_t i = 0; i < v.size(); i++) v[i] = elu(x[i]);
    conv1d_stream(a, &state, v.data(), T, h.data());
    for (auto& e : h) e = elu(e);
    conv1d_stream(b, nullptr, h.data(), T, o.data());
    for (size_t i = 0; i < x.size(); i++) x[i] = x[i] + o[i];
}

// latent[32] -> 1920 samples. src/pocket_tts.cpp:472-485, models/mimi.h:77-104
void mimi_frame(oracle_stream* s, const float* latent, float* pcm) {
    oracle_ctx* c = s->ctx;
    // de-normalise + quantizer output_proj (1x1 conv, F16 weights, F16 im2col)
    float z[32], e[512];
    for (int i = 0; i < 32; i++) z[i] = f16r(c->emb_std[i] * latent[i] + c->emb_mean[i]);
    for (int o = 0; o < 512; o++) { float acc = 0.f; for (int i = 0; i < 32; i++) acc += c->quant_w[o * 32 + i] * z[i]; e[o] = acc; }
    // depthwise convT K=32 s=16 on ONE input step (src/pocket_tts/modules/conv.h:283-331)
    std::vector<float> y((size_t)32 * 512);
    for (int k = 0; k < 32; k++) for (int ch = 0; ch < 512; ch++) y[(size_t)k * 512 + ch] = e[ch] * c->up_w[ch * 32 + k];
    for (int k = 0; k < 16; k++) for (int ch = 0; ch < 512; ch++) y[(size_t)k * 512 + ch] += s->up_state[(size_t)(16 + k) * 512 + ch];
    s->up_state = y;
    const int T = 16, D = 512, H = 8, Dh = 64;
    std::vector<float> x((size_t)T * D);
    for (int k = 0; k < 16; k++) for (int ch = 0; ch < 512; ch++) x[(size_t)k * 512 + ch] = y[(size_t)k * 512 + ch] + (c->up_b ? c->up_b[ch] : 0.f);
    if (c->taps) s->tap["mimi.upsample"] = x;

    // Mimi decoder transformer (src/pocket_tts/modules/mimi_transformer.h:586-712,912-973,1184-1217)
    const int off = s->mimi_offset;
    std::vector<float> n((size_t)T * D), qkv((size_t)T * 3 * D), q((size_t)T * D), att((size_t)T * D), upd((size_t)T * D), hid((size_t)T * 4 * D);
    for (int l = 0; l < 2; l++) {
        const MimiLayer& L = c->ml[l];
        layer_norm(x.data(), T, D, 0.0f, L.n1w, L.n1b, n.data());    // eps = 0 (models/defaults.h:14)
        linear(L.in_proj, n.data(), T, qkv.data(), c->ar);
        float* Kc = s->mK[l].data(); float* Vc = s->mV[l].data();
        for (int t = 0; t < T; t++) {
            const float ts = (float)t + (float)off;
            const int slot = (off + t) % 250;                         // indices (mimi_transformer.h:1206-1210)
            for (int h = 0; h < H; h++) {
                const float* qs = &qkv[(size_t)t * 3 * D + h * Dh];
                const float* ks = qs + D; const float* vs = qs + 2 * D;
                float* qd = &q[(size_t)t * D + h * Dh];
                float* kd = &Kc[(size_t)slot * D + h * Dh];
                float* vd = &Vc[(size_t)slot * D + h * Dh];
                for (int i = 0; i < 32; i++) {
                    float rad = ts * c->rope_freq_mimi[i];            // ggml_timestep_embedding (rope.h:8-20)
                    float cr = cosf(rad), sr = sinf(rad);
                    float a = qs[2 * i], b = qs[2 * i + 1];
                    qd[i] = bf16r(a * cr - b * sr); qd[32 + i] = bf16r(a * sr + b * cr);  // q -> bf16 by the bf16 mul_mat
                    a = ks[2 * i]; b = ks[2 * i + 1];
                    kd[i] = bf16r(a * cr - b * sr); kd[32 + i] = bf16r(a * sr + b * cr);  // ggml_set_rows casts to BF16
                }
                for (int d = 0; d < Dh; d++) vd[d] = bf16r(vs[d]);
            }
        }
        for (int t = 0; t < T; t++) {
            const float* brow = pattern_row(c, off, t);
            for (int h = 0; h < H; h++) {
                float sc[250];
                const float* qv = &q[(size_t)t * D + h * Dh];
                for (int j = 0; j < 250; j++) {
                    const float* kv = &Kc[(size_t)j * D + h * Dh];
                    float acc = 0.f; for (int d = 0; d < Dh; d++) acc += kv[d] * qv[d];
                    sc[j] = acc;
                }
                softmax_row(sc, 250, 0.125f, brow);
                for (int j = 0; j < 250; j++) sc[j] = bf16r(sc[j]);   // probs -> bf16 by the bf16 V mul_mat
                float* o = &att[(size_t)t * D + h * Dh];
                for (int d = 0; d < Dh; d++) {
                    float acc = 0.f; for (int j = 0; j < 250; j++) acc += Vc[(size_t)j * D + h * Dh + d] * sc[j];
                    o[d] = acc;
                }
            }
        }
        linear(L.out_proj, att.data(), T, upd.data(), c->ar);
        for (int t = 0; t < T; t++) for (int i = 0; i < D; i++) x[(size_t)t * D + i] += upd[(size_t)t * D + i] * L.ls1[i];
        layer_norm(x.data(), T, D, 0.0f, L.n2w, L.n2b, n.data());
        linear(L.lin1, n.data(), T, hid.data(), c->ar);
        for (auto& v : hid) v = gelu_ggml(v);
        linear(L.lin2, hid.data(), T, upd.data(), c->ar);
        for (int t = 0; t < T; t++) for (int i = 0; i < D; i++) x[(size_t)t * D + i] += upd[(size_t)t * D + i] * L.ls2[i];
    }
    s->mimi_offset += T;
    if (c->taps) s->tap["mimi.transformer"] = x;

    // SEANet decoder (src/pocket_tts/modules/seanet.h:187-211)
    std::vector<float> a((size_t)16 * 512);
    conv1d_stream(c->c0, &s->s0, x.data(), 16, a.data());
    for (auto& v : a) v = elu(v);
    if (c->taps) s->tap["seanet.conv0"] = a;
    std::vector<float> b((size_t)96 * 256);
    convt_stream(c->t2, s->st2, a.data(), 16, b.data());
    if (c->taps) s->tap["seanet.convt2"] = b;
    resblock(c->r3a, c->r3b, s->s3, b, 96);
    for (auto& v : b) v = elu(v);
    if (c->taps) s->tap["seanet.res3"] = b;
    std::vector<float> d((size_t)480 * 128);
    convt_stream(c->t5, s->st5, b.data(), 96, d.data());
    resblock(c->r6a, c->r6b, s->s6, d, 480);
    for (auto& v : d) v = elu(v);
    if (c->taps) s->tap["seanet.res6"] = d;
    std::vector<float> g((size_t)1920 * 64);
    convt_stream(c->t8, s->st8, d.data(), 480, g.data());
    resblock(c->r9a, c->r9b, s->s9, g, 1920);
    for (auto& v : g) v = elu(v);
    if (c->taps) s->tap["seanet.res9"] = g;
    conv1d_stream(c->c11, &s->s11, g.data(), 1920, pcm);
}

void mimi_reset(oracle_stream* s) {   // init(mimi_states): zero conv states, offset = 0 (models/mimi.h:71-75)
    s->up_state.assign((size_t)32 * 512, 0.f);
    // NOTE: the reference does NOT clear the KV ring on init (mimi_transformer.h init only resets offset);
    // stale slots are masked (-inf) until rewritten while offset <= 250, so zeros vs stale are equivalent there.
    s->mimi_offset = 0;
    s->s0.assign((size_t)6 * 512, 0.f); s->s3.assign((size_t)2 * 256, 0.f); s->s6.assign((size_t)2 * 128, 0.f);
    s->s9.assign((size_t)2 * 64, 0.f); s->s11.assign((size_t)2 * 64, 0.f);
    s->st2.assign((size_t)(15 * 6 + 12) * 256, 0.f); s->st5.assign((size_t)(95 * 5 + 10) * 128, 0.f);
    s->st8.assign((size_t)(479 * 4 + 8) * 64, 0.f);
}

}  // namespace

// =====================================================================================
// C API
// =====================================================================================
extern "C" {

oracle_ctx* oracle_create(int file_bf16, int n_threads) {
    auto* c = new oracle_ctx;
    c->file_bf16 = file_bf16; c->ar = file_bf16 ? ACT_BF16 : ACT_F32;
    g_threads = n_threads > 0 ? n_threads : 1;
    return c;
}
void oracle_set_threads(int n) { g_threads = n > 0 ? n : 1; }
void oracle_enable_taps(oracle_ctx* c, int on) { c->taps = on != 0; }

// data: float32 values as stored in the file (BF16 files: already bf16-exact), torch-order shape
void oracle_set_tensor(oracle_ctx* c, const char* name, const float* data, const int64_t* shape, int ndim) {
    Named n; size_t cnt = 1;
    for (int i = 0; i < ndim; i++) { n.shape.push_back(shape[i]); cnt *= (size_t)shape[i]; }
    n.data.assign(data, data + cnt);
    c->T[name] = std::move(n);
}

int oracle_finalize(oracle_ctx* c) {
    c->n_embed = (int)c->T.at("flow_lm.conditioner.embed.weight").shape[0];
    c->embed = need(c, "flow_lm.conditioner.embed.weight", {c->n_embed, 1024});
    c->emb_std = need(c, "flow_lm.emb_std", {32}); c->emb_mean = need(c, "flow_lm.emb_mean", {32});
    c->bos_emb = need(c, "flow_lm.bos_emb", {32});
    c->input_linear = get_lin(c, "flow_lm.input_linear", 1024, 32);
    c->out_eos = get_lin(c, "flow_lm.out_eos", 1, 1024);
    c->onw = need(c, "flow_lm.out_norm.weight", {1024}); c->onb = need(c, "flow_lm.out_norm.bias", {1024}, true);
    for (int l = 0; l < 6; l++) {
        std::string p = "flow_lm.transformer.layers." + std::to_string(l) + ".";
        FlowLayer& L = c->fl[l];
        L.in_proj = get_lin(c, p + "self_attn.in_proj", 3072, 1024); L.out_proj = get_lin(c, p + "self_attn.out_proj", 1024, 1024);
        L.lin1 = get_lin(c, p + "linear1", 4096, 1024); L.lin2 = get_lin(c, p + "linear2", 1024, 4096);
        L.n1w = need(c, p + "norm1.weight", {1024}); L.n1b = need(c, p + "norm1.bias", {1024}, true);
        L.n2w = need(c, p + "norm2.weight", {1024}); L.n2b = need(c, p + "norm2.bias", {1024}, true);
    }
    std::string f = "flow_lm.flow_net.";
    c->input_proj = get_lin(c, f + "input_proj", 512, 32); c->cond_embed = get_lin(c, f + "cond_embed", 512, 1024);
    for (int r = 0; r < 6; r++) {
        std::string p = f + "res_blocks." + std::to_string(r) + ".";
        ResBlock& R = c->rb[r];
        R.lnw = need(c, p + "in_ln.weight", {512}, true); R.lnb = need(c, p + "in_ln.bias", {512}, true);
        R.mlp0 = get_lin(c, p + "mlp.0", 512, 512); R.mlp2 = get_lin(c, p + "mlp.2", 512, 512);
        R.ada = get_lin(c, p + "adaLN_modulation.1", 1536, 512);
    }
    c->final_lin = get_lin(c, f + "final_layer.linear", 32, 512);
    c->final_ada = get_lin(c, f + "final_layer.adaLN_modulation.1", 1024, 512);
    c->fnw = need(c, f + "final_layer.norm_final.weight", {512}, true);
    c->fnb = need(c, f + "final_layer.norm_final.bias", {512}, true);
    // t_combined = (TE1(t=1) + TE0(s=0)) / 2  (modules/mlp.h:241-245)
    float te1[512], te0[512];
    timestep_embed(c, 1, 1.0f, te1); timestep_embed(c, 0, 0.0f, te0);
    c->t_combined.resize(512);
    for (int i = 0; i < 512; i++) c->t_combined[i] = (te1[i] + te0[i]) * 0.5f;
    // Mimi
    const float* qw = need(c, "mimi.quantizer.output_proj.weight", {512, 32, 1});
    c->quant_w.resize(512 * 32); for (int i = 0; i < 512 * 32; i++) c->quant_w[i] = f16r(qw[i]);
    const float* uw = need(c, "mimi.upsample.convtr.convtr.weight", {512, 1, 32});
    c->up_w.assign(uw, uw + 512 * 32);
    c->up_b = need(c, "mimi.upsample.convtr.convtr.bias", {512}, true);
    for (int l = 0; l < 2; l++) {
        std::string p = "mimi.decoder_transformer.transformer.layers." + std::to_string(l) + ".";
        MimiLayer& L = c->ml[l];
        L.in_proj = get_lin(c, p + "self_attn.in_proj", 1536, 512); L.out_proj = get_lin(c, p + "self_attn.out_proj", 512, 512);
        L.lin1 = get_lin(c, p + "linear1", 2048, 512); L.lin2 = get_lin(c, p + "linear2", 512, 2048);
        L.n1w = need(c, p + "norm1.weight", {512}); L.n1b = need(c, p + "norm1.bias", {512}, true);
        L.n2w = need(c, p + "norm2.weight", {512}); L.n2b = need(c, p + "norm2.bias", {512}, true);
        L.ls1 = need(c, p + "layer_scale_1.scale", {512}); L.ls2 = need(c, p + "layer_scale_2.scale", {512});
    }
    std::string d = "mimi.decoder.model.";
    c->c0 = get_conv(c, d + "0", 512, 512, 7);
    c->t2 = get_convt(c, d + "2", 512, 256, 12, 6);
    c->r3a = get_conv(c, d + "3.block.1", 128, 256, 3); c->r3b = get_conv(c, d + "3.block.3", 256, 128, 1);
    c->t5 = get_convt(c, d + "5", 256, 128, 10, 5);
    c->r6a = get_conv(c, d + "6.block.1", 64, 128, 3); c->r6b = get_conv(c, d + "6.block.3", 128, 64, 1);
    c->t8 = get_convt(c, d + "8", 128, 64, 8, 4);
    c->r9a = get_conv(c, d + "9.block.1", 32, 64, 3); c->r9b = get_conv(c, d + "9.block.3", 64, 32, 1);
    c->c11 = get_conv(c, d + "11", 1, 64, 3);
    make_pattern(c);
    for (int i = 0; i < 32; i++) {
        // FlowLM: freqs = exp(arange * (-logf(max_period)/D_half))  (rope.h:36-38: ggml_scale then ggml_exp)
        c->rope_freq_flow[i] = expf((float)i * (-logf(10000.0f) / 32));
        // Mimi: ggml_timestep_embedding: freq = expf(-logf(max_period) * j / half)
        c->rope_freq_mimi[i] = (float)expf(-logf(10000.0f) * i / 32);
    }
    c->finalized = true;
    return 0;
}

void oracle_destroy(oracle_ctx* c) { delete c; }
void oracle_get_t_combined(oracle_ctx* c, float* out512) { memcpy(out512, c->t_combined.data(), 512 * sizeof(float)); }
// bias row j (250 floats) of the Mimi attention mask at stream offset `offset` (src/torch.h:203-221)
void oracle_mimi_bias_row(oracle_ctx* c, int offset, int j, float* out250) { memcpy(out250, pattern_row(c, offset, j), 250 * sizeof(float)); }

// ptts_stream_from_safetensors (src/pocket_tts.cpp:351-394) + get_state_for_audio_prompt (:100-124):
// allocate states and prefill the voice prefix audio_prompt [T_voice][1024] into the conditioned KV.
oracle_stream* oracle_stream_create(oracle_ctx* c, const float* audio_prompt, int T_voice, int kv_capacity) {
    auto* s = new oracle_stream; s->ctx = c; s->cap = kv_capacity;
    for (int l = 0; l < 6; l++) { s->K[l].assign((size_t)kv_capacity * 1024, 0.f); s->V[l].assign((size_t)kv_capacity * 1024, 0.f); }
    for (int l = 0; l < 2; l++) { s->mK[l].assign((size_t)250 * 512, 0.f); s->mV[l].assign((size_t)250 * 512, 0.f); }
    mimi_reset(s);
    s->current_end = 0;
    std::vector<float> x(audio_prompt, audio_prompt + (size_t)T_voice * 1024);
    if (T_voice > 0) flowlm_forward(s, x, T_voice);
    for (int l = 0; l < 6; l++) { s->cK[l] = s->K[l]; s->cV[l] = s->V[l]; }
    s->cond_end = s->current_end;
    s->max_gen_len = 0; s->generation_step = 0;           // ptts_stream_reset
    return s;
}
void oracle_stream_destroy(oracle_stream* s) { delete s; }
// A second stream of the same voice without repeating the voice prefill (the reference would run get_state_for_audio_prompt again and
// arrive at the same conditioned state: the prefill is deterministic). Test convenience for long prefixes.
oracle_stream* oracle_stream_clone(const oracle_stream* s) { return new oracle_stream(*s); }

// _stream_sentence_init (src/pocket_tts.cpp:416-444): restore conditioned KV, reset Mimi, text prefill.
void oracle_sentence_init(oracle_stream* s, const int* tokens, int n_tokens, int max_gen_len, int frames_after_eos) {
    oracle_ctx* c = s->ctx;
    for (int l = 0; l < 6; l++) { s->K[l] = s->cK[l]; s->V[l] = s->cV[l]; }   // copy_states (flow_lm.h:70-78)
    s->current_end = s->cond_end;
    mimi_reset(s);
    if (n_tokens > 0) {
        std::vector<float> x((size_t)n_tokens * 1024);
        for (int t = 0; t < n_tokens; t++) {
            int id = tokens[t];
            if (id < 0 || id >= c->n_embed) { fprintf(stderr, "oracle: token id %d out of range\n", id); exit(1); }
            memcpy(&x[(size_t)t * 1024], c->embed + (size_t)id * 1024, 1024 * sizeof(float));   // ggml_get_rows (text.h:29-37)
        }
        flowlm_forward(s, x, n_tokens);
    }
    s->frames_after_eos = frames_after_eos; s->max_gen_len = max_gen_len;
    memcpy(s->backbone_input, c->bos_emb, 32 * sizeof(float));
    s->generation_step = 0; s->eos_step = -1;
}

// _stream_sentence_step (src/pocket_tts.cpp:446-492). noise32 may be NULL (== temp 0). Returns 1 and fills
// pcm[1920]/latent[32] when a frame was produced, 0 when the sentence is finished.
int oracle_step(oracle_stream* s, const float* noise32, float* latent32, float* pcm1920, float* eos_logit) {
    oracle_ctx* c = s->ctx;
    if (s->generation_step >= s->max_gen_len) return 0;
    std::vector<float> x(1024);
    linear(c->input_linear, s->backbone_input, 1, x.data(), c->ar);
    flowlm_forward(s, x, 1);
    float zero[32] = {0}; float lat[32]; float e;
    flow_head(c, x.data(), noise32 ? noise32 : zero, lat, &e);
    if (eos_logit) *eos_logit = e;
    bool is_eos = e > 0.f;
    if (is_eos && s->eos_step == -1) s->eos_step = s->generation_step;
    if (s->eos_step != -1 && s->generation_step >= s->eos_step + s->frames_after_eos) {
        s->generation_step = s->max_gen_len;
        return 0;
    }
    mimi_frame(s, lat, pcm1920);
    memcpy(s->backbone_input, lat, sizeof(lat));
    if (latent32) memcpy(latent32, lat, sizeof(lat));
    s->generation_step++;
    return 1;
}

// ---- module-level entry points for kernel-level parity tests ----
// FlowLM transformer over T rows x[T][1024] (in place -> pre-out_norm hidden), appends KV, advances current_end.
void oracle_flowlm_rows(oracle_stream* s, float* x, int T) {
    std::vector<float> v(x, x + (size_t)T * 1024);
    flowlm_forward(s, v, T);
    memcpy(x, v.data(), v.size() * sizeof(float));
}
void oracle_input_linear(oracle_ctx* c, const float* latent32, float* out1024) { linear(c->input_linear, latent32, 1, out1024, c->ar); }
void oracle_flow_head(oracle_ctx* c, const float* h1024, const float* noise32, float* latent32, float* eos_logit) {
    flow_head(c, h1024, noise32, latent32, eos_logit);
}
void oracle_mimi_reset(oracle_stream* s) { mimi_reset(s); }
void oracle_mimi_frame(oracle_stream* s, const float* latent32, float* pcm1920) { mimi_frame(s, latent32, pcm1920); }
int oracle_current_end(oracle_stream* s) { return s->current_end; }
// teacher forcing: the next step's backbone input (the reference keeps it in states->output, src/pocket_tts.cpp:487)
void oracle_set_backbone_input(oracle_stream* s, const float* latent32) { memcpy(s->backbone_input, latent32, 32 * sizeof(float)); }
int oracle_mimi_offset(oracle_stream* s) { return s->mimi_offset; }
void oracle_set_mimi_offset(oracle_stream* s, int off) { s->mimi_offset = off; }
// copies KV rows [0,current_end) of layer l: out [current_end][1024]
void oracle_get_kv(oracle_stream* s, int layer, int which, float* out) {
    const auto& src = which == 0 ? s->K[layer] : s->V[layer];
    memcpy(out, src.data(), (size_t)s->current_end * 1024 * sizeof(float));
}
int oracle_get_tap(oracle_stream* s, const char* name, float* out, int max_elems) {
    auto it = s->tap.find(name);
    if (it == s->tap.end()) return -1;
    int n = (int)std::min<size_t>(it->second.size(), (size_t)max_elems);
    if (out) memcpy(out, it->second.data(), (size_t)n * sizeof(float));
    return (int)it->second.size();
}
int oracle_num_threads_available() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
