// ggml.h — STAND-IN for the ggml headers, TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// The reference (Codes4Fun/pocket-tts.cpp) builds its graphs with ggml, which is neither vendored nor pinned
// (cmake/FindGGML.cmake:11-34) and absent from this image. This header lets the reference's OWN graph code
// (src/pocket_tts.cpp and the headers it includes) compile UNCHANGED: it declares the ~90 ggml entry points those sources
// use and implements them as a small eager CPU interpreter — tensors, views, a node list, and one scalar/OpenMP kernel per
// op. Only the op SEMANTICS are restated here (from ggml's published behaviour): shapes are
// ne[0]-fastest, mul_mat rounds the f32 operand to the weight type's vec_dot type (F16 / BF16) and accumulates in f32,
// ggml_norm uses double sums and a biased variance, ggml_gelu goes through an f16 table, conv_1d = im2col(F16) + f16
// mul_mat, conv_transpose_1d accumulates in f32, soft_max_ext = softmax(x * scale + mask) with a double sum.
// The graph STRUCTURE that runs is the reference's, not a restatement. Nothing in the product includes this file.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <time.h>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GGML_MAX_DIMS 4
#define GGML_MAX_NAME 64
#define GGML_MAX_SRC 10
#define GGML_DEFAULT_GRAPH_SIZE 2048

enum ggml_type {
    GGML_TYPE_F32 = 0, GGML_TYPE_F16 = 1, GGML_TYPE_Q4_0 = 2, GGML_TYPE_Q8_0 = 8, GGML_TYPE_Q4_K = 12, GGML_TYPE_Q8_K = 15,
    GGML_TYPE_I32 = 26, GGML_TYPE_I64 = 27, GGML_TYPE_BF16 = 30,
};
enum ggml_op {
    GGML_OP_NONE = 0, GGML_OP_DUP, GGML_OP_ADD, GGML_OP_SUB, GGML_OP_MUL, GGML_OP_DIV, GGML_OP_SCALE, GGML_OP_NEG, GGML_OP_EXP, GGML_OP_SIN, GGML_OP_COS,
    GGML_OP_SQRT, GGML_OP_SILU, GGML_OP_GELU, GGML_OP_ELU, GGML_OP_CLAMP, GGML_OP_NORM, GGML_OP_RMS_NORM, GGML_OP_MUL_MAT, GGML_OP_SOFT_MAX,
    GGML_OP_CPY, GGML_OP_CONT, GGML_OP_CONCAT, GGML_OP_VIEW, GGML_OP_RESHAPE, GGML_OP_PERMUTE, GGML_OP_TRANSPOSE, GGML_OP_REPEAT,
    GGML_OP_IM2COL, GGML_OP_CONV_TRANSPOSE_1D, GGML_OP_GET_ROWS, GGML_OP_SET_ROWS, GGML_OP_TIMESTEP_EMBEDDING, GGML_OP_ARANGE, GGML_OP_SUM,
    GGML_OP_MEAN, GGML_OP_PAD,
};

struct ggml_backend_buffer { std::vector<char> mem; };
typedef ggml_backend_buffer* ggml_backend_buffer_t;
struct ggml_backend { int n_threads = 1; };
typedef ggml_backend* ggml_backend_t;

struct ggml_tensor {
    ggml_type type = GGML_TYPE_F32;
    ggml_backend_buffer* buffer = nullptr;
    int64_t ne[GGML_MAX_DIMS] = {1, 1, 1, 1};
    size_t nb[GGML_MAX_DIMS] = {0, 0, 0, 0};
    ggml_op op = GGML_OP_NONE;
    int32_t op_params[16] = {0};
    ggml_tensor* src[GGML_MAX_SRC] = {nullptr};
    ggml_tensor* view_src = nullptr;
    size_t view_offs = 0;
    void* data = nullptr;
    char name[GGML_MAX_NAME] = {0};
    bool owns_data = false;
};

struct ggml_init_params { size_t mem_size; void* mem_buffer; bool no_alloc; };
struct ggml_context { bool no_alloc = true; std::vector<ggml_tensor*> tensors; };
struct ggml_cgraph { std::vector<ggml_tensor*> nodes; std::unordered_set<ggml_tensor*> seen; };

// ---------------------------------------------------------------------------------------------------------------
// scalar conversions (GGML_FP32_TO_BF16: round to nearest even; FP16: IEEE half via the compiler's _Float16)
// ---------------------------------------------------------------------------------------------------------------
static inline uint16_t ggml_shim_f32_to_bf16(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    if ((u & 0x7fffffff) > 0x7f800000) return (uint16_t)((u >> 16) | 64);
    return (uint16_t)((u + (0x7fff + ((u >> 16) & 1))) >> 16);
}
static inline float ggml_shim_bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static inline uint16_t ggml_shim_f32_to_f16(float x) { _Float16 h = (_Float16)x; uint16_t u; memcpy(&u, &h, 2); return u; }
static inline float ggml_shim_f16_to_f32(uint16_t u) { _Float16 h; memcpy(&h, &u, 2); return (float)h; }

static inline size_t ggml_type_size(ggml_type t) {
    switch (t) { case GGML_TYPE_F32: case GGML_TYPE_I32: return 4; case GGML_TYPE_F16: case GGML_TYPE_BF16: return 2; case GGML_TYPE_I64: return 8;
                 default: fprintf(stderr, "ggml shim: quantised types are not supported\n"); abort(); }
}
static inline size_t ggml_row_size(ggml_type t, int64_t ne) { return ggml_type_size(t) * (size_t)ne; }
static inline int64_t ggml_nelements(const ggml_tensor* t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }
static inline size_t ggml_nbytes(const ggml_tensor* t) {
    size_t n = ggml_type_size(t->type);
    for (int i = 0; i < GGML_MAX_DIMS; i++) n += (size_t)(t->ne[i] - 1) * t->nb[i];
    return n;
}
static inline size_t ggml_tensor_overhead(void) { return sizeof(ggml_tensor) + 32; }
static inline bool ggml_is_contiguous(const ggml_tensor* t) {
    size_t s = ggml_type_size(t->type);
    for (int i = 0; i < GGML_MAX_DIMS; i++) { if (t->ne[i] != 1 && t->nb[i] != s) return false; s *= (size_t)t->ne[i]; }
    return true;
}
static inline int64_t ggml_time_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (int64_t)ts.tv_sec * 1000 + ts.tv_nsec / 1000000; }
static inline void ggml_time_init(void) {}

// ---------------------------------------------------------------------------------------------------------------
// contexts and tensors
// ---------------------------------------------------------------------------------------------------------------
static inline ggml_context* ggml_init(ggml_init_params p) { auto* c = new ggml_context; c->no_alloc = p.no_alloc; return c; }
static inline void ggml_reset(ggml_context* c) {
    for (auto* t : c->tensors) { if (t->owns_data) free(t->data); delete t; }
    c->tensors.clear();
}
static inline void ggml_free(ggml_context* c) { if (!c) return; ggml_reset(c); delete c; }

static inline ggml_tensor* ggml_shim_new(ggml_context* c, ggml_type type, const int64_t* ne, int n_dims, ggml_tensor* view_src = nullptr, size_t view_offs = 0) {
    auto* t = new ggml_tensor;
    t->type = type;
    for (int i = 0; i < GGML_MAX_DIMS; i++) t->ne[i] = i < n_dims ? ne[i] : 1;
    t->nb[0] = ggml_type_size(type);
    for (int i = 1; i < GGML_MAX_DIMS; i++) t->nb[i] = t->nb[i - 1] * (size_t)t->ne[i - 1];
    if (view_src) {
        if (view_src->view_src) { view_offs += view_src->view_offs; view_src = view_src->view_src; }     // ggml_new_tensor_impl
        t->view_src = view_src; t->view_offs = view_offs;
        if (view_src->data) t->data = (char*)view_src->data + view_offs;
        t->buffer = view_src->buffer;
    } else if (!c->no_alloc) {
        t->data = calloc(1, std::max<size_t>(ggml_nbytes(t), 16)); t->owns_data = true;
    }
    c->tensors.push_back(t);
    return t;
}
static inline ggml_tensor* ggml_new_tensor(ggml_context* c, ggml_type type, int n_dims, const int64_t* ne) { return ggml_shim_new(c, type, ne, n_dims); }
static inline ggml_tensor* ggml_new_tensor_1d(ggml_context* c, ggml_type type, int64_t ne0) { return ggml_shim_new(c, type, &ne0, 1); }
static inline ggml_tensor* ggml_new_tensor_2d(ggml_context* c, ggml_type type, int64_t ne0, int64_t ne1) { const int64_t ne[2] = {ne0, ne1}; return ggml_shim_new(c, type, ne, 2); }
static inline ggml_tensor* ggml_new_tensor_3d(ggml_context* c, ggml_type type, int64_t ne0, int64_t ne1, int64_t ne2) { const int64_t ne[3] = {ne0, ne1, ne2}; return ggml_shim_new(c, type, ne, 3); }
static inline ggml_tensor* ggml_new_tensor_4d(ggml_context* c, ggml_type type, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3) { const int64_t ne[4] = {ne0, ne1, ne2, ne3}; return ggml_shim_new(c, type, ne, 4); }
static inline ggml_tensor* ggml_dup_tensor(ggml_context* c, const ggml_tensor* a) { return ggml_shim_new(c, a->type, a->ne, GGML_MAX_DIMS); }
static inline ggml_tensor* ggml_new_i32(ggml_context* c, int32_t v) { auto* t = ggml_new_tensor_1d(c, GGML_TYPE_I32, 1); assert(t->data); *(int32_t*)t->data = v; return t; }
static inline ggml_tensor* ggml_new_f32(ggml_context* c, float v) { auto* t = ggml_new_tensor_1d(c, GGML_TYPE_F32, 1); assert(t->data); *(float*)t->data = v; return t; }
static inline ggml_tensor* ggml_set_name(ggml_tensor* t, const char* name) { snprintf(t->name, sizeof t->name, "%s", name); return t; }
static inline ggml_tensor* ggml_get_first_tensor(const ggml_context* c) { return c->tensors.empty() ? nullptr : c->tensors[0]; }
static inline ggml_tensor* ggml_get_next_tensor(const ggml_context* c, ggml_tensor* t) {
    for (size_t i = 0; i + 1 < c->tensors.size(); i++) if (c->tensors[i] == t) return c->tensors[i + 1];
    return nullptr;
}
static inline ggml_tensor* ggml_get_tensor(ggml_context* c, const char* name) { for (auto* t : c->tensors) if (!strcmp(t->name, name)) return t; return nullptr; }

// ---------------------------------------------------------------------------------------------------------------
// backend: one host buffer per context (ggml_backend_alloc_ctx_tensors), plain memcpy set/get
// ---------------------------------------------------------------------------------------------------------------
static inline ggml_backend_t ggml_backend_cpu_init(void) { return new ggml_backend; }
static inline void ggml_backend_cpu_set_n_threads(ggml_backend_t b, int n) { if (b) b->n_threads = n > 0 ? n : 1; }
static inline void ggml_backend_free(ggml_backend_t b) { delete b; }
static inline void ggml_backend_load_all(void) {}
static inline ggml_backend_buffer_t ggml_backend_alloc_ctx_tensors(ggml_context* c, ggml_backend_t) {
    size_t total = 0;
    for (auto* t : c->tensors) if (!t->data && !t->view_src) total += (ggml_nbytes(t) + 63) / 64 * 64;
    ggml_backend_buffer* buf = nullptr;
    if (total) { buf = new ggml_backend_buffer; buf->mem.assign(total + 64, 0); }
    size_t off = 0;
    for (auto* t : c->tensors) {
        if (t->data) continue;
        if (t->view_src) { assert(t->view_src->data && "view of an unallocated tensor"); t->data = (char*)t->view_src->data + t->view_offs; t->buffer = t->view_src->buffer; continue; }
        t->data = buf->mem.data() + off; t->buffer = buf; off += (ggml_nbytes(t) + 63) / 64 * 64;
    }
    return buf;
}
static inline void ggml_backend_buffer_free(ggml_backend_buffer_t b) { delete b; }
static inline void ggml_backend_tensor_set(ggml_tensor* t, const void* data, size_t offset, size_t size) { assert(t->data); memcpy((char*)t->data + offset, data, size); }
static inline void ggml_backend_tensor_get(const ggml_tensor* t, void* data, size_t offset, size_t size) { assert(t->data); memcpy(data, (const char*)t->data + offset, size); }

// ---------------------------------------------------------------------------------------------------------------
// graph
// ---------------------------------------------------------------------------------------------------------------
static inline ggml_cgraph* ggml_new_graph_custom(ggml_context*, size_t, bool) { return new ggml_cgraph; }   // leaked per graph (a few hundred bytes); test infrastructure
static inline void ggml_shim_visit(ggml_cgraph* g, ggml_tensor* t) {
    if (!t || g->seen.count(t)) return;
    g->seen.insert(t);
    for (int i = 0; i < GGML_MAX_SRC; i++) ggml_shim_visit(g, t->src[i]);
    if (t->op != GGML_OP_NONE) g->nodes.push_back(t);
}
static inline void ggml_build_forward_expand(ggml_cgraph* g, ggml_tensor* t) { ggml_shim_visit(g, t); }

// ---------------------------------------------------------------------------------------------------------------
// op constructors (deferred; executed by ggml_backend_graph_compute)
// ---------------------------------------------------------------------------------------------------------------
static inline ggml_tensor* ggml_shim_op(ggml_context* c, ggml_op op, ggml_type type, const int64_t* ne, ggml_tensor* a, ggml_tensor* b = nullptr, ggml_tensor* d = nullptr) {
    auto* t = ggml_shim_new(c, type, ne, GGML_MAX_DIMS);
    t->op = op; t->src[0] = a; t->src[1] = b; t->src[2] = d;
    return t;
}
static inline bool ggml_shim_can_repeat(const ggml_tensor* small, const ggml_tensor* big) {
    for (int i = 0; i < GGML_MAX_DIMS; i++) if (big->ne[i] % small->ne[i] != 0) return false;
    return true;
}
#define GGML_SHIM_BINARY(fn, OP) \
    static inline ggml_tensor* fn(ggml_context* c, ggml_tensor* a, ggml_tensor* b) { assert(ggml_shim_can_repeat(b, a)); return ggml_shim_op(c, OP, a->type, a->ne, a, b); }
GGML_SHIM_BINARY(ggml_add, GGML_OP_ADD)
GGML_SHIM_BINARY(ggml_sub, GGML_OP_SUB)
GGML_SHIM_BINARY(ggml_mul, GGML_OP_MUL)
GGML_SHIM_BINARY(ggml_div, GGML_OP_DIV)
static inline ggml_tensor* ggml_shim_view_of(ggml_context* c, ggml_tensor* a, ggml_op op) {
    auto* t = ggml_shim_new(c, a->type, a->ne, GGML_MAX_DIMS, a, 0);
    for (int i = 0; i < GGML_MAX_DIMS; i++) t->nb[i] = a->nb[i];
    t->op = op; t->src[0] = a;
    return t;
}
static inline ggml_tensor* ggml_add_inplace(ggml_context* c, ggml_tensor* a, ggml_tensor* b) {
    assert(ggml_shim_can_repeat(b, a));
    auto* t = ggml_shim_view_of(c, a, GGML_OP_ADD); t->src[1] = b; return t;
}
#define GGML_SHIM_UNARY(fn, OP) static inline ggml_tensor* fn(ggml_context* c, ggml_tensor* a) { return ggml_shim_op(c, OP, a->type, a->ne, a); }
GGML_SHIM_UNARY(ggml_neg, GGML_OP_NEG)
GGML_SHIM_UNARY(ggml_exp, GGML_OP_EXP)
GGML_SHIM_UNARY(ggml_sin, GGML_OP_SIN)
GGML_SHIM_UNARY(ggml_cos, GGML_OP_COS)
GGML_SHIM_UNARY(ggml_sqrt, GGML_OP_SQRT)
GGML_SHIM_UNARY(ggml_silu, GGML_OP_SILU)
GGML_SHIM_UNARY(ggml_gelu, GGML_OP_GELU)
GGML_SHIM_UNARY(ggml_elu, GGML_OP_ELU)
GGML_SHIM_UNARY(ggml_cont, GGML_OP_CONT)
GGML_SHIM_UNARY(ggml_dup, GGML_OP_DUP)
static inline void ggml_shim_setf(ggml_tensor* t, int i, float v) { memcpy(&t->op_params[i], &v, 4); }
static inline float ggml_shim_getf(const ggml_tensor* t, int i) { float v; memcpy(&v, &t->op_params[i], 4); return v; }
static inline ggml_tensor* ggml_scale(ggml_context* c, ggml_tensor* a, float s) { auto* t = ggml_shim_op(c, GGML_OP_SCALE, a->type, a->ne, a); ggml_shim_setf(t, 0, s); return t; }
static inline ggml_tensor* ggml_clamp(ggml_context* c, ggml_tensor* a, float lo, float hi) { auto* t = ggml_shim_op(c, GGML_OP_CLAMP, a->type, a->ne, a); ggml_shim_setf(t, 0, lo); ggml_shim_setf(t, 1, hi); return t; }
static inline ggml_tensor* ggml_norm(ggml_context* c, ggml_tensor* a, float eps) { auto* t = ggml_shim_op(c, GGML_OP_NORM, a->type, a->ne, a); ggml_shim_setf(t, 0, eps); return t; }
static inline ggml_tensor* ggml_rms_norm(ggml_context* c, ggml_tensor* a, float eps) { auto* t = ggml_shim_op(c, GGML_OP_RMS_NORM, a->type, a->ne, a); ggml_shim_setf(t, 0, eps); return t; }
static inline ggml_tensor* ggml_cast(ggml_context* c, ggml_tensor* a, ggml_type type) { auto* t = ggml_shim_op(c, GGML_OP_CPY, type, a->ne, a); t->src[1] = t; return t; }
static inline ggml_tensor* ggml_cpy(ggml_context* c, ggml_tensor* a, ggml_tensor* b) {
    assert(ggml_nelements(a) == ggml_nelements(b));
    auto* t = ggml_shim_view_of(c, b, GGML_OP_CPY); t->src[0] = a; t->src[1] = b; return t;
}
static inline ggml_tensor* ggml_mul_mat(ggml_context* c, ggml_tensor* a, ggml_tensor* b) {
    assert(a->ne[0] == b->ne[0] && b->ne[2] % a->ne[2] == 0 && b->ne[3] % a->ne[3] == 0);
    const int64_t ne[4] = {a->ne[1], b->ne[1], b->ne[2], b->ne[3]};
    return ggml_shim_op(c, GGML_OP_MUL_MAT, GGML_TYPE_F32, ne, a, b);
}
static inline ggml_tensor* ggml_soft_max_ext(ggml_context* c, ggml_tensor* a, ggml_tensor* mask, float scale, float max_bias) {
    assert(max_bias == 0.0f); (void)max_bias;
    if (mask) { assert(mask->ne[0] == a->ne[0] && mask->ne[1] >= a->ne[1] && a->ne[2] % mask->ne[2] == 0 && a->ne[3] % mask->ne[3] == 0); }
    auto* t = ggml_shim_op(c, GGML_OP_SOFT_MAX, a->type, a->ne, a, mask); ggml_shim_setf(t, 0, scale); return t;
}
static inline ggml_tensor* ggml_concat(ggml_context* c, ggml_tensor* a, ggml_tensor* b, int dim) {
    int64_t ne[4];
    for (int i = 0; i < 4; i++) { if (i == dim) ne[i] = a->ne[i] + b->ne[i]; else { assert(a->ne[i] == b->ne[i]); ne[i] = a->ne[i]; } }
    assert(a->type == b->type);
    auto* t = ggml_shim_op(c, GGML_OP_CONCAT, a->type, ne, a, b); t->op_params[0] = dim; return t;
}
static inline ggml_tensor* ggml_shim_view(ggml_context* c, ggml_tensor* a, int n_dims, const int64_t* ne, size_t offset) {
    auto* t = ggml_shim_new(c, a->type, ne, n_dims, a, offset);
    t->op = GGML_OP_VIEW; t->src[0] = a;
    return t;
}
static inline ggml_tensor* ggml_view_1d(ggml_context* c, ggml_tensor* a, int64_t ne0, size_t offset) { return ggml_shim_view(c, a, 1, &ne0, offset); }
static inline ggml_tensor* ggml_view_2d(ggml_context* c, ggml_tensor* a, int64_t ne0, int64_t ne1, size_t nb1, size_t offset) {
    const int64_t ne[2] = {ne0, ne1}; auto* t = ggml_shim_view(c, a, 2, ne, offset); t->nb[1] = nb1; t->nb[2] = t->nb[1] * ne1; t->nb[3] = t->nb[2]; return t;
}
static inline ggml_tensor* ggml_view_3d(ggml_context* c, ggml_tensor* a, int64_t ne0, int64_t ne1, int64_t ne2, size_t nb1, size_t nb2, size_t offset) {
    const int64_t ne[3] = {ne0, ne1, ne2}; auto* t = ggml_shim_view(c, a, 3, ne, offset); t->nb[1] = nb1; t->nb[2] = nb2; t->nb[3] = t->nb[2] * ne2; return t;
}
static inline ggml_tensor* ggml_view_4d(ggml_context* c, ggml_tensor* a, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3, size_t nb1, size_t nb2, size_t nb3, size_t offset) {
    const int64_t ne[4] = {ne0, ne1, ne2, ne3}; auto* t = ggml_shim_view(c, a, 4, ne, offset); t->nb[1] = nb1; t->nb[2] = nb2; t->nb[3] = nb3; return t;
}
static inline ggml_tensor* ggml_shim_reshape(ggml_context* c, ggml_tensor* a, int n_dims, const int64_t* ne) {
    assert(ggml_is_contiguous(a));
    auto* t = ggml_shim_new(c, a->type, ne, n_dims, a, 0);
    assert(ggml_nelements(t) == ggml_nelements(a));
    t->op = GGML_OP_RESHAPE; t->src[0] = a;
    return t;
}
static inline ggml_tensor* ggml_reshape_2d(ggml_context* c, ggml_tensor* a, int64_t ne0, int64_t ne1) { const int64_t ne[2] = {ne0, ne1}; return ggml_shim_reshape(c, a, 2, ne); }
static inline ggml_tensor* ggml_reshape_3d(ggml_context* c, ggml_tensor* a, int64_t ne0, int64_t ne1, int64_t ne2) { const int64_t ne[3] = {ne0, ne1, ne2}; return ggml_shim_reshape(c, a, 3, ne); }
static inline ggml_tensor* ggml_reshape_4d(ggml_context* c, ggml_tensor* a, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3) { const int64_t ne[4] = {ne0, ne1, ne2, ne3}; return ggml_shim_reshape(c, a, 4, ne); }
static inline ggml_tensor* ggml_permute(ggml_context* c, ggml_tensor* a, int ax0, int ax1, int ax2, int ax3) {
    auto* t = ggml_shim_view_of(c, a, GGML_OP_PERMUTE);
    const int ax[4] = {ax0, ax1, ax2, ax3};
    for (int i = 0; i < 4; i++) { t->ne[ax[i]] = a->ne[i]; t->nb[ax[i]] = a->nb[i]; }
    return t;
}
static inline ggml_tensor* ggml_transpose(ggml_context* c, ggml_tensor* a) {
    auto* t = ggml_shim_view_of(c, a, GGML_OP_TRANSPOSE);
    t->ne[0] = a->ne[1]; t->ne[1] = a->ne[0]; t->nb[0] = a->nb[1]; t->nb[1] = a->nb[0];
    return t;
}
static inline ggml_tensor* ggml_repeat_4d(ggml_context* c, ggml_tensor* a, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3) {
    const int64_t ne[4] = {ne0, ne1, ne2, ne3};
    auto* t = ggml_shim_op(c, GGML_OP_REPEAT, a->type, ne, a); assert(ggml_shim_can_repeat(a, t)); return t;
}
// ggml_conv_1d = im2col (F16) + mul_mat(im2col, kernel) reshaped to [OL, OC, N]
static inline ggml_tensor* ggml_conv_1d(ggml_context* c, ggml_tensor* a, ggml_tensor* b, int s0, int p0, int d0) {
    assert(a->ne[1] == b->ne[1]);
    const int64_t K = a->ne[0], IC = a->ne[1], OC = a->ne[2], L = b->ne[0], N = b->ne[2];
    const int64_t OL = (L + 2 * p0 - d0 * (K - 1) - 1) / s0 + 1;
    const int64_t ne_im[4] = {IC * K, OL, N, 1};
    auto* im = ggml_shim_op(c, GGML_OP_IM2COL, GGML_TYPE_F16, ne_im, a, b);
    im->op_params[0] = s0; im->op_params[1] = p0; im->op_params[2] = d0;
    auto* r = ggml_mul_mat(c, ggml_reshape_2d(c, im, IC * K, OL * N), ggml_reshape_2d(c, a, IC * K, OC));     // [OL * N, OC]
    return ggml_reshape_3d(c, r, OL, OC, N);
}
static inline ggml_tensor* ggml_conv_transpose_1d(ggml_context* c, ggml_tensor* a, ggml_tensor* b, int s0, int p0, int d0) {
    assert(p0 == 0 && d0 == 1 && a->ne[2] == b->ne[1] && b->ne[2] == 1); (void)p0; (void)d0;
    const int64_t ne[4] = {(b->ne[0] - 1) * s0 + a->ne[0], a->ne[1], 1, 1};
    auto* t = ggml_shim_op(c, GGML_OP_CONV_TRANSPOSE_1D, GGML_TYPE_F32, ne, a, b); t->op_params[0] = s0; return t;
}
static inline ggml_tensor* ggml_get_rows(ggml_context* c, ggml_tensor* a, ggml_tensor* b) {
    assert(b->type == GGML_TYPE_I32 && a->ne[2] == b->ne[1]);
    const int64_t ne[4] = {a->ne[0], b->ne[0], b->ne[1], b->ne[2]};
    return ggml_shim_op(c, GGML_OP_GET_ROWS, GGML_TYPE_F32, ne, a, b);
}
// dst rows selected by c (I32/I64 [n, ...]) receive the rows of b (F32), converted to dst's type; returns a view of dst
static inline ggml_tensor* ggml_set_rows(ggml_context* c, ggml_tensor* a, ggml_tensor* b, ggml_tensor* idx) {
    assert(a->ne[0] == b->ne[0] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3] && b->ne[1] == idx->ne[0] && b->type == GGML_TYPE_F32);
    assert(b->ne[2] % idx->ne[1] == 0 && b->ne[3] % idx->ne[2] == 0 && (idx->type == GGML_TYPE_I32 || idx->type == GGML_TYPE_I64));
    auto* t = ggml_shim_view_of(c, a, GGML_OP_SET_ROWS); t->src[0] = b; t->src[1] = idx; t->src[2] = a; return t;
}
static inline ggml_tensor* ggml_timestep_embedding(ggml_context* c, ggml_tensor* ts, int dim, int max_period) {
    const int64_t ne[4] = {dim, ts->ne[0], 1, 1};
    auto* t = ggml_shim_op(c, GGML_OP_TIMESTEP_EMBEDDING, GGML_TYPE_F32, ne, ts); t->op_params[0] = dim; t->op_params[1] = max_period; return t;
}
static inline ggml_tensor* ggml_arange(ggml_context* c, float start, float stop, float step) {
    const int64_t n = (int64_t)ceilf((stop - start) / step);
    auto* t = ggml_shim_op(c, GGML_OP_ARANGE, GGML_TYPE_F32, &n, nullptr); t->ne[1] = t->ne[2] = t->ne[3] = 1;
    t->nb[1] = t->nb[2] = t->nb[3] = t->nb[0] * (size_t)n;
    ggml_shim_setf(t, 0, start); ggml_shim_setf(t, 1, stop); ggml_shim_setf(t, 2, step); return t;
}
static inline ggml_tensor* ggml_sum(ggml_context* c, ggml_tensor* a) { const int64_t ne[4] = {1, 1, 1, 1}; return ggml_shim_op(c, GGML_OP_SUM, a->type, ne, a); }
static inline ggml_tensor* ggml_mean(ggml_context* c, ggml_tensor* a) { const int64_t ne[4] = {1, a->ne[1], a->ne[2], a->ne[3]}; return ggml_shim_op(c, GGML_OP_MEAN, GGML_TYPE_F32, ne, a); }
static inline ggml_tensor* ggml_pad(ggml_context* c, ggml_tensor* a, int p0, int p1, int p2, int p3) {
    const int64_t ne[4] = {a->ne[0] + p0, a->ne[1] + p1, a->ne[2] + p2, a->ne[3] + p3};
    return ggml_shim_op(c, GGML_OP_PAD, a->type, ne, a);
}

// ---------------------------------------------------------------------------------------------------------------
// execution
// ---------------------------------------------------------------------------------------------------------------
namespace ggml_shim {

inline int g_threads = 1;

inline char* at(const ggml_tensor* t, int64_t i0, int64_t i1, int64_t i2, int64_t i3) {
    return (char*)t->data + i0 * (int64_t)t->nb[0] + i1 * (int64_t)t->nb[1] + i2 * (int64_t)t->nb[2] + i3 * (int64_t)t->nb[3];
}
inline float load(const ggml_tensor* t, const char* p) {
    switch (t->type) {
        case GGML_TYPE_F32: { float v; memcpy(&v, p, 4); return v; }
        case GGML_TYPE_F16: { uint16_t u; memcpy(&u, p, 2); return ggml_shim_f16_to_f32(u); }
        case GGML_TYPE_BF16: { uint16_t u; memcpy(&u, p, 2); return ggml_shim_bf16_to_f32(u); }
        case GGML_TYPE_I32: { int32_t v; memcpy(&v, p, 4); return (float)v; }
        default: abort();
    }
}
inline void store(const ggml_tensor* t, char* p, float v) {
    switch (t->type) {
        case GGML_TYPE_F32: memcpy(p, &v, 4); break;
        case GGML_TYPE_F16: { uint16_t u = ggml_shim_f32_to_f16(v); memcpy(p, &u, 2); break; }
        case GGML_TYPE_BF16: { uint16_t u = ggml_shim_f32_to_bf16(v); memcpy(p, &u, 2); break; }
        case GGML_TYPE_I32: { int32_t i = (int32_t)v; memcpy(p, &i, 4); break; }
        default: abort();
    }
}
// linear element index (ne[0] fastest) -> address
inline char* at_lin(const ggml_tensor* t, int64_t i) {
    const int64_t i0 = i % t->ne[0]; i /= t->ne[0];
    const int64_t i1 = i % t->ne[1]; i /= t->ne[1];
    const int64_t i2 = i % t->ne[2]; const int64_t i3 = i / t->ne[2];
    return at(t, i0, i1, i2, i3);
}

template <typename F> inline void unary(ggml_tensor* d, F f) {
    const ggml_tensor* a = d->src[0];
    assert(a->type == GGML_TYPE_F32 && d->type == GGML_TYPE_F32);
    const int64_t n = ggml_nelements(d);
    if (ggml_is_contiguous(a) && ggml_is_contiguous(d)) {
        const float* x = (const float*)a->data; float* y = (float*)d->data;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (n > 32768)
        for (int64_t i = 0; i < n; i++) y[i] = f(x[i]);
        return;
    }
    for (int64_t i = 0; i < n; i++) { float v = load(a, at_lin(a, i)); store(d, at_lin(d, i), f(v)); }
}
template <typename F> inline void binary(ggml_tensor* d, F f) {
    const ggml_tensor* a = d->src[0]; const ggml_tensor* b = d->src[1];
    assert(a->type == GGML_TYPE_F32 && b->type == GGML_TYPE_F32 && d->type == GGML_TYPE_F32);
    const bool rows_contig = a->nb[0] == 4 && d->nb[0] == 4 && b->nb[0] == 4;
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) {
        const int64_t j1 = i1 % b->ne[1], j2 = i2 % b->ne[2], j3 = i3 % b->ne[3];
        if (rows_contig && (b->ne[0] == d->ne[0] || b->ne[0] == 1)) {
            const float* x = (const float*)at(a, 0, i1, i2, i3); const float* y = (const float*)at(b, 0, j1, j2, j3); float* o = (float*)at(d, 0, i1, i2, i3);
            if (b->ne[0] == 1) { const float yv = y[0]; for (int64_t i0 = 0; i0 < d->ne[0]; i0++) o[i0] = f(x[i0], yv); }
            else for (int64_t i0 = 0; i0 < d->ne[0]; i0++) o[i0] = f(x[i0], y[i0]);
            continue;
        }
        for (int64_t i0 = 0; i0 < d->ne[0]; i0++) {
            const float x = *(const float*)at(a, i0, i1, i2, i3), y = *(const float*)at(b, i0 % b->ne[0], j1, j2, j3);
            *(float*)at(d, i0, i1, i2, i3) = f(x, y);
        }
    }
}

// ggml_gelu_f32 on the CPU backend: table lookup indexed by the f16 bit pattern of x, values stored as f16
inline float gelu_ggml(float x) {
    if (x <= -10.0f) return 0.0f;
    if (x >= 10.0f) return x;
    const float xf = ggml_shim_f16_to_f32(ggml_shim_f32_to_f16(x));
    const float g = 0.5f * xf * (1.0f + tanhf(0.79788456080286535587989211986876f * xf * (1.0f + 0.044715f * xf * xf)));
    return ggml_shim_f16_to_f32(ggml_shim_f32_to_f16(g));
}

inline void copy_convert(const ggml_tensor* a, ggml_tensor* d) {         // same element count, element order ne[0]-fastest on both sides
    const int64_t n = ggml_nelements(a);
    if (a->type == d->type && ggml_is_contiguous(a) && ggml_is_contiguous(d)) { memcpy(d->data, a->data, ggml_nbytes(a)); return; }
    const bool same_shape = a->ne[0] == d->ne[0] && a->ne[1] == d->ne[1] && a->ne[2] == d->ne[2] && a->ne[3] == d->ne[3];
    if (same_shape && a->nb[0] == ggml_type_size(a->type) && d->nb[0] == ggml_type_size(d->type)) {
        for (int64_t i3 = 0; i3 < a->ne[3]; i3++) for (int64_t i2 = 0; i2 < a->ne[2]; i2++) for (int64_t i1 = 0; i1 < a->ne[1]; i1++) {
            const char* s = at(a, 0, i1, i2, i3); char* o = at(d, 0, i1, i2, i3);
            if (a->type == d->type) memcpy(o, s, (size_t)a->ne[0] * a->nb[0]);
            else for (int64_t i0 = 0; i0 < a->ne[0]; i0++) store(d, o + i0 * d->nb[0], load(a, s + i0 * a->nb[0]));
        }
        return;
    }
    if (same_shape) {                                   // e.g. cont(transpose / permute): strided source, no index arithmetic per element
        const size_t es = ggml_type_size(a->type);
        for (int64_t i3 = 0; i3 < a->ne[3]; i3++) for (int64_t i2 = 0; i2 < a->ne[2]; i2++) for (int64_t i1 = 0; i1 < a->ne[1]; i1++) {
            const char* sp = at(a, 0, i1, i2, i3); char* dp = at(d, 0, i1, i2, i3);
            if (a->type == d->type && es == 4) for (int64_t i0 = 0; i0 < a->ne[0]; i0++) *(uint32_t*)(dp + i0 * d->nb[0]) = *(const uint32_t*)(sp + i0 * a->nb[0]);
            else if (a->type == d->type && es == 2) for (int64_t i0 = 0; i0 < a->ne[0]; i0++) *(uint16_t*)(dp + i0 * d->nb[0]) = *(const uint16_t*)(sp + i0 * a->nb[0]);
            else for (int64_t i0 = 0; i0 < a->ne[0]; i0++) store(d, dp + i0 * d->nb[0], load(a, sp + i0 * a->nb[0]));
        }
        return;
    }
    for (int64_t i = 0; i < n; i++) store(d, at_lin(d, i), load(a, at_lin(a, i)));
}

// mul_mat: d[i0 = row of a, i1 = row of b] = sum_k a[k, i0] * b[k, i1]; the f32 operand b is rounded to a's type when a is F16 / BF16
inline void mul_mat(ggml_tensor* d) {
    const ggml_tensor* a = d->src[0]; const ggml_tensor* b = d->src[1];
    const int64_t K = a->ne[0], N = a->ne[1], M = b->ne[1];
    assert(a->nb[0] == ggml_type_size(a->type));
    const int64_t r2 = b->ne[2] / a->ne[2], r3 = b->ne[3] / a->ne[3];
    std::vector<float> brow((size_t)M * K), arow;
    for (int64_t i3 = 0; i3 < b->ne[3]; i3++) for (int64_t i2 = 0; i2 < b->ne[2]; i2++) {
        // gather + round the activation rows of this (i2, i3) plane
#pragma omp parallel for schedule(static) num_threads(g_threads) if (M * K > 65536)
        for (int64_t m = 0; m < M; m++) {
            float* dst = &brow[(size_t)m * K];
            if (b->type == GGML_TYPE_F32 && b->nb[0] == 4) memcpy(dst, at(b, 0, m, i2, i3), (size_t)K * 4);
            else for (int64_t k = 0; k < K; k++) dst[k] = load(b, at(b, k, m, i2, i3));
            if (a->type == GGML_TYPE_F16) for (int64_t k = 0; k < K; k++) dst[k] = (float)(_Float16)dst[k];
            else if (a->type == GGML_TYPE_BF16) for (int64_t k = 0; k < K; k++) dst[k] = ggml_shim_bf16_to_f32(ggml_shim_f32_to_bf16(dst[k]));
        }
        const int64_t a2 = i2 / r2, a3 = i3 / r3;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (N * M * K > 65536)
        for (int64_t n = 0; n < N; n++) {
            std::vector<float> w((size_t)K);
            const char* ap = at(a, 0, n, a2, a3);
            if (a->type == GGML_TYPE_F32) memcpy(w.data(), ap, (size_t)K * 4);
            else if (a->type == GGML_TYPE_BF16) { const uint16_t* p16 = (const uint16_t*)ap; for (int64_t k = 0; k < K; k++) w[k] = ggml_shim_bf16_to_f32(p16[k]); }
            else if (a->type == GGML_TYPE_F16) { const _Float16* p16 = (const _Float16*)ap; for (int64_t k = 0; k < K; k++) w[k] = (float)p16[k]; }
            else for (int64_t k = 0; k < K; k++) w[k] = load(a, ap + k * a->nb[0]);
            for (int64_t m = 0; m < M; m++) {
                const float* x = &brow[(size_t)m * K];
                float acc = 0.f;
#pragma omp simd reduction(+ : acc)
                for (int64_t k = 0; k < K; k++) acc += w[k] * x[k];
                *(float*)at(d, n, m, i2, i3) = acc;
            }
        }
    }
}

inline void soft_max(ggml_tensor* d) {
    const ggml_tensor* a = d->src[0]; const ggml_tensor* mask = d->src[1];
    const float scale = ggml_shim_getf(d, 0);
    const int64_t n = a->ne[0];
    for (int64_t i3 = 0; i3 < a->ne[3]; i3++) for (int64_t i2 = 0; i2 < a->ne[2]; i2++) for (int64_t i1 = 0; i1 < a->ne[1]; i1++) {
        std::vector<float> w((size_t)n);
        float mx = -INFINITY;
        for (int64_t i = 0; i < n; i++) {
            float v = *(const float*)at(a, i, i1, i2, i3) * scale;
            if (mask) v += load(mask, at(mask, i, i1, i2 % mask->ne[2], i3 % mask->ne[3]));
            w[i] = v; mx = std::max(mx, v);
        }
        double sum = 0.0;
        for (int64_t i = 0; i < n; i++) { const float e = expf(w[i] - mx); w[i] = e; sum += (double)e; }
        const float inv = (float)(1.0 / sum);
        for (int64_t i = 0; i < n; i++) *(float*)at(d, i, i1, i2, i3) = w[i] * inv;
    }
}

inline void compute(ggml_tensor* d) {
    const ggml_tensor* a = d->src[0]; const ggml_tensor* b = d->src[1];
    assert(d->data && "graph node without storage");
    switch (d->op) {
        case GGML_OP_VIEW: case GGML_OP_RESHAPE: case GGML_OP_PERMUTE: case GGML_OP_TRANSPOSE: case GGML_OP_NONE: break;
        case GGML_OP_ADD: binary(d, [](float x, float y) { return x + y; }); break;
        case GGML_OP_SUB: binary(d, [](float x, float y) { return x - y; }); break;
        case GGML_OP_MUL: binary(d, [](float x, float y) { return x * y; }); break;
        case GGML_OP_DIV: binary(d, [](float x, float y) { return x / y; }); break;
        case GGML_OP_SCALE: { const float s = ggml_shim_getf(d, 0); unary(d, [s](float x) { return x * s; }); break; }
        case GGML_OP_NEG: unary(d, [](float x) { return -x; }); break;
        case GGML_OP_EXP: unary(d, [](float x) { return expf(x); }); break;
        case GGML_OP_SIN: unary(d, [](float x) { return sinf(x); }); break;
        case GGML_OP_COS: unary(d, [](float x) { return cosf(x); }); break;
        case GGML_OP_SQRT: unary(d, [](float x) { return sqrtf(x); }); break;
        case GGML_OP_SILU: unary(d, [](float x) { return x / (1.0f + expf(-x)); }); break;
        case GGML_OP_GELU: unary(d, [](float x) { return gelu_ggml(x); }); break;
        case GGML_OP_ELU: unary(d, [](float x) { return x > 0.f ? x : expm1f(x); }); break;
        case GGML_OP_CLAMP: { const float lo = ggml_shim_getf(d, 0), hi = ggml_shim_getf(d, 1); unary(d, [lo, hi](float x) { return x < lo ? lo : (x > hi ? hi : x); }); break; }
        case GGML_OP_NORM: case GGML_OP_RMS_NORM: {
            const float eps = ggml_shim_getf(d, 0);
            assert(a->type == GGML_TYPE_F32 && a->nb[0] == 4 && d->nb[0] == 4);
            for (int64_t i3 = 0; i3 < a->ne[3]; i3++) for (int64_t i2 = 0; i2 < a->ne[2]; i2++) for (int64_t i1 = 0; i1 < a->ne[1]; i1++) {
                const float* x = (const float*)at(a, 0, i1, i2, i3); float* y = (float*)at(d, 0, i1, i2, i3);
                const int64_t n = a->ne[0];
                if (d->op == GGML_OP_NORM) {
                    double sum = 0.0; for (int64_t i = 0; i < n; i++) sum += (double)x[i];
                    const float mean = (float)(sum / n);
                    double sum2 = 0.0;
                    for (int64_t i = 0; i < n; i++) { const float v = x[i] - mean; y[i] = v; sum2 += (double)(v * v); }
                    const float sc = 1.0f / sqrtf((float)(sum2 / n) + eps);
                    for (int64_t i = 0; i < n; i++) y[i] *= sc;
                } else {
                    double sum = 0.0; for (int64_t i = 0; i < n; i++) sum += (double)(x[i] * x[i]);
                    const float sc = 1.0f / sqrtf((float)(sum / n) + eps);
                    for (int64_t i = 0; i < n; i++) y[i] = x[i] * sc;
                }
            }
            break;
        }
        case GGML_OP_MUL_MAT: mul_mat(d); break;
        case GGML_OP_SOFT_MAX: soft_max(d); break;
        case GGML_OP_CPY: case GGML_OP_CONT: case GGML_OP_DUP: copy_convert(a, d); break;
        case GGML_OP_CONCAT: {
            const int dim = d->op_params[0];
            if (dim == 0 && a->nb[0] == ggml_type_size(a->type) && b->nb[0] == a->nb[0] && d->nb[0] == a->nb[0]) {      // rows are runs of bytes
                for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) {
                    memcpy(at(d, 0, i1, i2, i3), at(a, 0, i1, i2, i3), (size_t)a->ne[0] * a->nb[0]);
                    memcpy(at(d, a->ne[0], i1, i2, i3), at(b, 0, i1, i2, i3), (size_t)b->ne[0] * b->nb[0]);
                }
                break;
            }
            for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) for (int64_t i0 = 0; i0 < d->ne[0]; i0++) {
                int64_t idx[4] = {i0, i1, i2, i3};
                const ggml_tensor* s = a;
                if (idx[dim] >= a->ne[dim]) { s = b; idx[dim] -= a->ne[dim]; }
                memcpy(at(d, i0, i1, i2, i3), at(s, idx[0], idx[1], idx[2], idx[3]), ggml_type_size(d->type));
            }
            break;
        }
        case GGML_OP_REPEAT:
            for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) for (int64_t i0 = 0; i0 < d->ne[0]; i0++)
                memcpy(at(d, i0, i1, i2, i3), at(a, i0 % a->ne[0], i1 % a->ne[1], i2 % a->ne[2], i3 % a->ne[3]), ggml_type_size(d->type));
            break;
        case GGML_OP_IM2COL: {      // a = kernel [K, IC, OC], b = input [L, IC, N] (F32) -> d [IC*K, OL, N] F16
            const int s0 = d->op_params[0], p0 = d->op_params[1], d0 = d->op_params[2];
            const int64_t K = a->ne[0], IC = a->ne[1], L = b->ne[0], OL = d->ne[1], N = d->ne[2];
            assert(b->type == GGML_TYPE_F32);
            for (int64_t n = 0; n < N; n++) {
#pragma omp parallel for schedule(static) num_threads(g_threads) if (OL * IC * K > 32768)
                for (int64_t ol = 0; ol < OL; ol++) {
                    _Float16* o = (_Float16*)at(d, 0, ol, n, 0);
                    for (int64_t ic = 0; ic < IC; ic++) for (int64_t k = 0; k < K; k++) {
                        const int64_t il = ol * s0 + k * d0 - p0;
                        o[ic * K + k] = (_Float16)((il < 0 || il >= L) ? 0.f : *(const float*)at(b, il, ic, n, 0));
                    }
                }
            }
            break;
        }
        case GGML_OP_CONV_TRANSPOSE_1D: {   // a = kernel [K, OC, IC] (F16 | F32), b = input [L, IC] F32 -> d [(L-1)*s0 + K, OC] F32
            const int s0 = d->op_params[0];
            const int64_t K = a->ne[0], OC = a->ne[1], IC = a->ne[2], L = b->ne[0];
            const int64_t n_out = ggml_nelements(d);
            for (int64_t i = 0; i < n_out; i++) *(float*)at_lin(d, i) = 0.f;
            // kernel -> [oc][k][ic], source -> [l][ic] (ggml's wdata permutation). The permuted kernel is cached per weight tensor: weights
            // are leaves that never change after loading (keyed by data pointer + shape).
            static std::vector<std::pair<const void*, std::vector<float>>> wcache;
            std::vector<float>* wp = nullptr;
            for (auto& e : wcache) if (e.first == a->data && e.second.size() == (size_t)K * OC * IC) wp = &e.second;
            if (!wp) {
                wcache.emplace_back(a->data, std::vector<float>((size_t)K * OC * IC));
                wp = &wcache.back().second;
                for (int64_t ic = 0; ic < IC; ic++) for (int64_t oc = 0; oc < OC; oc++) for (int64_t k = 0; k < K; k++) (*wp)[((size_t)oc * K + k) * IC + ic] = load(a, at(a, k, oc, ic, 0));
            }
            const std::vector<float>& w = *wp;
            std::vector<float> x((size_t)L * IC);
            for (int64_t ic = 0; ic < IC; ic++) for (int64_t l = 0; l < L; l++) x[(size_t)l * IC + ic] = load(b, at(b, l, ic, 0, 0));
#pragma omp parallel for schedule(static) num_threads(g_threads)
            for (int64_t oc = 0; oc < OC; oc++)
                for (int64_t l = 0; l < L; l++) for (int64_t k = 0; k < K; k++) {
                    const float* wv = &w[((size_t)oc * K + k) * IC]; const float* xv = &x[(size_t)l * IC];
                    float acc = 0.f;
#pragma omp simd reduction(+ : acc)
                    for (int64_t ic = 0; ic < IC; ic++) acc += xv[ic] * wv[ic];
                    *(float*)at(d, l * s0 + k, oc, 0, 0) += acc;
                }
            break;
        }
        case GGML_OP_GET_ROWS:
            for (int64_t i2 = 0; i2 < b->ne[2]; i2++) for (int64_t i1 = 0; i1 < b->ne[1]; i1++) for (int64_t i = 0; i < b->ne[0]; i++) {
                const int32_t r = *(const int32_t*)at(b, i, i1, i2, 0);
                assert(r >= 0 && r < a->ne[1]);
                for (int64_t k = 0; k < a->ne[0]; k++) *(float*)at(d, k, i, i1, i2) = load(a, at(a, k, r, i1, i2));
            }
            break;
        case GGML_OP_SET_ROWS: {            // src[0] = rows (F32), src[1] = indices, src[2] = destination (d is a view of it)
            const ggml_tensor* rows = d->src[0]; const ggml_tensor* idx = d->src[1];
            for (int64_t i3 = 0; i3 < rows->ne[3]; i3++) for (int64_t i2 = 0; i2 < rows->ne[2]; i2++) for (int64_t i = 0; i < rows->ne[1]; i++) {
                const char* ip = at(idx, i, i2 % idx->ne[1], i3 % idx->ne[2], 0);
                const int64_t r = idx->type == GGML_TYPE_I64 ? *(const int64_t*)ip : (int64_t)*(const int32_t*)ip;
                assert(r >= 0 && r < d->ne[1]);
                for (int64_t k = 0; k < rows->ne[0]; k++) store(d, at(d, k, r, i2, i3), *(const float*)at(rows, k, i, i2, i3));
            }
            break;
        }
        case GGML_OP_TIMESTEP_EMBEDDING: {
            const int dim = d->op_params[0], max_period = d->op_params[1], half = dim / 2;
            for (int64_t i = 0; i < a->ne[0]; i++) {
                const float t = *(const float*)at(a, i, 0, 0, 0);
                float* e = (float*)at(d, 0, i, 0, 0);
                for (int j = 0; j < half; j++) {
                    const float freq = (float)expf(-logf((float)max_period) * j / half);
                    const float arg = t * freq;
                    e[j] = cosf(arg); e[j + half] = sinf(arg);
                }
                if (dim % 2) e[2 * half] = 0.f;
            }
            break;
        }
        case GGML_OP_ARANGE: {
            const float start = ggml_shim_getf(d, 0), step = ggml_shim_getf(d, 2);
            for (int64_t i = 0; i < d->ne[0]; i++) ((float*)d->data)[i] = start + step * i;
            break;
        }
        case GGML_OP_SUM: {                 // ggml_vec_sum_f32_ggf: double accumulation over every element
            double s = 0.0; const int64_t n = ggml_nelements(a);
            for (int64_t i = 0; i < n; i++) s += (double)load(a, at_lin(a, i));
            *(float*)d->data = (float)s;
            break;
        }
        case GGML_OP_MEAN:
            for (int64_t i3 = 0; i3 < a->ne[3]; i3++) for (int64_t i2 = 0; i2 < a->ne[2]; i2++) for (int64_t i1 = 0; i1 < a->ne[1]; i1++) {
                float s = 0.f; for (int64_t i0 = 0; i0 < a->ne[0]; i0++) s += *(const float*)at(a, i0, i1, i2, i3);        // ggml_vec_sum_f32 (float), then / ne00
                *(float*)at(d, 0, i1, i2, i3) = s / (float)a->ne[0];
            }
            break;
        case GGML_OP_PAD:
            for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) for (int64_t i0 = 0; i0 < d->ne[0]; i0++) {
                const bool in = i0 < a->ne[0] && i1 < a->ne[1] && i2 < a->ne[2] && i3 < a->ne[3];
                *(float*)at(d, i0, i1, i2, i3) = in ? *(const float*)at(a, i0, i1, i2, i3) : 0.f;
            }
            break;
        default: fprintf(stderr, "ggml shim: op %d not implemented\n", (int)d->op); abort();
    }
}

}  // namespace ggml_shim

enum ggml_status { GGML_STATUS_SUCCESS = 0 };
static inline ggml_status ggml_backend_graph_compute(ggml_backend_t backend, ggml_cgraph* g) {
    ggml_shim::g_threads = backend ? backend->n_threads : 1;
    static const bool prof = getenv("GGML_SHIM_PROFILE") != nullptr;       // per-op wall time, printed at exit (development aid)
    if (!prof) { for (auto* t : g->nodes) ggml_shim::compute(t); return GGML_STATUS_SUCCESS; }
    static double acc[64] = {0}; static long cnt[64] = {0}; static bool reg = false;
    if (!reg) { reg = true; atexit([] { for (int i = 0; i < 64; i++) if (cnt[i]) fprintf(stderr, "ggml shim op %2d: %8ld calls %9.1f ms\n", i, cnt[i], acc[i] * 1e3); }); }
    for (auto* t : g->nodes) {
        struct timespec a, b; clock_gettime(CLOCK_MONOTONIC, &a);
        ggml_shim::compute(t);
        clock_gettime(CLOCK_MONOTONIC, &b);
        acc[t->op] += (b.tv_sec - a.tv_sec) + (b.tv_nsec - a.tv_nsec) * 1e-9; cnt[t->op]++;
    }
    return GGML_STATUS_SUCCESS;
}
