// Stand-in for gguf.h (oracle/_ref build, test infrastructure). The reference's GGUF path (src/loader.h:78-92,228-272) is dormant:
// nothing on the generation path calls it (weights come from safetensors), so these only have to link; calling one aborts.
#pragma once
#include "ggml.h"
struct gguf_context;
struct gguf_init_params { bool no_alloc; ggml_context** ctx; };
#define GGUF_SHIM_DEAD(sig) static inline sig { fprintf(stderr, "ggml shim: GGUF is not supported (%s)\n", __func__); abort(); }
GGUF_SHIM_DEAD(gguf_context* gguf_init_from_file(const char*, gguf_init_params))
GGUF_SHIM_DEAD(gguf_context* gguf_init_empty(void))
static inline void gguf_free(gguf_context*) {}
GGUF_SHIM_DEAD(void gguf_add_tensor(gguf_context*, const ggml_tensor*))
GGUF_SHIM_DEAD(bool gguf_write_to_file(const gguf_context*, const char*, bool))
GGUF_SHIM_DEAD(size_t gguf_get_data_offset(const gguf_context*))
GGUF_SHIM_DEAD(int64_t gguf_get_n_tensors(const gguf_context*))
GGUF_SHIM_DEAD(const char* gguf_get_tensor_name(const gguf_context*, int64_t))
GGUF_SHIM_DEAD(size_t gguf_get_tensor_offset(const gguf_context*, int64_t))
GGUF_SHIM_DEAD(size_t gguf_get_tensor_size(const gguf_context*, int64_t))
