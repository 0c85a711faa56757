// pocket_tts.h — the reference's public streaming API (Codes4Fun/pocket-tts.cpp include/pocket_tts/pocket_tts.h:18-42)
// re-declared for the B200 engine with identical names and C++ signatures, so callers written against the
// reference (demos/pocket-tts.cpp) compile and link against libptts_b200.so unchanged. The reference header pulls
// in <ggml.h> only for the `ggml_backend*` parameter type; when ggml is not installed the forward declaration
// below yields the same mangled symbols. The backend pointers are ignored by this implementation.
#pragma once

#if defined(__has_include)
#  if __has_include(<ggml-backend.h>)
#    include <ggml.h>
#    include <ggml-backend.h>
#    define PTTS_B200_HAVE_GGML 1
#  endif
#endif
#ifndef PTTS_B200_HAVE_GGML
struct ggml_backend;
#endif

#define PTTS_API __attribute__ ((visibility ("default"))) extern

PTTS_API void ptts_set_seed( unsigned int seed );
PTTS_API unsigned int ptts_get_seed();

struct ptts_context_t;

PTTS_API ptts_context_t * ptts_init(
    ggml_backend * backend,
    ggml_backend * backend_cpu,
    const char * model_path
);
PTTS_API int ptts_get_sample_rate( ptts_context_t * ptts_ctx );
PTTS_API int ptts_get_frame_size( ptts_context_t * ptts_ctx );

struct ptts_stream_t;

PTTS_API ptts_stream_t * ptts_stream_from_safetensors(
    ptts_context_t * ptts_ctx,
    const char * voice,
    float temp = 0.7f
);
PTTS_API void ptts_stream_reset( ptts_stream_t * stream );
PTTS_API void ptts_stream_flush( ptts_stream_t * stream );
PTTS_API void ptts_stream_send( ptts_stream_t * stream, const char * chunk );
PTTS_API bool ptts_stream_receive( ptts_stream_t * stream, float * samples );
