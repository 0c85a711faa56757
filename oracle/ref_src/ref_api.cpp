// ref_api.cpp — C entry points over the REFERENCE'S OWN SOURCES (TEST INFRASTRUCTURE; oracle/_ref/libptts_ref.so).
//
// This translation unit includes /root/reference/src/pocket_tts.cpp where it lies (nothing of the reference is copied into the repo);
// it compiles against oracle/ggml_shim (ggml / gguf stand-ins, a SentencePiece stand-in over the repo's unigram encoder) because the
// reference's real dependencies are un-vendored and absent here. What runs is the reference's graph-building code, state handling and
// driver loop (ptts_init, ptts_stream_from_safetensors, ptts_stream_send/flush/receive); only ggml's op kernels are restated.
// The extra accessors read the reference's internal state for parity checks (KV rows, current_end, the latent) and inject noise.
#include <pocket_tts.cpp>   // -I/root/reference/src

extern "C" {

struct ref_handle { ggml_backend* backend; ggml_backend* backend_cpu; ptts_context_t* ctx; };

void* ref_init(const char* model_path, int threads) {
    auto* h = new ref_handle;
    h->backend = ggml_backend_cpu_init(); h->backend_cpu = ggml_backend_cpu_init();
    ggml_backend_cpu_set_n_threads(h->backend, threads); ggml_backend_cpu_set_n_threads(h->backend_cpu, threads);
    h->ctx = ptts_init(h->backend, h->backend_cpu, model_path);
    return h;
}
void ref_set_threads(void* hv, int threads) { auto* h = (ref_handle*)hv; ggml_backend_cpu_set_n_threads(h->backend, threads); ggml_backend_cpu_set_n_threads(h->backend_cpu, threads); }
void ref_set_seed(unsigned int seed) { ptts_set_seed(seed); }
void* ref_stream(void* hv, const char* voice, float temp) { return ptts_stream_from_safetensors(((ref_handle*)hv)->ctx, voice, temp); }
void ref_stream_reset(void* s) { ptts_stream_reset((ptts_stream_t*)s); }
void ref_send(void* s, const char* chunk) { ptts_stream_send((ptts_stream_t*)s, chunk); }
void ref_flush(void* s) { ptts_stream_flush((ptts_stream_t*)s); }
int ref_receive(void* s, float* samples) { return ptts_stream_receive((ptts_stream_t*)s, samples) ? 1 : 0; }
int ref_sample_rate(void* hv) { return ptts_get_sample_rate(((ref_handle*)hv)->ctx); }
int ref_frame_size(void* hv) { return ptts_get_frame_size(((ref_handle*)hv)->ctx); }

int ref_tokenize(void* hv, const char* text, int* ids, int max_ids) {
    std::vector<int> v;
    conditioner_prepare(((ref_handle*)hv)->ctx->flow_lm->conditioner, text, v);
    for (int i = 0; i < (int)v.size() && i < max_ids; i++) ids[i] = v[i];
    return (int)v.size();
}
int ref_count_words(const char* text) { std::string t = text; return count_words(t); }
int ref_pending_sentences(void* sv, int index, char* buf, int buflen) {
    auto* s = (ptts_stream_t*)sv;
    const int n = (int)s->sproc.sentences.size();
    if (index >= 0 && index < n && buf && buflen > 0) snprintf(buf, buflen, "%s", s->sproc.sentences[index].c_str());
    return n;
}

// ---- internal state, for parity checks ----
int ref_current_end(void* sv) { return ((ptts_stream_t*)sv)->model_states->transformer->layers[0]->self_attn->current_end; }
int ref_generation_step(void* sv) { return ((ptts_stream_t*)sv)->generation_step; }
int ref_max_gen_len(void* sv) { return ((ptts_stream_t*)sv)->max_gen_len; }
int ref_frames_after_eos(void* sv) { return ((ptts_stream_t*)sv)->frames_after_eos; }
int ref_mimi_offset(void* sv) { return ((ptts_stream_t*)sv)->mimi_states->decoder_transformer->offset; }
// K (which = 0) or V (1) cache rows [0, n_pos) of FlowLM layer `layer`: the reference keeps [64, 16, capacity] f32 (modules/transformer.h:21-33)
int ref_read_kv(void* sv, int layer, int which, int n_pos, float* out) {
    auto* st = ((ptts_stream_t*)sv)->model_states->transformer->layers[layer]->self_attn;
    ggml_tensor* t = which ? st->values : st->keys;
    if (n_pos > t->ne[2]) return -1;
    ggml_backend_tensor_get(t, out, 0, (size_t)n_pos * 1024 * sizeof(float));
    return 0;
}
void ref_get_latent(void* sv, float* out32) { ggml_backend_tensor_get(((ptts_stream_t*)sv)->model_states->output, out32, 0, 32 * sizeof(float)); }
// teacher forcing: the next step's backbone input (stream->backbone_input aliases model_states->output after the first step)
void ref_set_latent(void* sv, const float* in32) {
    auto* s = (ptts_stream_t*)sv;
    ggml_backend_tensor_set(s->model_states->output, in32, 0, 32 * sizeof(float));
    s->backbone_input = s->model_states->output;
}
// values the next noise draws return, in order (see inject_normal.h); use a stream with temp = 1 so that they arrive unscaled
void ref_inject_noise(const float* z, int n) { auto& q = ptts_ref_noise::queue(); q.clear(); for (int i = 0; i < n; i++) q.push_back((double)z[i]); }
int ref_noise_pending(void) { return (int)ptts_ref_noise::queue().size(); }

}  // extern "C"
