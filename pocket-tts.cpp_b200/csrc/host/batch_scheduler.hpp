// Continuous batching over the engine's slots.
//
// The reference serves ONE stream and rolls to the next sentence inside ptts_stream_receive (src/pocket_tts.cpp:494-519): every
// sentence restarts from the voice-conditioned state and resets the Mimi decoder (_stream_sentence_init :416-444), so sentences are
// independent units of work even inside one utterance. This scheduler exploits that: sentences ("jobs") of many utterances are queued,
// longest first; every step covers the slot range in use; a slot whose sentence finished (EOS rule or cap, decided on the device) is
// refilled with the next queued sentence while the other slots keep generating. Nothing here synchronises the device except the
// collect of the oldest in-flight frame: sentence starts are enqueued between steps (b200_begin_sentences_ex is asynchronous).
//
// The engine is reached through a table of three calls so that the scheduling logic can be driven by a mock engine in CPU tests.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

namespace ptts_host {

struct BatchOps {
    void* user = nullptr;
    // per-sentence arrays of length n; tokens concatenated, tok_off[n+1]; returns 0 on success
    int (*begin)(void* user, int n, const int32_t* slots, const int32_t* voices, const int32_t* tokens, const int32_t* tok_off,
                 const int32_t* max_gen_len, const int32_t* frames_after_eos, const float* temp, const uint32_t* rng_stream) = nullptr;
    int (*submit)(void* user, int slot0, int n) = nullptr;                           // enqueue one step of slots [slot0, slot0 + n)
    int (*collect)(void* user, float* pcm, int32_t* produced) = nullptr;              // oldest in-flight step: [n][1920], [n]; returns n
};

struct BatchStats {
    long long steps = 0;          // engine steps submitted
    long long frames = 0;         // audio frames produced (all utterances)
    long long slot_steps = 0;     // sum over steps of the stepped slot count (idle fraction = 1 - frames / slot_steps)
    long long sentences = 0;      // sentences completed
    long long refills = 0;        // sentence-start calls
    double wall_ms = 0.0;         // run() wall time
    double begin_ms = 0.0, submit_ms = 0.0, collect_ms = 0.0;   // host time inside the three engine calls (collect includes waiting for the GPU)
};

class BatchScheduler {
public:
    struct Job {
        int utt = 0, index = 0;              // utterance id, sentence index inside it
        int voice = 0; float temp = 0.7f;
        std::vector<int32_t> ids;            // SentencePiece ids
        int max_gen = 0, fae = 0;            // reference stop rule: cap int((words + 2) * 12.5), frames_after_eos (src/pocket_tts.cpp:429-430,504-506)
        uint32_t rng_stream = 0;
        std::vector<float> pcm; int frames = 0; bool done = false;
    };

    BatchScheduler(const BatchOps& ops, int n_slots, int frame = 1920) : ops_(ops), n_slots_(n_slots), frame_(frame), slots_(n_slots) {}

    // Tunables: refill as soon as `refill_min` slots are free or `refill_every` steps passed since the last sentence start (each
    // sentence start costs a prefill pass over the weights, so single-slot refills every step would not pay); range granularity.
    int refill_min = 8, refill_every = 8, range_quantum = 32, depth = 2;
    bool keep_pcm = true;

    int add_job(Job j) { j.rng_stream = j.rng_stream ? j.rng_stream : (uint32_t)jobs_.size() + 1; jobs_.push_back(std::move(j)); return (int)jobs_.size() - 1; }
    const std::vector<Job>& jobs() const { return jobs_; }
    const BatchStats& stats() const { return stats_; }
    // effective cap of a job after the engine's KV-capacity clamp (host copy of the rule in b200_begin_sentences): set by the owner
    std::vector<int> room_of_voice;          // kv_capacity - voice_len[v]; empty = no clamp known
    int effective_cap(const Job& j) const {
        int cap = j.max_gen;
        if (j.voice >= 0 && j.voice < (int)room_of_voice.size()) cap = std::min(cap, room_of_voice[j.voice] - (int)j.ids.size());
        return std::max(cap, 0);
    }

    // Runs every queued job to completion. Returns the number of frames produced, or a negative engine error.
    long long run() {
        using clk = std::chrono::steady_clock;
        const auto t0 = clk::now();
        std::vector<int> order;
        for (int i = 0; i < (int)jobs_.size(); i++) if (!jobs_[i].done && !started_[i]) order.push_back(i);
        // longest processing time first (by the frame cap): the long sentences start early, the tail of the run is made of short ones
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return jobs_[a].max_gen > jobs_[b].max_gen; });
        std::deque<int> pending(order.begin(), order.end());
        std::vector<float> pcm((size_t)n_slots_ * frame_); std::vector<int32_t> produced(n_slots_);
        std::deque<int> inflight_n;                              // stepped slot count of each in-flight step, oldest first
        long long submit_idx = 0, collect_idx = 0; int since_refill = 1 << 30;
        const long long frames0 = stats_.frames;
        for (;;) {
            int busy = 0, top = 0, n_free = 0;
            for (int s = 0; s < n_slots_; s++) { if (slots_[s].job >= 0) { busy++; top = s + 1; } else n_free++; }
            // ---- refill free slots ----
            if (!pending.empty() && n_free > 0 && (busy == 0 || n_free >= std::min(refill_min, (int)pending.size()) || since_refill >= refill_every)) {
                std::vector<int32_t> sl, vo, toks, off{0}, mg, fae; std::vector<float> tp; std::vector<uint32_t> rs;
                for (int s = 0; s < n_slots_ && !pending.empty(); s++) {
                    if (slots_[s].job >= 0) continue;
                    const int j = pending.front(); pending.pop_front();
                    Job& J = jobs_[j];
                    if (effective_cap(J) <= 0) { J.done = true; stats_.sentences++; s--; continue; }   // no room (or empty cap): nothing to generate
                    slots_[s].job = j; slots_[s].live_from = submit_idx; started_[j] = true;
                    // one allocation per sentence, sized by its cap (virtual memory only until frames arrive): growing the buffer frame by
                    // frame re-faulted every page several times and cost more host time per step than the GPU step itself
                    if (keep_pcm) J.pcm.reserve((size_t)effective_cap(J) * frame_);
                    sl.push_back(s); vo.push_back(J.voice); toks.insert(toks.end(), J.ids.begin(), J.ids.end()); off.push_back((int32_t)toks.size());
                    mg.push_back(J.max_gen); fae.push_back(J.fae); tp.push_back(J.temp); rs.push_back(J.rng_stream);
                }
                if (!sl.empty()) {
                    const auto tb = clk::now();
                    const int rc = ops_.begin(ops_.user, (int)sl.size(), sl.data(), vo.data(), toks.data(), off.data(), mg.data(), fae.data(), tp.data(), rs.data());
                    stats_.begin_ms += std::chrono::duration<double, std::milli>(clk::now() - tb).count();
                    if (rc != 0) return rc;
                    stats_.refills++; since_refill = 0;
                }
                busy = 0; top = 0;
                for (int s = 0; s < n_slots_; s++) if (slots_[s].job >= 0) { busy++; top = s + 1; }
            }
            if (busy == 0 && pending.empty() && inflight_n.empty()) break;
            // ---- submit a step while work remains and the pipeline has room ----
            if (busy > 0 && (int)inflight_n.size() < depth) {
                int n = std::min(n_slots_, (top + range_quantum - 1) / range_quantum * range_quantum);
                const auto ts = clk::now();
                const int rc = ops_.submit(ops_.user, 0, n);
                stats_.submit_ms += std::chrono::duration<double, std::milli>(clk::now() - ts).count();
                if (rc != 0) return rc;
                inflight_n.push_back(n); submit_idx++; since_refill++;
                stats_.steps++; stats_.slot_steps += n;
                if ((int)inflight_n.size() < depth) continue;     // fill the pipeline before the first collect
            }
            // ---- collect the oldest step ----
            if (inflight_n.empty()) continue;
            const int n = inflight_n.front(); inflight_n.pop_front();
            const auto tc0 = clk::now();
            const int rc = ops_.collect(ops_.user, pcm.data(), produced.data());
            stats_.collect_ms += std::chrono::duration<double, std::milli>(clk::now() - tc0).count();
            if (rc < 0) return rc;
            for (int s = 0; s < n; s++) {
                Slot& S = slots_[s];
                if (S.job < 0 || collect_idx < S.live_from) continue;          // frame of a previous (finished) sentence in this slot
                Job& J = jobs_[S.job];
                bool finished = !produced[s];
                if (produced[s]) {
                    if (keep_pcm) J.pcm.insert(J.pcm.end(), pcm.begin() + (size_t)s * frame_, pcm.begin() + (size_t)(s + 1) * frame_);
                    J.frames++; stats_.frames++;
                    if (J.frames >= effective_cap(J)) finished = true;           // the device stops it at the cap as well: no need to wait for a 0 flag
                }
                if (finished) { J.done = true; stats_.sentences++; S.job = -1; }
            }
            collect_idx++;
        }
        stats_.wall_ms += std::chrono::duration<double, std::milli>(clk::now() - t0).count();
        return stats_.frames - frames0;
    }

    void on_jobs_added() { started_.resize(jobs_.size(), false); }

private:
    struct Slot { int job = -1; long long live_from = 0; };
    BatchOps ops_; int n_slots_, frame_;
    std::vector<Slot> slots_;
    std::vector<Job> jobs_;
    std::vector<bool> started_;
    BatchStats stats_;
};

}  // namespace ptts_host
