// Forced-include for the oracle/_ref build (TEST INFRASTRUCTURE): makes the reference's noise source injectable WITHOUT touching its
// sources. src/context.h:8-12 declares `std::normal_distribution nd{0.0, 1.0};` and draws `(float)nd(gen) * std + mean` per noise
// element (src/context.h:501-503). With -Dnormal_distribution=ptts_ref_injectable_normal that global becomes the class below: draws
// come from a queue filled by ref_inject_noise() (0.0 when the queue is empty, e.g. the draws a prefill wastes), so a stream created
// with temp = 1 (std = 1) receives exactly the injected values — "identical injected noise" (BASELINE north_star) for the reference.
#pragma once
#include <deque>
#include <random>
namespace ptts_ref_noise { inline std::deque<double>& queue() { static std::deque<double> q; return q; } }
namespace std {
struct ptts_ref_injectable_normal {
    ptts_ref_injectable_normal(double, double) {}
    template <class G> double operator()(G&) {
        auto& q = ptts_ref_noise::queue();
        if (q.empty()) return 0.0;
        const double v = q.front(); q.pop_front(); return v;
    }
};
}  // namespace std
#define normal_distribution ptts_ref_injectable_normal
