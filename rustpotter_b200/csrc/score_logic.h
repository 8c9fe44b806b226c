// Window judgement shared by the device kernel (K3) and the host replay hook — one source, two
// compilers. Restates WakewordComparator::run_detection after the per-template scores exist
// (reference src/wakewords/comp/wakeword_comp.rs:83-151) and the best-wakeword pick of
// Rustpotter::run_wakeword_detectors (src/detector.rs:433-447).
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define RP_HD __host__ __device__ __forceinline__
#else
#define RP_HD inline
#endif

namespace rp {

constexpr int kJudgeMaxTemplates = 64;

// One per loaded wakeword; `slot_begin` indexes the per-window score row:
//   [avg score (only if has_avg)] [template 0] ... [template T-1]
struct WakewordMeta {
    int32_t slot_begin;
    int32_t n_templates;
    int32_t has_avg;        // avg_features.is_some()
    float threshold;        // wakeword.threshold.unwrap_or(config threshold)      (wakeword_comp.rs:95)
    float avg_threshold;    // wakeword.avg_threshold.unwrap_or(config avg_thr)    (wakeword_comp.rs:83)
};

// get_percentile (wakeword_comp.rs:38-49) on an ascending-sorted slice
RP_HD float percentile(const float* sorted, int n, float pct) {
    float index = pct / 100.0f * (float)(n - 1);
    float index_floor = floorf(index);
    if (index_floor == index) return sorted[(int)index];
    int i = (int)index_floor;
    float d = index - index_floor;
    return sorted[i] * (1.0f - d) + sorted[i + 1] * d;
}

// Score aggregation (wakeword_comp.rs:108-139). `v` is scratch and is reordered.
RP_HD float aggregate_scores(float* v, int n, int mode) {
    if (mode == 0 /* Average */) {
        float s = 0.f;
        for (int i = 0; i < n; i++) s += v[i];
        return s / (float)n;
    }
    if (mode == 1 /* Max */) {
        float m = v[0];
        for (int i = 1; i < n; i++) m = v[i] > m ? v[i] : m;
        return m;
    }
    for (int i = 1; i < n; i++) {  // insertion sort ascending (n <= 64, usually 3..8)
        float x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; j--; }
        v[j + 1] = x;
    }
    float pct = (mode == 2 || mode == 4) ? 50.f : mode == 3 ? 25.f : mode == 5 ? 75.f : mode == 6 ? 80.f : mode == 7 ? 90.f : 95.f;
    return percentile(v, n, pct);
}

struct Judgement {
    int wakeword;  // -1: no wakeword detected on this window
    float avg_score;
    float score;
};

// row: the window's score row (all slots). Picks the detection with the highest score; on a tie
// the earlier wakeword wins (the reference's order is that of an unordered HashMap).
RP_HD Judgement judge_window(const float* row, const WakewordMeta* metas, int n_wakewords, int score_mode) {
    Judgement best{-1, 0.f, 0.f};
    float scratch[kJudgeMaxTemplates];
    for (int w = 0; w < n_wakewords; w++) {
        const WakewordMeta m = metas[w];
        float avg_score = 0.f;
        int s0 = m.slot_begin;
        if (m.has_avg) {
            if (m.avg_threshold != 0.f) {
                avg_score = row[s0];
                if (avg_score < m.avg_threshold) continue;  // wakeword_comp.rs:91-93
            }
            s0 += 1;
        }
        for (int t = 0; t < m.n_templates; t++) scratch[t] = row[s0 + t];
        float score = aggregate_scores(scratch, m.n_templates, score_mode);
        if (score > m.threshold && (best.wakeword < 0 || score > best.score)) best = Judgement{w, avg_score, score};
    }
    return best;
}

}  // namespace rp
