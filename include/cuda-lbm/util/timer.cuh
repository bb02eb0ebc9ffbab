// util/timer.cuh — wall-clock bookkeeping used by main.cu-style drivers (reference src/util/timer.cuh:11-78).
#pragma once
#include <algorithm>
#include <chrono>
#include <deque>
#include <iomanip>
#include <iostream>
#include <vector>

class Timer {
    using clock = std::chrono::high_resolution_clock;
    clock::time_point t_start, t_last;
    std::deque<float> window;              // the last 100 checkpoint intervals (moving median)

public:
    Timer() { reset(); }
    void reset() {
        t_start = t_last = clock::now();
        window.clear();
    }
    float elapsed_seconds() { return std::chrono::duration<float>(clock::now() - t_start).count(); }
    float checkpoint_seconds() {
        const auto now = clock::now();
        const float dt = std::chrono::duration<float>(now - t_last).count();
        t_last = now;
        window.push_back(dt);
        if (window.size() > 100) window.pop_front();
        return dt;
    }
    float median_timestep() {
        if (window.empty()) return 0.0f;
        std::vector<float> v(window.begin(), window.end());
        std::sort(v.begin(), v.end());
        const size_t n = v.size();
        return (n % 2) ? v[n / 2] : 0.5f * (v[n / 2 - 1] + v[n / 2]);
    }
};

// Million lattice-node updates per second — the metric of BASELINE.json (the reference never computes it).
inline double lbm_b200_mlups(long long nodes, long long steps, double seconds) { return (double)nodes * (double)steps / seconds / 1e6; }
// Achieved bandwidth under the engine's algorithmic traffic model: 9 reads + 9 writes of fp32 per node update (72 B).
// (The reference's helper, src/util/timer.cuh:61-78, models its own multi-buffer pipeline at 128 B per node.)
inline float calculate_memory_bandwidth(long long nodes, int num_populations, float elapsed_seconds) {
    const double bytes = (double)nodes * 2.0 * num_populations * sizeof(float);
    return (float)(bytes / (1024.0 * 1024.0 * 1024.0) / elapsed_seconds);
}
