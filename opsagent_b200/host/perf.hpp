// perf.hpp — C++ host-side mirror of the reference's PerfStats (reference pkg/utils/perf.go): StartTimer / StopTimer accumulate per-operation
// durations and call counts (perf.go:64-121), TraceFunc returns the closer (perf.go:288), GetStats returns {"timers", "callCounts", "lastResetTime"}
// (perf.go:296-320), Reset clears them (perf.go:323-335), GetPerfStats() is the process-wide instance (perf.go:38-47).  Two deliberate differences
// (same as opsagent_b200/perf.py): running timers are keyed per (thread, operation), so concurrent requests do not overwrite each
// other's start time (the reference keeps ONE start time per operation name), and durations are integer nanoseconds — what Go's time.Duration
// marshals to in the /api/perf/stats JSON (pkg/handlers/perf.go:12-25).
#pragma once
#include <chrono>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <utility>

namespace opsagent {

class PerfStats {
public:
    void StartTimer(const std::string& operation) {
        std::lock_guard<std::mutex> lk(mu_);
        start_[{std::this_thread::get_id(), operation}] = std::chrono::steady_clock::now();
    }
    // -> elapsed ns; 0 if the timer was never started (perf.go:87-93)
    long long StopTimer(const std::string& operation) {
        const auto now = std::chrono::steady_clock::now();
        std::lock_guard<std::mutex> lk(mu_);
        auto it = start_.find({std::this_thread::get_id(), operation});
        if (it == start_.end()) return 0;
        const long long d = std::chrono::duration_cast<std::chrono::nanoseconds>(now - it->second).count();
        start_.erase(it);
        timers_[operation] += d; callCounts_[operation] += 1;
        return d;
    }
    void RecordMetric(const std::string& operation, long long duration_ns) { std::lock_guard<std::mutex> lk(mu_); timers_[operation] += duration_ns; callCounts_[operation] += 1; }
    std::function<void()> TraceFunc(const std::string& operation) { StartTimer(operation); return [this, operation] { StopTimer(operation); }; }
    struct Stats { std::map<std::string, long long> timers, callCounts; std::chrono::system_clock::time_point lastResetTime; };
    Stats GetStats() const { std::lock_guard<std::mutex> lk(mu_); return Stats{timers_, callCounts_, lastReset_}; }
    void Reset() { std::lock_guard<std::mutex> lk(mu_); timers_.clear(); callCounts_.clear(); start_.clear(); lastReset_ = std::chrono::system_clock::now(); }

private:
    mutable std::mutex mu_;
    std::map<std::pair<std::thread::id, std::string>, std::chrono::steady_clock::time_point> start_;
    std::map<std::string, long long> timers_, callCounts_;
    std::chrono::system_clock::time_point lastReset_ = std::chrono::system_clock::now();
};

inline PerfStats& GetPerfStats() { static PerfStats g; return g; }

}  // namespace opsagent
