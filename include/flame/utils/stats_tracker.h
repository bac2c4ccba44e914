// flame/utils/stats_tracker.h -- flame::utils::StatsTracker as used by the frontends
// (/root/reference/src/flame_nodelet.cc:533-590,747-749; src/utils.h:72-83).
#pragma once
#include <chrono>
#include <string>
#include <unordered_map>

namespace flame {
namespace utils {

class StatsTracker {
 public:
  explicit StatsTracker(const std::string& prefix = "") : prefix_(prefix) {}
  void tick(const std::string& key) { ticks_[key] = std::chrono::steady_clock::now(); }
  // Returns (and records) the elapsed milliseconds since tick(key).
  double tock(const std::string& key) {
    auto it = ticks_.find(key);
    if (it == ticks_.end()) return 0.0;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - it->second).count();
    timings_[prefix_ + key] = ms;
    return ms;
  }
  void set(const std::string& key, double value) { stats_[prefix_ + key] = value; }
  void setTiming(const std::string& key, double ms) { timings_[prefix_ + key] = ms; }
  const std::unordered_map<std::string, double>& stats() const { return stats_; }
  const std::unordered_map<std::string, double>& timings() const { return timings_; }
  double stats(const std::string& key) const { auto it = stats_.find(key); return it == stats_.end() ? 0.0 : it->second; }
  double timings(const std::string& key) const { auto it = timings_.find(key); return it == timings_.end() ? 0.0 : it->second; }

 private:
  std::string prefix_;
  std::unordered_map<std::string, std::chrono::steady_clock::time_point> ticks_;
  std::unordered_map<std::string, double> stats_, timings_;
};

}  // namespace utils
}  // namespace flame
