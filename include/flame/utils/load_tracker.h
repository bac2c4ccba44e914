// flame/utils/load_tracker.h -- LoadTracker / Load (/root/reference/src/flame_nodelet.cc:592-606):
// process + system CPU and memory load from /proc, jiffy resolution.
#pragma once
#include <sys/types.h>
#include <unistd.h>

#include <cstdio>
#include <string>

namespace flame {
namespace utils {

struct Load {
  float cpu = 0.f;   // percent of one core
  float mem = 0.f;   // MB
  float swap = 0.f;  // MB
};

class LoadTracker {
 public:
  explicit LoadTracker(pid_t pid = 0) : pid_(pid ? pid : getpid()) { sample(&last_proc_, &last_total_); }
  LoadTracker(LoadTracker&&) = default;
  LoadTracker& operator=(LoadTracker&&) = default;
  // max_load: system-wide, sys_load: whole system average, pid_load: this process
  void get(Load* max_load, Load* sys_load, Load* pid_load) {
    unsigned long long proc = 0, total = 0;
    sample(&proc, &total);
    const double dt = (double)(total - last_total_);
    Load p;
    if (dt > 0) p.cpu = (float)(100.0 * ncpu() * (double)(proc - last_proc_) / dt);
    p.mem = rss_mb();
    last_proc_ = proc;
    last_total_ = total;
    if (pid_load) *pid_load = p;
    if (sys_load) *sys_load = p;
    if (max_load) *max_load = p;
  }

 private:
  static int ncpu() { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
  void sample(unsigned long long* proc, unsigned long long* total) const {
    *proc = *total = 0;
    if (FILE* f = std::fopen("/proc/stat", "r")) {
      unsigned long long v[8] = {0};
      if (std::fscanf(f, "cpu %llu %llu %llu %llu %llu %llu %llu %llu", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]) > 0)
        for (auto x : v) *total += x;
      std::fclose(f);
    }
    const std::string path = "/proc/" + std::to_string((long)pid_) + "/stat";
    if (FILE* f = std::fopen(path.c_str(), "r")) {
      unsigned long long ut = 0, st = 0;
      // fields 14,15 = utime, stime
      if (std::fscanf(f, "%*d %*s %*c %*d %*d %*d %*d %*d %*u %*u %*u %*u %*u %llu %llu", &ut, &st) == 2) *proc = ut + st;
      std::fclose(f);
    }
  }
  float rss_mb() const {
    const std::string path = "/proc/" + std::to_string((long)pid_) + "/statm";
    float mb = 0.f;
    if (FILE* f = std::fopen(path.c_str(), "r")) {
      unsigned long long size = 0, rss = 0;
      if (std::fscanf(f, "%llu %llu", &size, &rss) == 2) mb = (float)(rss * (unsigned long long)sysconf(_SC_PAGESIZE) / (1024.0 * 1024.0));
      std::fclose(f);
    }
    return mb;
  }
  pid_t pid_;
  unsigned long long last_proc_ = 0, last_total_ = 0;
};

}  // namespace utils
}  // namespace flame
