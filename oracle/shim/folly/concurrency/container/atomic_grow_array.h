// Stand-in for the one folly header the reference's index classes include
// (hnswalg_slim.h:12, hnswalg_slimq.h:11). Only operator[] is used there, and
// only by the delta/patch members that the search hot path never calls.
// Test infrastructure: lets oracle/_ref compile the UNMODIFIED reference headers.
#pragma once
#include <cstddef>
#include <deque>
#include <mutex>
namespace folly {
template <typename T> class atomic_grow_array {
  std::deque<T> d_;
  std::mutex m_;
 public:
  T &operator[](size_t i) {
    std::lock_guard<std::mutex> g(m_);
    if (i >= d_.size()) d_.resize(i + 1);
    return d_[i];
  }
};
}  // namespace folly
