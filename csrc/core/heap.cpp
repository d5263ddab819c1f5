#include "heap.hpp"

#include <sys/mman.h>

#include <algorithm>
#include <cstdlib>

#include "common.hpp"

namespace mlslb {

void SlabAllocator::reset(size_t base_offset, size_t bytes) {
  std::lock_guard<std::mutex> g(mu_);
  free_.clear();
  live_.clear();
  base_ = base_offset;
  cap_ = bytes;
  in_use_ = 0;
  if (bytes) free_[base_offset] = bytes;
}

size_t SlabAllocator::alloc(size_t bytes, size_t align) {
  if (bytes == 0) bytes = 1;
  if (align < 64) align = 64;
  bytes = round_up(bytes, 64);
  std::lock_guard<std::mutex> g(mu_);
  // best fit among blocks that can hold the aligned request
  auto best = free_.end();
  size_t best_len = SIZE_MAX;
  for (auto it = free_.begin(); it != free_.end(); ++it) {
    size_t start = round_up(it->first, align);
    size_t pad = start - it->first;
    if (it->second >= pad + bytes && it->second < best_len) {
      best = it;
      best_len = it->second;
    }
  }
  if (best == free_.end()) return SIZE_MAX;
  size_t blk_off = best->first, blk_len = best->second;
  size_t start = round_up(blk_off, align);
  size_t pad = start - blk_off;
  free_.erase(best);
  if (pad) free_[blk_off] = pad;
  size_t tail = blk_len - pad - bytes;
  if (tail) free_[start + bytes] = tail;
  live_[start] = bytes;
  in_use_ += bytes;
  return start;
}

void SlabAllocator::extend(size_t bytes) {
  if (!bytes) return;
  std::lock_guard<std::mutex> g(mu_);
  const size_t off = base_ + cap_;
  cap_ += bytes;
  // merge with a free block that ends where the new space starts
  if (!free_.empty()) {
    auto last = std::prev(free_.end());
    if (last->first + last->second == off) {
      last->second += bytes;
      return;
    }
  }
  free_[off] = bytes;
}

bool SlabAllocator::free(size_t offset) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.find(offset);
  if (it == live_.end()) return false;
  size_t len = it->second;
  live_.erase(it);
  in_use_ -= len;
  auto ins = free_.emplace(offset, len).first;
  // coalesce with next
  auto nxt = std::next(ins);
  if (nxt != free_.end() && ins->first + ins->second == nxt->first) {
    ins->second += nxt->second;
    free_.erase(nxt);
  }
  // coalesce with previous
  if (ins != free_.begin()) {
    auto prv = std::prev(ins);
    if (prv->first + prv->second == ins->first) {
      prv->second += ins->second;
      free_.erase(ins);
    }
  }
  return true;
}

size_t SlabAllocator::size_of(size_t offset) const {
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.find(offset);
  return it == live_.end() ? 0 : it->second;
}

bool SlabAllocator::contains(size_t off, size_t len) const {
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.upper_bound(off);
  if (it == live_.begin()) return false;
  --it;
  return off >= it->first && off + len <= it->first + it->second;
}

void* aligned_host_alloc(size_t bytes, size_t align, size_t thp_bytes) {
  constexpr size_t kTwoMb = (size_t)2 << 20;
  align = std::max<size_t>(align ? align : 64, 64);
  const bool thp = thp_bytes && bytes >= thp_bytes && kTwoMb % align == 0;
  if (thp) align = kTwoMb;
  void* p = nullptr;
  if (posix_memalign(&p, align, round_up(std::max<size_t>(bytes, 1), 64)) != 0) return nullptr;
#ifdef MADV_HUGEPAGE
  if (thp && bytes >= kTwoMb) madvise(p, (bytes / kTwoMb) * kTwoMb, MADV_HUGEPAGE);     // the whole 2 MiB pages of the block
#endif
  return p;
}

}  // namespace mlslb
