// Offset-based sub-allocator for the symmetric heap.
//
// The reference carves its shared-memory heap with dlmalloc mspaces living *inside* the region
// (reference eplib/memory.c:147-263, eplib/dlmalloc.c).  A device slab cannot hold host-walkable metadata, so
// ours keeps the book-keeping on the host (ordered free list with coalescing, best-fit) and only hands out
// offsets; the same allocator therefore serves the CUDA slab and the host shared-memory slab.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <mutex>

namespace mlslb {

class SlabAllocator {
 public:
  SlabAllocator() = default;
  void reset(size_t base_offset, size_t bytes);
  // returns offset or SIZE_MAX when the slab is exhausted
  size_t alloc(size_t bytes, size_t align);
  void extend(size_t bytes);           // the slab grew by `bytes` at its end (device heap expansion)
  bool free(size_t offset);            // false if offset is not a live allocation
  size_t size_of(size_t offset) const; // 0 if unknown
  // find the live allocation containing [off, off+len); returns false if none
  bool contains(size_t off, size_t len) const;
  size_t bytes_in_use() const { return in_use_; }
  size_t capacity() const { return cap_; }
  size_t live_count() const { return live_.size(); }

 private:
  mutable std::mutex mu_;
  std::map<size_t, size_t> free_;   // offset -> length
  std::map<size_t, size_t> live_;   // offset -> length
  size_t base_ = 0, cap_ = 0, in_use_ = 0;
};

// Host memory for user buffers outside a shared slab: cache-line aligned, and from `thp_bytes` on 2 MiB aligned and marked
// for transparent huge pages (reference eplib/common.h:78-93, MLSL_THP_THRESHOLD_MB).  Freed with free().
void* aligned_host_alloc(size_t bytes, size_t align, size_t thp_bytes);

}  // namespace mlslb
