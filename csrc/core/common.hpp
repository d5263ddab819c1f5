// Common internal types for the mlsl-b200 runtime.
//
// Parity notes: data types / reduction ops mirror the public enums of the reference
// (reference include/mlsl.hpp:82-146); BF16/F16/I32/F8 are B200-side extensions.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace mlslb {

constexpr int kMaxDevRanks = 16;    // ranks addressable by one device kernel (single NVSwitch domain)
constexpr int kMaxHostRanks = 64;   // ranks supported by the host shared-memory backend
constexpr int kMaxGroupRows = 64;   // concurrent process-group "rows" (one row per collective group creation)
constexpr int kMaxChannels = 152;    // max CTAs ("channels", the GPU analogue of endpoints) per collective kernel

enum class DType : int { F32 = 0, F64 = 1, U8 = 2, BF16 = 3, F16 = 4, I32 = 5, F8E4M3 = 6 };
enum class RedOp : int { SUM = 0, MIN = 1, MAX = 2 };

inline size_t dtype_size(DType d) {
  switch (d) {
    case DType::F32: return 4;
    case DType::F64: return 8;
    case DType::U8: return 1;
    case DType::BF16: return 2;
    case DType::F16: return 2;
    case DType::I32: return 4;
    case DType::F8E4M3: return 1;
  }
  return 0;
}

inline const char* dtype_name(DType d) {
  switch (d) {
    case DType::F32: return "f32";
    case DType::F64: return "f64";
    case DType::U8: return "u8";
    case DType::BF16: return "bf16";
    case DType::F16: return "f16";
    case DType::I32: return "i32";
    case DType::F8E4M3: return "f8e4m3";
  }
  return "?";
}

enum class OpKind : int {
  BARRIER = 0,
  BCAST,
  REDUCE,
  ALLREDUCE,
  ALLTOALL,
  ALLTOALLV,
  GATHER,
  ALLGATHER,
  ALLGATHERV,
  SCATTER,
  REDUCE_SCATTER,
  SENDRECV_LIST,   // ring-shift style point-to-point list (the reference declares it but never wires it up)
  FUSED_UPDATE,    // B200 extension: reduce-scatter + optimizer step + all-gather in one kernel
  GEMM_RS,         // B200 extension: tcgen05 GEMM whose epilogue reduce-scatters over peer memory
  AG_GEMM,         // B200 extension (experimental): all-gather of the row shards overlapped with the GEMM that eats them
};

inline const char* opkind_name(OpKind k) {
  static const char* n[] = {"Barrier", "Bcast", "Reduce", "AllReduce", "AlltoAll", "AlltoAllv", "Gather",
                            "AllGather", "AllGatherv", "Scatter", "ReduceScatter", "SendRecvList",
                            "FusedUpdate", "GemmRS", "AllGatherGemm"};
  return n[(int)k];
}

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

uint64_t now_ns();          // monotonic clock
uint64_t cycles_now();      // rdtsc on x86 (API compatibility with the reference's cycle counters)

}  // namespace mlslb
