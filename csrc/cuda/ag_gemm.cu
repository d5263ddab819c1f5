// Fused all-gather + GEMM (experimental, opt-in) - the other half of SURVEY K14 and the forward of a column-parallel layer
// under sequence parallelism: every rank holds a row shard X_r[M/P, K] and needs
//
//   Y[M, N] = concat_rows(X_0 .. X_{P-1}) * W[N, K]^T          (bf16 inputs, fp32 accumulation, bf16 or fp32 result)
//
// Instead of an all-gather followed by a GEMM, ONE persistent kernel does both and overlaps them tile by tile:
//   * "copy" CTAs (blockIdx < C) are copy engines: one thread streams the row blocks of all ranks - its own first, then
//     rank me+1, me+2, ... - from the owners' slabs into the local gathered buffer with 1-D bulk copies
//     (cp.async.bulk global -> shared -> global, 12 x 16 KB in flight per CTA, no SM cycles spent on the data) and
//     publishes one flag per 128-row tile as soon as that tile has landed;
//   * "gemm" CTAs run the same TMA / tcgen05 / TMEM pipeline as k_gemm_rs (csrc/cuda/gemm_rs.cu); their TMA producer
//     visits the tiles in arrival order and only waits for the flag of the tile it is about to load, so the tensor cores
//     work on the local rows while the remote ones are still in flight over NVLink.
// The gathered X stays in `gathered` for the backward pass (dW = dY^T * X needs all rows).
// Peer lines are not cached in the reader's L2, so reading A tiles straight from the peers would fetch every tile
// tiles_n times over NVLink; the local copy is fetched once and re-read from HBM / L2.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "core/log.hpp"
#include "cuda/kernels.hpp"
#include "cuda/umma.cuh"

namespace mlslb {

namespace {

using namespace umma;

constexpr int BM = 128, BN = 256;
constexpr int kStages = 4;
constexpr int kAccBufs = 2;
constexpr int kTmemCols = kAccBufs * BN;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr uint32_t kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
constexpr uint32_t kEpiBytesPerWarp = 32 * 64 * 2;
constexpr size_t kSmemBytes = 1024 + kStages * (kStageBytesA + kStageBytesB) + kEpiWarps * kEpiBytesPerWarp + 256;
// copy engine: a ring of kSlots pieces; the store of a piece is issued kLag steps after its load
constexpr int kSlots = 12, kLag = 6;
constexpr uint32_t kPiece = 16384;
static_assert((size_t)kSlots * kPiece + 1024 + 256 <= kSmemBytes, "copy ring must fit into the GEMM CTA's shared memory");

struct AgGemmArgs {
  int M, N, K;
  unsigned long long x_off;       // slab offset of this rank's shard X_r [M/P, K] bf16
  __nv_bfloat16* gathered;        // [M, K] bf16, local
  void* out;                      // [M, N] bf16 or fp32, local
  int out_fp32;
  int copy_ctas;                  // C: blockIdx < C are copy engines
  unsigned long long* counters;   // [gridDim.x] private launch counters (epoch of the flags)
  unsigned long long* ready;      // [M / 128] one flag per row tile: == epoch when the tile is in `gathered`
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreads, 1)
k_ag_gemm(DevComm dc, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, AgGemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ PeerTable pt;
  __shared__ unsigned long long s_epoch;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = dc.nranks, me = dc.me;
  const int C = g.copy_ctas, G = (int)gridDim.x - C;
  const int tiles_pb = (g.M / P) / BM;            // row tiles per rank block
  const int tiles_n = g.N / BN;
  const int kblocks = g.K / BK;

  if (threadIdx.x == 0) {
    const unsigned long long e = g.counters[blockIdx.x] + 1;   // every CTA counts its own launches: same value everywhere
    g.counters[blockIdx.x] = e;
    s_epoch = e;
  }
  __syncthreads();
  const unsigned long long epoch = s_epoch;

  if ((int)blockIdx.x < C) {
    // =============================== copy engine ===============================
    const unsigned long long ticket = comm_begin(dc, pt, g.x_off, g.x_off, NoAux());   // where every rank's shard lives
    uint64_t* lbar = reinterpret_cast<uint64_t*>(smem + (size_t)kSlots * kPiece);
    if (threadIdx.x == 0) {
      for (int i = 0; i < kSlots; ++i) mbar_init(&lbar[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const size_t tile_bytes = (size_t)BM * g.K * 2;
      const int pieces_per_tile = (int)(tile_bytes / kPiece);
      // my tiles in arrival order: sequence number s = j * tiles_pb + ti (j-th row block after mine), s % C == blockIdx
      const int total_seq = P * tiles_pb;
      const int first = (int)blockIdx.x;
      const int my_count = first < total_seq ? (total_seq - first + C - 1) / C : 0;
      const long long total = (long long)my_count * pieces_per_tile;
      auto locate = [&](long long piece, const char*& src, char*& dst, int& tm, bool& last) {
        const int k = (int)(piece / pieces_per_tile), pp = (int)(piece % pieces_per_tile);
        const int s = first + k * C;
        const int j = s / tiles_pb, ti = s % tiles_pb;
        int q = me + j;
        if (q >= P) q -= P;
        tm = q * tiles_pb + ti;
        src = pt.send[q] + (size_t)ti * tile_bytes + (size_t)pp * kPiece;
        dst = reinterpret_cast<char*>(g.gathered) + (size_t)tm * tile_bytes + (size_t)pp * kPiece;
        last = pp == pieces_per_tile - 1;
      };
      for (long long i = 0; i < total + kLag; ++i) {
        if (i >= kSlots && i - kSlots < total) {
          // piece i - kSlots was stored kSlots - kLag steps ago and kSlots - 1 - kLag groups were committed after it
          asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kSlots - 1 - kLag) : "memory");
          const char* s0;
          char* d0;
          int tm;
          bool last;
          locate(i - kSlots, s0, d0, tm, last);
          if (last) {
            asm volatile("fence.proxy.async;" ::: "memory");
            st_release_gpu(&g.ready[tm], epoch);
          }
        }
        if (i < total) {
          const char* src;
          char* dst;
          int tm;
          bool last;
          locate(i, src, dst, tm, last);
          const int slot = (int)(i % kSlots);
          mbar_expect_tx(&lbar[slot], kPiece);
          bulk_load(smem + (size_t)slot * kPiece, src, kPiece, &lbar[slot]);
        }
        const long long jst = i - kLag;
        if (jst >= 0 && jst < total) {
          const char* src;
          char* dst;
          int tm;
          bool last;
          locate(jst, src, dst, tm, last);
          const int slot = (int)(jst % kSlots);
          mbar_wait(&lbar[slot], (uint32_t)((jst / kSlots) & 1));
          bulk_store(dst, smem + (size_t)slot * kPiece, kPiece);
        }
      }
      // drain: everything stored, publish the flags that are still outstanding
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
      const long long first_unflagged = total + kLag - kSlots > 0 ? total + kLag - kSlots : 0;
      for (long long p = first_unflagged; p < total; ++p) {
        const char* s0;
        char* d0;
        int tm;
        bool last;
        locate(p, s0, d0, tm, last);
        if (last) st_release_gpu(&g.ready[tm], epoch);
      }
    }
    // my shard may be overwritten by the caller only when every peer has copied it
    comm_sync(dc, pt, ticket, 1, false);
    return;
  }

  // =============================== GEMM CTA ===============================
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kStageBytesA;
  uint8_t* sEpi = sB + kStages * kStageBytesB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + kEpiWarps * kEpiBytesPerWarp);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + kAccBufs;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + kAccBufs);
  const int gidx = (int)blockIdx.x - C;
  const int ntiles = P * tiles_pb * tiles_n;
  const int my_tiles = gidx < ntiles ? (ntiles - gidx + G - 1) / G : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAccBufs; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_smem;

  // tile `it` of this CTA -> (row tile tm in global numbering, column tile tn), visited in arrival order
  auto tile_of = [&](int it, int& tm, int& tn) {
    const int tseq = gidx + it * G;
    const int rs = tseq / tiles_n;                 // arrival sequence of the row tile
    tn = tseq % tiles_n;
    const int j = rs / tiles_pb, ti = rs % tiles_pb;
    int q = me + j;
    if (q >= P) q -= P;
    tm = q * tiles_pb + ti;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      int have_tm = -1;
      for (int it = 0; it < my_tiles; ++it) {
        int tm, tn;
        tile_of(it, tm, tn);
        if (tm != have_tm) {
          // the rows of this tile must have landed in `gathered` (deadline + error word like every other spin)
          spin_on(dc, g.ready + tm, [epoch](unsigned long long w) { return w == epoch; });
          (void)ld_acquire_gpu(g.ready + tm);
          asm volatile("fence.proxy.async;" ::: "memory");
          have_tm = tm;
        }
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], kStageBytesA + kStageBytesB);
          tma_load_2d(sA + stage * kStageBytesA, &map_a, &full[stage], kb * BK, tm * BM);
          tma_load_2d(sB + stage * kStageBytesB, &map_w, &full[stage], kb * BK, tn * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BM, BN);
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t buf = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
      mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + buf * BN;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a0 = smem_u32(sA + stage * kStageBytesA), b0 = smem_u32(sB + stage * kStageBytesB);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(a0 + k * UMMA_K * 2), db = make_smem_desc(b0 + k * UMMA_K * 2);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb == kblocks - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> bf16 / fp32 -> Y (local) =====================
    const int ew = (warp - 4) & 3;
    const int ch = (warp - 4) >> 2;
    uint8_t* myepi = sEpi + (warp - 4) * kEpiBytesPerWarp;
    for (int it = 0; it < my_tiles; ++it) {
      int tm, tn;
      tile_of(it, tm, tn);
      const uint32_t buf = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
      mbar_wait(&acc_full[buf], use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const size_t row0 = (size_t)tm * BM + ew * 32;
#pragma unroll 1
      for (int hh = 0; hh < BN / 128; ++hh) {
        const int half = ch * (BN / 128) + hh;
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * BN + half * 64;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const size_t col0 = (size_t)tn * BN + half * 64;
        if (g.out_fp32) {
          // lane = row: 64 consecutive floats of one row
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + (row0 + lane) * g.N + col0);
#pragma unroll
          for (int c = 0; c < 16; ++c)
            o[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                               __uint_as_float(v[4 * c + 3]));
        } else {
          uint4* rowp = reinterpret_cast<uint4*>(myepi + lane * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 q;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[8 * c + 0]), __uint_as_float(v[8 * c + 1]));
            __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3]));
            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5]));
            __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7]));
            q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
            q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
            rowp[c ^ (lane & 7)] = q;
          }
          __syncwarp();
          __nv_bfloat16* dst_base = reinterpret_cast<__nv_bfloat16*>(g.out) + row0 * g.N + col0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3), c = lane & 7;
            const uint4 q = *reinterpret_cast<const uint4*>(myepi + r * 128 + ((c ^ (r & 7)) * 16));
            *reinterpret_cast<uint4*>(reinterpret_cast<char*>(dst_base + (size_t)r * g.N) + c * 16) = q;
          }
          __syncwarp();
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

}  // namespace

const char* ag_gemm_check(int M, int N, int K, int P) {
  if (M <= 0 || N <= 0 || K <= 0) return "empty problem";
  if (M % (BM * P) != 0) return "M must be a multiple of 128 * group size";
  if (N % BN != 0) return "N must be a multiple of 256";
  if (K % BK != 0) return "K must be a multiple of 64";
  return nullptr;
}

size_t ag_gemm_scratch_bytes(int M, int max_ctas) { return ((size_t)max_ctas + (size_t)(M / BM)) * sizeof(unsigned long long); }

// grid: `copy_ctas` copy engines (= handshake channels) + GEMM CTAs; both are pure functions of (shape, SM budget)
void ag_gemm_grid(int M, int N, int P, int max_ctas, int* copy_ctas, int* total_ctas) {
  const int row_tiles = M / BM;
  int c = std::max(1, std::min(std::min(16, row_tiles), max_ctas / 4));
  int tiles = row_tiles * (N / BN);
  int gemm = std::max(1, std::min(tiles, max_ctas - c));
  *copy_ctas = c;
  *total_ctas = c + gemm;
}

cudaError_t launch_ag_gemm(const DevComm& dc, unsigned long long x_off, const void* w, void* gathered, void* out, bool out_fp32,
                           int M, int N, int K, int copy_ctas, int total_ctas, void* scratch, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ag_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap ma, mw;
  if (!make_map(&ma, gathered, M, K, BM) || !make_map(&mw, w, N, K, BN)) return cudaErrorInvalidValue;
  AgGemmArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.x_off = x_off;
  g.gathered = (__nv_bfloat16*)gathered;
  g.out = out;
  g.out_fp32 = out_fp32 ? 1 : 0;
  g.copy_ctas = copy_ctas;
  g.counters = (unsigned long long*)scratch;
  g.ready = (unsigned long long*)scratch + total_ctas;
  k_ag_gemm<<<total_ctas, kThreads, kSmemBytes, s>>>(dc, ma, mw, g);
  return cudaGetLastError();
}

}  // namespace mlslb
