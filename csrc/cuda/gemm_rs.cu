// Fused GEMM + reduce-scatter for tensor (model) parallel layers - SURVEY K14.
//
// The reference's model parallelism (OT_CC with modelParts > 1: every rank holds a K-slice of the weights, its output
// is a full-size PARTIAL sum that must be reduce-scattered over the model group, reference src/mlsl_impl.cpp:139-175)
// costs a GEMM and then a collective.  Here both are ONE persistent sm_100a kernel:
//
//   C_r[M, N] = A_r[M, K_r] * W_r[N, K_r]^T        (bf16 inputs, fp32 accumulation)
//   out_r[M/P, N] = sum_q C_q[rows of rank r]       (rank r owns rows [r*M/P, (r+1)*M/P))
//
//   * warp 0  : TMA producer  - cp.async.bulk.tensor 2D loads of A / W tiles (128B swizzle) into a 4-stage smem ring,
//               completion on mbarriers;
//   * warp 1  : MMA issuer    - one elected thread issues tcgen05.mma (cta_group::1, kind::f16, M128 x N256 x K16)
//               with the accumulator in TMEM; two accumulator buffers (2 x 256 columns = all of TMEM) so the epilogue of tile i
//               overlaps the main loop of tile i+1; tcgen05.commit frees smem stages / publishes the accumulator;
//   * warp 2  : TMEM allocation (512 columns) / deallocation;
//   * warps 4-11: epilogue    - tcgen05.ld (32x32b.x32) the accumulator, convert to bf16, transpose through smem and
//               write the tile with 128-byte row segments straight into the OWNER rank's staging slot over NVLink
//               (peer store; local store when this rank owns the rows) while the tensor core already runs the next tile.
//   After its last tile every CTA does the channel handshake (fence + flags) with the same CTA of the peers - the
//   tile -> CTA map is identical on every rank, so the P partials of a tile are all produced by "its" channel - and
//   then sums the P staged partials of the tiles it owns into the output (fp32 accumulate, bf16 or fp32 result).
// Partials travel as bf16 (what a separate bf16 GEMM followed by a bf16 reduce-scatter would move); the order of the
// final summation is fixed (rank 0..P-1), so results are deterministic.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <tuple>

#include "core/log.hpp"
#include "cuda/kernels.hpp"
#include "cuda/umma.cuh"

namespace mlslb {

namespace {

using namespace umma;

constexpr int BM = 128, BN = 256;          // CTA tile; BK * 2 bytes = one 128-byte swizzle row.  N = 256 keeps the
                                                    // smem operand traffic (A 4 KB + B 8 KB per 128-cycle MMA) under 128 B/clk
constexpr int kStages = 4;
constexpr int kAccBufs = 2;
constexpr int kTmemCols = kAccBufs * BN;           // 512: the whole tensor memory, double-buffered accumulator
constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter, each takes half of the columns
constexpr int kThreads = 128 + kEpiWarps * 32;     // warps 0..3: producer / mma / tmem / idle, warps 4..11: epilogue
constexpr uint32_t kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
constexpr uint32_t kEpiBytesPerWarp = 32 * 64 * 2; // 32 rows x 64 bf16 columns
constexpr size_t kSmemBytes = 1024 /*align slack*/ + kStages * (kStageBytesA + kStageBytesB) + kEpiWarps * kEpiBytesPerWarp + 256;

struct GemmRsArgs {
  int M, N, K;                      // this rank's partial product: [M, N], reduction length K (= K_r)
  unsigned long long stage_off;     // slab offset of the staging area: [P sources][M/P rows][N] bf16
  void* out;                        // [M/P, N], bf16 or fp32
  int out_fp32;
};

__global__ void __launch_bounds__(kThreads, 1)
k_gemm_rs(DevComm dc, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, GemmRsArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kStageBytesA;
  uint8_t* sEpi = sB + kStages * kStageBytesB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + kEpiWarps * kEpiBytesPerWarp);
  uint64_t* full = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* acc_full = empty + kStages;     // [kAccBufs] MMA -> epilogue
  uint64_t* acc_empty = acc_full + kAccBufs;// [kAccBufs] epilogue -> MMA
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + kAccBufs);
  __shared__ PeerTable pt;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = dc.nranks, me = dc.me;
  const int tiles_m = g.M / BM, tiles_n = g.N / BN, ntiles = tiles_m * tiles_n;
  const int kblocks = g.K / BK;
  const int rows_per_rank = g.M / P;

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
      mbar_init(&acc_empty[i], kEpiWarps);   // one arrival per epilogue warp
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

  // Opening handshake: learn where every peer's staging area lives (and that its kernel - hence everything queued
  // before it on its stream - has started).  All threads take part in the bar.syncs inside.
  const unsigned long long ticket = comm_begin(dc, pt, g.stage_off, g.stage_off, NoAux());

  // Tile schedule: tile t belongs to channel t % gridDim.x on EVERY rank; within a CTA the tiles are visited starting
  // at an offset that depends on the rank, so at any moment the ranks push to different owners.
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int rot = my_tiles > 0 ? (me * ((my_tiles + P - 1) / P)) % my_tiles : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int t = (int)blockIdx.x + ((it + rot) % my_tiles) * (int)gridDim.x;
        const int tm = t / tiles_n, tn = t % tiles_n;
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
      mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);       // epilogue has drained this accumulator buffer
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
          umma_commit(&empty[stage]);                    // smem stage reusable once these MMAs have read it
          if (kb == kblocks - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> bf16 -> owner's staging slot (peer memory) =====================
    const int ew = (warp - 4) & 3;                       // TMEM lane quarter [32*ew, 32*ew + 32) (= warp id % 4)
    const int ch = (warp - 4) >> 2;                      // which half of the tile's columns this warp drains
    uint8_t* myepi = sEpi + (warp - 4) * kEpiBytesPerWarp;
    for (int it = 0; it < my_tiles; ++it) {
      const int t = (int)blockIdx.x + ((it + rot) % my_tiles) * (int)gridDim.x;
      const int tm = t / tiles_n, tn = t % tiles_n;
      const uint32_t buf = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
      mbar_wait(&acc_full[buf], use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row0 = tm * BM + ew * 32;                // first global row handled by this warp
      const int owner = row0 / rows_per_rank;
      // destination: staging[src = me][row - owner*rows_per_rank][col] in the OWNER's slab
      __nv_bfloat16* dst_base = reinterpret_cast<__nv_bfloat16*>(pt.send[owner]) +
                                ((size_t)me * rows_per_rank + (size_t)(row0 - owner * rows_per_rank)) * g.N + (size_t)tn * BN;
#pragma unroll 1
      for (int hh = 0; hh < BN / 128; ++hh) {
        const int half = ch * (BN / 128) + hh;           // 64-column chunk index inside the tile
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * BN + half * 64;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // lane = row: pack 64 fp32 -> 64 bf16 (128 B) and park the row in smem, 16-byte chunks XOR-swizzled by row
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
        // 8 lanes write one full 128-byte row segment: 4 rows per instruction, 8 instructions per 32 rows
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3), c = lane & 7;
          const uint4 q = *reinterpret_cast<const uint4*>(myepi + r * 128 + ((c ^ (r & 7)) * 16));
          st16(reinterpret_cast<char*>(dst_base + (size_t)r * g.N + half * 64) + c * 16, q);
        }
        __syncwarp();
      }
      // accumulator buffer drained: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }

  // ---- all partial tiles of this channel are on their way: fence + flag handshake with the same channel of the peers
  comm_sync(dc, pt, ticket, 1, true);

  // ---- owner-side reduction of the tiles this channel is responsible for ---------------------------------------------
  {
    const __nv_bfloat16* stage_me = reinterpret_cast<const __nv_bfloat16*>(pt.send[me]);
    const size_t slot = (size_t)rows_per_rank * g.N;      // elements per source slot
    for (int it = 0; it < my_tiles; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int tm = t / tiles_n, tn = t % tiles_n;
      const int row0 = tm * BM;
      if (row0 / rows_per_rank != me) continue;           // tiles never straddle owners (rows_per_rank % BM == 0)
      const int lrow0 = row0 - me * rows_per_rank;
      // 128 x 256 tile, 8 bf16 (16 B) per item; kU independent items per thread in flight (the loop is latency bound)
      constexpr int kU = 4, kItems = BM * (BN / 8);
      for (int idx0 = threadIdx.x; idx0 < kItems; idx0 += kThreads * kU) {
        float acc[kU][8];
        size_t off[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int idx = idx0 + u * kThreads;
          const int r = idx / (BN / 8), c8 = idx % (BN / 8);
          off[u] = (size_t)(lrow0 + r) * g.N + (size_t)tn * BN + c8 * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[u][j] = 0.f;
        }
        for (int q = 0; q < P; ++q) {
          uint4 v[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u)
            if (idx0 + u * kThreads < kItems) v[u] = __ldcg(reinterpret_cast<const uint4*>(stage_me + (size_t)q * slot + off[u]));
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[u][2 * j] += __uint_as_float(w[j] << 16);
              acc[u][2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          if (idx0 + u * kThreads >= kItems) continue;
          if (g.out_fp32) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + off[u]);
            o[0] = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
            o[1] = make_float4(acc[u][4], acc[u][5], acc[u][6], acc[u][7]);
          } else {
            uint4 o;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[u][0], acc[u][1]), h1 = __floats2bfloat162_rn(acc[u][2], acc[u][3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[u][4], acc[u][5]), h3 = __floats2bfloat162_rn(acc[u][6], acc[u][7]);
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + off[u]) = o;
          }
        }
      }
    }
  }
  // keep the staging area stable until every peer has finished reading... nobody reads remote staging in phase 2, but a
  // fast peer must not start the NEXT launch's pushes into my staging while I still reduce: closing handshake.
  comm_sync(dc, pt, ticket, 2, false);

  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ====================================================================================================================
// 2-CTA variant (experimental, MLSL_GEMM_2CTA=1): a cluster of two CTAs on one TPC computes a 256 x 256 tile with
// tcgen05.mma.cta_group::2 (M256 x N256 x K16).  Each CTA stages ITS 128 rows of A and ITS 128-row half of the W tile, so
// a CTA moves 16 KB + 16 KB per k-block instead of 16 KB + 32 KB - the 1-CTA kernel sits at the shared-memory bandwidth
// limit (profiles/README.md 5b), this form frees a third of it and deepens the ring to 6 stages in the same 192 KB.
// Accumulator rows [0,128) live in the leader's TMEM, rows [128,256) in the peer's: every CTA drains its own half with the
// same epilogue as the 1-CTA kernel and stays its own communication channel.  Pipeline:
//   full[s]      (leader only)  <- TMA bytes of BOTH CTAs (cp.async.bulk.tensor ... .cta_group::2 to the leader's barrier)
//   empty[s]     (both CTAs)    <- tcgen05.commit.cta_group::2 ... multicast::cluster, mask 0b11
//   acc_full[b]  (both CTAs)    <- the same multicast commit after the last k-block
//   acc_empty[b] (leader only)  <- 2 x kEpiWarps arrivals, the peer's epilogue warps arrive remotely
constexpr int kStages2 = 6;
constexpr int BN2H = BN / 2;                                       // W rows staged by one CTA
constexpr uint32_t kStageBytesA2 = BM * BK * 2, kStageBytesB2 = BN2H * BK * 2;
constexpr size_t kSmemBytes2 = 1024 + kStages2 * (kStageBytesA2 + kStageBytesB2) + kEpiWarps * kEpiBytesPerWarp + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_gemm_rs2(DevComm dc, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, GemmRsArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages2 * kStageBytesA2;
  uint8_t* sEpi = sB + kStages2 * kStageBytesB2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + kEpiWarps * kEpiBytesPerWarp);
  uint64_t* full = bars;                      // [kStages2]  used in the leader only
  uint64_t* empty = bars + kStages2;          // [kStages2]
  uint64_t* acc_full = empty + kStages2;      // [kAccBufs]
  uint64_t* acc_empty = acc_full + kAccBufs;  // [kAccBufs]  used in the leader only
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + kAccBufs);
  __shared__ PeerTable pt;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = dc.nranks, me = dc.me;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int cluster_id = (int)blockIdx.x >> 1, nclusters = (int)gridDim.x >> 1;
  const int tiles_n = g.N / BN, nsuper = (g.M / (2 * BM)) * tiles_n;   // 256 x 256 super tiles
  const int kblocks = g.K / BK;
  const int rows_per_rank = g.M / P;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAccBufs; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 2 * kEpiWarps);   // the epilogue warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // the same warp of both CTAs: a pair-wide allocation at the same column address
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated, before anyone signals the peer
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_smem;

  const unsigned long long ticket = comm_begin(dc, pt, g.stage_off, g.stage_off, NoAux());

  const int my_tiles = (nsuper - cluster_id + nclusters - 1) / nclusters;
  const int rot = my_tiles > 0 ? (me * ((my_tiles + P - 1) / P)) % my_tiles : 0;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own rows of A, own half of the W tile) =====================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int sidx = cluster_id + ((it + rot) % my_tiles) * nclusters;
        const int tm2 = sidx / tiles_n, tn = sidx % tiles_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t lbar = map_to_cta(smem_u32(&full[stage]), 0);
          if (leader) mbar_expect_tx(&full[stage], 2 * (kStageBytesA2 + kStageBytesB2));
          tma_load_2d_2cta(sA + stage * kStageBytesA2, &map_a, lbar, kb * BK, tm2 * 2 * BM + (int)cta * BM);
          tma_load_2d_2cta(sB + stage * kStageBytesB2, &map_w, lbar, kb * BK, tn * BN + (int)cta * BN2H);
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    constexpr uint32_t idesc = make_idesc(2 * BM, BN);
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
          const uint32_t a0 = smem_u32(sA + stage * kStageBytesA2), b0 = smem_u32(sB + stage * kStageBytesB2);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(a0 + k * UMMA_K * 2), db = make_smem_desc(b0 + k * UMMA_K * 2);
            umma_bf16_2cta(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(&empty[stage]);
          if (kb == kblocks - 1) umma_commit_2cta(&acc_full[buf]);
        }
        __syncwarp();
        if (++stage == kStages2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: this CTA's 128 accumulator rows -> owner's staging slot =====================
    const int ew = (warp - 4) & 3;
    const int ch = (warp - 4) >> 2;
    uint8_t* myepi = sEpi + (warp - 4) * kEpiBytesPerWarp;
    for (int it = 0; it < my_tiles; ++it) {
      const int sidx = cluster_id + ((it + rot) % my_tiles) * nclusters;
      const int tm = (sidx / tiles_n) * 2 + (int)cta, tn = sidx % tiles_n;
      const uint32_t buf = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
      mbar_wait(&acc_full[buf], use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row0 = tm * BM + ew * 32;
      const int owner = row0 / rows_per_rank;
      __nv_bfloat16* dst_base = reinterpret_cast<__nv_bfloat16*>(pt.send[owner]) +
                                ((size_t)me * rows_per_rank + (size_t)(row0 - owner * rows_per_rank)) * g.N + (size_t)tn * BN;
#pragma unroll 1
      for (int hh = 0; hh < BN / 128; ++hh) {
        const int half = ch * (BN / 128) + hh;
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * BN + half * 64;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3), c = lane & 7;
          const uint4 q = *reinterpret_cast<const uint4*>(myepi + r * 128 + ((c ^ (r & 7)) * 16));
          st16(reinterpret_cast<char*>(dst_base + (size_t)r * g.N + half * 64) + c * 16, q);
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) {
        if (leader) mbar_arrive(&acc_empty[buf]);
        else mbar_arrive_cluster(map_to_cta(smem_u32(&acc_empty[buf]), 0));
      }
    }
  }

  comm_sync(dc, pt, ticket, 1, true);

  // ---- owner-side reduction of the 128 x 256 tiles this CTA produced and owns -----------------------------------------
  {
    const __nv_bfloat16* stage_me = reinterpret_cast<const __nv_bfloat16*>(pt.send[me]);
    const size_t slot = (size_t)rows_per_rank * g.N;
    for (int it = 0; it < my_tiles; ++it) {
      const int sidx = cluster_id + it * nclusters;
      const int tm = (sidx / tiles_n) * 2 + (int)cta, tn = sidx % tiles_n;
      const int row0 = tm * BM;
      if (row0 / rows_per_rank != me) continue;
      const int lrow0 = row0 - me * rows_per_rank;
      constexpr int kU = 4, kItems = BM * (BN / 8);
      for (int idx0 = threadIdx.x; idx0 < kItems; idx0 += kThreads * kU) {
        float acc[kU][8];
        size_t off[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int idx = idx0 + u * kThreads;
          const int r = idx / (BN / 8), c8 = idx % (BN / 8);
          off[u] = (size_t)(lrow0 + r) * g.N + (size_t)tn * BN + c8 * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[u][j] = 0.f;
        }
        for (int q = 0; q < P; ++q) {
          uint4 v[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u)
            if (idx0 + u * kThreads < kItems) v[u] = __ldcg(reinterpret_cast<const uint4*>(stage_me + (size_t)q * slot + off[u]));
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[u][2 * j] += __uint_as_float(w[j] << 16);
              acc[u][2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          if (idx0 + u * kThreads >= kItems) continue;
          if (g.out_fp32) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + off[u]);
            o[0] = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
            o[1] = make_float4(acc[u][4], acc[u][5], acc[u][6], acc[u][7]);
          } else {
            uint4 o;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[u][0], acc[u][1]), h1 = __floats2bfloat162_rn(acc[u][2], acc[u][3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[u][4], acc[u][5]), h3 = __floats2bfloat162_rn(acc[u][6], acc[u][7]);
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + off[u]) = o;
          }
        }
      }
    }
  }
  comm_sync(dc, pt, ticket, 2, false);

  // nobody leaves (or frees tensor memory) while the pair's other half may still read this CTA's operands or barriers
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
}  // namespace

size_t gemm_rs_stage_bytes(int M, int N) { return (size_t)M * N * 2; }   // P slots of [M/P, N] bf16

int gemm_rs_channels(int M, int N, int max_channels) {
  int tiles = (M / BM) * (N / BN);
  return std::max(1, std::min(tiles, max_channels));
}

const char* gemm_rs_check(int M, int N, int K, int P) {
  if (M <= 0 || N <= 0 || K <= 0) return "empty problem";
  if (M % (BM * P) != 0) return "M must be a multiple of 128 * group size";
  if (N % BN != 0) return "N must be a multiple of 256";
  if (K % BK != 0) return "K must be a multiple of 64";
  return nullptr;
}

// ---- 2-CTA variant: host side ----
const char* gemm_rs2_check(int M, int N, int K, int P) {
  if (const char* why = gemm_rs_check(M, N, K, P)) return why;
  if (M % (2 * BM) != 0) return "M must be a multiple of 256";
  return nullptr;
}

int gemm_rs2_channels(int M, int N, int max_channels) {
  int nsuper = (M / (2 * BM)) * (N / BN);
  return 2 * std::max(1, std::min(nsuper, max_channels / 2));
}

cudaError_t launch_gemm_rs2(const DevComm& dc, const void* a, const void* w, unsigned long long stage_off, void* out,
                            bool out_fp32, int M, int N, int K, int channels, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_rs2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes2);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap ma, mw;
  if (!make_map(&ma, a, M, K, BM) || !make_map(&mw, w, N, K, BN2H)) return cudaErrorInvalidValue;
  GemmRsArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.stage_off = stage_off;
  g.out = out;
  g.out_fp32 = out_fp32 ? 1 : 0;
  k_gemm_rs2<<<channels, kThreads, kSmemBytes2, s>>>(dc, ma, mw, g);   // __cluster_dims__(2,1,1): channels is even
  return cudaGetLastError();
}

cudaError_t launch_gemm_rs(const DevComm& dc, const void* a, const void* w, unsigned long long stage_off, void* out,
                           bool out_fp32, int M, int N, int K, int channels, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_rs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap ma, mw;
  if (!make_map(&ma, a, M, K, BM) || !make_map(&mw, w, N, K, BN)) return cudaErrorInvalidValue;
  GemmRsArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.stage_off = stage_off;
  g.out = out;
  g.out_fp32 = out_fp32 ? 1 : 0;
  k_gemm_rs<<<channels, kThreads, kSmemBytes, s>>>(dc, ma, mw, g);
  return cudaGetLastError();
}

}  // namespace mlslb
