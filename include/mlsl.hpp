// mlsl-b200: public C++ API.
//
// Source-compatible with the public surface of Intel(R) MLSL 2018 (API 1.0; reference include/mlsl.hpp:28-913):
// the same namespace, class names, method signatures, enums and version macros, so a framework integration written
// against the reference recompiles unchanged.  Every pointer handed to a collective may be host memory (host
// backend) or device memory (CUDA backend: buffers from Environment::Alloc live in the symmetric, peer-mapped
// device heap and are used zero-copy by the NVLink kernels; any other device pointer is staged transparently).
//
// Blackwell-side extensions are marked [ext]: extra data types, stream binding, fused epilogues (scale / cast /
// fp8-compressed transport), the fused distributed-update op and send/recv lists.
#ifndef MLSL_HPP
#define MLSL_HPP

#include <cstddef>
#include <cstdint>

#define MLSL_MAJOR_VERSION 1
#define MLSL_MINOR_VERSION 0

#define MLSL_VERSION(major, minor) (((major) << 16) | (minor))
#define MLSL_MAJOR(version) ((version) >> 16)
#define MLSL_MINOR(version) ((version)&0xFFFF)
#define MLSL_VERSION_GE(v1, v2)                                                               \
  ((MLSL_MAJOR(v1) > MLSL_MAJOR(v2)) ||                                                       \
   (MLSL_MAJOR(v1) == MLSL_MAJOR(v2) && MLSL_MINOR(v1) >= MLSL_MINOR(v2)))
#define MLSL_VERSION_LT(v1, v2)                                                               \
  ((MLSL_MAJOR(v1) < MLSL_MAJOR(v2)) ||                                                       \
   (MLSL_MAJOR(v1) == MLSL_MAJOR(v2) && MLSL_MINOR(v1) < MLSL_MINOR(v2)))

// Objects are created and destroyed only through their factories (Environment / Session).
#define MLSL_FACTORY_ONLY(T) \
 protected:                  \
  T() {}                     \
  ~T() {}                    \
                             \
 private:                    \
  T(const T&);               \
  T& operator=(const T&);

namespace MLSL {

typedef int CommReq;   // opaque handle of an in-flight collective

enum DataType {
  DT_FLOAT = 0,
  DT_DOUBLE = 1,
  DT_BYTE = 2,
  DT_BF16 = 3,   // [ext]
  DT_FP16 = 4,   // [ext]
  DT_INT32 = 5   // [ext]
};

enum PhaseType { PT_TRAIN = 0, PT_TEST = 1 };

// GT_DATA: ranks holding the same parameter shard for different samples (gradient exchange).
// GT_MODEL: ranks holding different shards for the same samples (activation exchange).  GT_GLOBAL: everyone.
enum GroupType { GT_DATA = 0, GT_MODEL = 1, GT_GLOBAL = 2 };

enum ReductionType { RT_SUM = 0, RT_MIN = 1, RT_MAX = 2 };

enum OpType {
  OT_CC = 0,      // cross-correlation / GEMM-like: inputs and outputs independent, has parameters
  OT_BIAS = 1,    // same in/out layout, has parameters
  OT_ACT = 2,     // same in/out layout, no parameters
  OT_POOL = 3,    // same in/out layout, no parameters
  OT_SPLIT = 4,   // output depends on input (OA1+OA2+...), no parameters
  OT_CONCAT = 5,  // output is IA1+IA2+..., no parameters
  OT_BCAST = 6,   // OA1=IA, OA2=IA, ...
  OT_REDUCE = 7,  // OA=IA1+IA2+...
  OT_DATA = 8,    // outputs only
  OT_EVAL = 9     // inputs only
};

enum CompressionType { CT_NONE = 0, CT_QUANTIZATION = 1 };

// The reference dlopen()s a user quantisation library described by this struct (three function names + block
// geometry).  The host backend does the same when lib_path is set (csrc/tests/quant_plugin_sample.c is a complete
// plug-in); the CUDA backend always uses its built-in block-scaled FP8 E4M3 format fused into the all-reduce kernel.
// With an empty lib_path the built-in format is used everywhere; block_size / elem_in_block then report 132 / 128.
typedef struct {
  char* lib_path;
  char* quant_buffer_func_name;
  char* dequant_buffer_func_name;
  char* reduce_sum_func_name;
  size_t block_size;
  size_t elem_in_block;
} QuantParams;

// [ext] optimizer description for ParameterSet::StartFusedUpdate
enum OptimizerType { OPT_SGD = 0, OPT_ADAMW = 1 };
typedef struct {
  OptimizerType type;
  float lr, momentum, beta1, beta2, eps, weight_decay;
  long long step;      // 1-based step count (AdamW bias correction)
  float grad_scale;    // multiplier applied to the summed gradient (e.g. 1/world)
} FusedUpdateParams;

// One rectangular piece of an activation in the (minibatch, feature-map, feature-map-size) index space together
// with its offset (in elements) inside the communication buffer.
class CommBlockInfo {
  MLSL_FACTORY_ONLY(CommBlockInfo)
 public:
  size_t GetMbOffset();
  size_t GetMbCount();
  size_t GetFmOffset();
  size_t GetFmCount();
  size_t GetFmSize();
  DataType GetDataType();
  size_t GetBufOffset();
};

class Activation {
  MLSL_FACTORY_ONLY(Activation)
 public:
  size_t GetGlobalFmCount();
  size_t GetGlobalFmOffset();
  size_t GetLocalFmCount();
  size_t GetPackBlockCount();
  size_t GetUnpackBlockCount();
  CommBlockInfo* GetPackBlock(size_t idx);
  CommBlockInfo* GetUnpackBlock(size_t idx);
  DataType GetDataType();
  size_t GetFmSize();
  void* GetCommBuf();
  size_t GetCommBufSize();
  void StartComm(void* buf);
  void* WaitComm();   // waits for the PEER activation's transfer; NULL when no communication is needed
  // [ext] pack + exchange + unpack in one call: `localBuf` is the UNPACKED local tensor; with `localDst` the consumer's
  // unpacked tensor is filled directly and the peer's WaitComm returns it.  For the all-to-all patterns on the CUDA backend
  // this is one kernel that pulls the rectangles straight out of the peers' tensors (no pack / unpack kernels).
  void StartCommFused(void* localBuf, void* localDst = nullptr);
  // [ext] device-side pack/unpack of a local (mb, fm, fmSize) tensor to/from the comm buffer (one kernel)
  void Pack(const void* localBuf, void* commBuf);
  void Unpack(const void* commBuf, void* localBuf);
};

class ParameterSet {
  MLSL_FACTORY_ONLY(ParameterSet)
 public:
  size_t GetGlobalKernelCount();
  size_t GetGlobalKernelOffset();
  size_t GetLocalKernelCount();
  size_t GetOwnedKernelCount();
  size_t GetOwnedKernelOffset();
  DataType GetDataType();
  size_t GetKernelSize();
  bool IsDistributedUpdate();
  void StartGradientComm(void* buf);
  void StartIncrementComm(void* buf);
  void* WaitGradientComm();
  void* TestGradientComm(bool* isCompleted);
  void* WaitIncrementComm();
  // [ext] reduce-scatter + optimizer step on the owned shard + all-gather of the new parameters as ONE
  // operation (one kernel on the CUDA backend).  grad: LocalKernelCount*KernelSize elements of the set's data
  // type; param: same count of paramType; state buffers cover the owned shard only (fp32).
  void StartFusedUpdate(void* grad, void* param, DataType paramType, void* master, void* state1, void* state2,
                        const FusedUpdateParams* opt);
  void WaitFusedUpdate();
  // [ext] multiplier fused into the gradient reduction (e.g. 1/world for averaging); default 1
  void SetGradientScale(float scale);
};

class Window;

class Distribution {
  MLSL_FACTORY_ONLY(Distribution)
 public:
  size_t GetProcessIdx(GroupType groupType);
  size_t GetProcessCount(GroupType groupType);
  CommReq* Bcast(void* buffer, size_t count, DataType dataType, size_t rootIdx, GroupType groupType);
  CommReq* Reduce(void* sendBuffer, void* recvBuffer, size_t count, DataType dataType, ReductionType redType,
                  size_t rootIdx, GroupType groupType);
  CommReq* AllReduce(void* sendBuffer, void* recvBuffer, size_t count, DataType dataType, ReductionType redType,
                     GroupType groupType);
  CommReq* AlltoAll(void* sendBuffer, size_t sendCount, void* recvBuffer, DataType dataType, GroupType groupType);
  CommReq* AlltoAllv(void* sendBuffer, size_t* sendCounts, size_t* sendOffsets, void* recvBuffer,
                     size_t* recvCounts, size_t* recvOffsets, DataType dataType, GroupType groupType);
  CommReq* Gather(void* sendBuffer, size_t sendCount, void* recvBuffer, DataType dataType, size_t rootIdx,
                  GroupType groupType);
  CommReq* AllGather(void* sendBuffer, size_t sendCount, void* recvBuffer, DataType dataType, GroupType groupType);
  CommReq* AllGatherv(void* sendBuffer, size_t sendCount, void* recvBuffer, size_t* recvCounts, DataType dataType,
                      GroupType groupType);
  CommReq* Scatter(void* sendBuffer, void* recvBuffer, size_t recvCount, DataType dataType, size_t rootIdx,
                   GroupType groupType);
  CommReq* ReduceScatter(void* sendBuffer, void* recvBuffer, size_t recvCount, DataType dataType,
                         ReductionType redType, GroupType groupType);
  void Barrier(GroupType groupType);
  // [ext] all-reduce with the epilogue fused into the reduction kernel: result = scale * sum, optional fp8
  // block-quantised transport (compress) - no separate elementwise kernel runs.
  CommReq* AllReduceEx(void* sendBuffer, void* recvBuffer, size_t count, DataType dataType, ReductionType redType,
                       GroupType groupType, float scale, CompressionType compress);
  CommReq* ReduceScatterEx(void* sendBuffer, void* recvBuffer, size_t recvCount, DataType dataType,
                           ReductionType redType, GroupType groupType, float scale);
  // [ext] sparse point-to-point list (ring shifts, halo / KV rotation): entry p describes what goes to / comes
  // from group member p (count 0 = nothing).  The reference declares this op (src/comm.hpp:212-248) but never
  // exposes it.
  CommReq* SendRecvList(void* sendBuffer, size_t* sendCounts, size_t* sendOffsets, void* recvBuffer,
                        size_t* recvCounts, size_t* recvOffsets, DataType dataType, GroupType groupType);
  // [ext] tensor-parallel GEMM fused with the reduce-scatter of its partial sums (CUDA backend, one tcgen05 kernel):
  // out[M/P, N] (rows of this rank) = sum over the group of A_r[M, K] * W_r[N, K]^T; A, W bf16 row-major;
  // out bf16 or fp32 (outType).  M % (128 * P) == 0, N % 256 == 0, K % 64 == 0.
  CommReq* GemmReduceScatter(const void* a, const void* w, void* out, size_t M, size_t N, size_t K, DataType outType,
                             GroupType groupType);
  // [ext, experimental] all-gather fused with the GEMM that consumes it (CUDA backend, one kernel: copy CTAs stream the
  // peers' row shards into `gathered` while tensor-core CTAs already multiply the rows that have landed):
  // out[M, N] = concat_rows(X_0 .. X_{P-1}) * W[N, K]^T, xShard = X_r [M/P, K] bf16, gathered [M, K] bf16 (kept for
  // backward), out bf16 or fp32.  Same shape rules as GemmReduceScatter.
  CommReq* AllGatherGemm(const void* xShard, const void* w, void* gathered, void* out, size_t M, size_t N, size_t K,
                         DataType outType, GroupType groupType);
  // [ext] one-sided access (the reference keeps its RMA window table behind ENABLE_MPIRMA_ENDPOINTS,
  // eplib/window.c).  Collective over the group: every member exposes `bytes` at `base` (memory from
  // Environment::Alloc).  FreeWindow is collective too.
  Window* CreateWindow(void* base, size_t bytes, GroupType groupType);
  void FreeWindow(Window* window);
};

// [ext] RMA window: Put / Get address a member by its index in the window's group and a byte displacement inside
// the memory it exposed; they are ordered like the caller's other work (stream order on the CUDA backend) and
// complete - locally and at the target - at the next Fence(), which is collective.
class Window {
  MLSL_FACTORY_ONLY(Window)
 public:
  void Put(const void* origin, size_t bytes, size_t targetIdx, size_t targetDisp);
  void Get(void* origin, size_t bytes, size_t targetIdx, size_t targetDisp);
  void Fence();
  size_t GetSize(size_t memberIdx);
};

class OperationRegInfo {
  MLSL_FACTORY_ONLY(OperationRegInfo)
 public:
  void SetName(const char* name);
  size_t AddInput(size_t featureMapCount, size_t featureMapSize, DataType dataType);
  size_t AddOutput(size_t featureMapCount, size_t featureMapSize, DataType dataType);
  size_t AddParameterSet(size_t kernelCount, size_t kernelSize, DataType dataType, bool distributedUpdate = false,
                         CompressionType compressType = CT_NONE);
  void Validate(Distribution* dist = NULL);
};

class Session;

class Operation {
  MLSL_FACTORY_ONLY(Operation)
 public:
  void SetDistribution(Distribution* dist);
  Distribution* GetDistribution();
  Session* GetSession();
  OpType GetOpType();
  void SetPrev(Operation* prev, size_t actIdx, size_t prevOpActIdx);
  void SetNext(Operation* next, size_t actIdx, size_t nextOpActIdx);
  const char* GetName();
  size_t GetGlobalMinibatchSize();
  size_t GetLocalMinibatchSize();
  size_t GetGlobalMinibatchOffset();
  size_t GetInputCount();
  Activation* GetInput(size_t idx);
  size_t GetOutputCount();
  Activation* GetOutput(size_t idx);
  bool HasParameterSets();
  size_t GetParameterSetCount();
  ParameterSet* GetParameterSet(size_t idx);
};

class Statistics {
  MLSL_FACTORY_ONLY(Statistics)
 public:
  void Start();
  void Stop();
  void Reset();
  bool IsStarted();
  bool IsEnabled();
  void Print();
  unsigned long long GetIsolationCommCycles(size_t opIdx);
  size_t GetCommSize(size_t opIdx);
  unsigned long long GetCommCycles(size_t opIdx);
  unsigned long long GetComputeCycles(size_t opIdx);
  unsigned long long GetTotalIsolationCommCycles();
  size_t GetTotalCommSize();
  unsigned long long GetTotalCommCycles();
  unsigned long long GetTotalComputeCycles();
  // [ext] same counters in nanoseconds (cycle counters are TSC ticks as in the reference)
  unsigned long long GetCommNanos(size_t opIdx);
  unsigned long long GetComputeNanos(size_t opIdx);
  // [ext] duration of the operation's collectives measured ON THE DEVICE (CUDA event pair around each kernel); 0 on
  // host backends.  With stream-ordered waits this is also what GetCommCycles / GetCommNanos carry.
  unsigned long long GetDeviceCommNanos(size_t opIdx);
};

class Session {
  MLSL_FACTORY_ONLY(Session)
 public:
  void SetGlobalMinibatchSize(size_t globalMinibatchSize);
  size_t GetGlobalMinibatchSize();
  PhaseType GetPhaseType();
  OperationRegInfo* CreateOperationRegInfo(OpType opType);
  void DeleteOperationRegInfo(OperationRegInfo* info);
  size_t AddOperation(OperationRegInfo* info, Distribution* dist = NULL);
  void RemoveOperations();
  size_t GetOperationCount();
  Operation* GetOperation(size_t idx);
  void Commit();
  Statistics* GetStats();
};

class Environment {
  MLSL_FACTORY_ONLY(Environment)
 public:
  static Environment& GetEnv();
  static int GetVersion();
  void Configure(const char* config = NULL);
  void Init(int* argc, char** argv[]);
  void Finalize();
  bool IsInitialized();
  size_t GetProcessIdx();
  size_t GetProcessCount();
  Session* CreateSession(PhaseType phaseType = PT_TRAIN);
  void DeleteSession(Session* session);
  Distribution* CreateDistribution(size_t dataPartitions, size_t modelPartitions);
  Distribution* CreateDistributionWithColors(int dataColor, int modelColor);
  void DeleteDistribution(Distribution* distribution);
  void Wait(CommReq* req);
  void Test(CommReq* req, bool* isCompleted);
  void* Alloc(size_t size, size_t alignment);
  void Free(void* ptr);
  void SetQuantizationParams(QuantParams* params);
  QuantParams* GetQuantizationParams();
  // [ext] --------------------------------------------------------------------------------------------------
  // CUDA stream the caller computes on (cudaStream_t).  Start*() orders the collective after the work already
  // submitted to it; Wait*() either blocks the host (default) or - WaitMode "stream" - only orders the stream.
  void SetStream(void* cudaStream);
  void* GetStream();
  void SetWaitMode(const char* mode);   // "host" | "stream"
  // [ext] device-path tuning knobs by name ("ar_channels", "mid_max_kb", ... or their MLSL_* environment names; the
  // list is printed at MLSL_LOG_LEVEL=1).  Every rank must apply the same change at the same point of the program.
  // [ext] operation uids of the collectives the progress threads launched most recently, oldest first (what order did
  // message prioritisation choose?).  Returns the number written (<= capacity).
  size_t GetLaunchOrder(long long* uids, size_t capacity);
  void SetTuning(const char* key, long value);
  long GetTuning(const char* key);
  const char* GetBackendName();         // "host" | "cuda"
  const char* DescribeBackend();        // human-readable: device, heap kind, NVLS availability
  bool IsDeviceBackend();
  // Park / resume the background progress threads (reference EPLIB_suspend / EPLIB_execute).
  void SuspendServers();
  void ResumeServers();
  // Group creation by its members only, for hosts that bring their own rendezvous (the torch.distributed backend
  // uses the job's store): every member reads GetGroupState(), the members exchange the two words, and each calls
  // CreateDistributionFromRanks with the OR of the row bitmaps and the maximum of the ticket marks.  `ranks` are
  // process indices of the global group in group order; the new distribution has them as its data group.
  void GetGroupState(unsigned long long* rowsInUse, unsigned long long* ticketMark);
  Distribution* CreateDistributionFromRanks(const size_t* ranks, size_t count, unsigned long long rowsInUse,
                                            unsigned long long ticketMark);
};

}  // namespace MLSL

#endif /* MLSL_HPP */
