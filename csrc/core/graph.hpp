// DL graph layer: Session -> Operation -> {Activation, ParameterSet}, Distribution, Statistics.
//
// Same semantics as the reference's L4 (reference src/mlsl_impl.hpp / src/mlsl_impl.cpp): from layer shapes and a
// (data x model) partition derive WHICH collective every activation / gradient / parameter needs, size and own the
// comm buffers, and expose pack/unpack block descriptors plus Start/Wait/Test per tensor.  The requests it builds
// are persistent CommRequests executed by the active backend (CUDA peer-memory kernels or host shared memory).
#pragma once
#include <map>
#include <tuple>
#include <string>
#include <vector>

#include "../../include/mlsl.hpp"
#include "runtime.hpp"

namespace MLSL {
namespace impl {

using mlslb::CommDesc;
using mlslb::CommRequest;
using mlslb::ProcessGroup;
using mlslb::RankContext;

inline mlslb::DType to_dtype(DataType d) {
  switch (d) {
    case DT_FLOAT: return mlslb::DType::F32;
    case DT_DOUBLE: return mlslb::DType::F64;
    case DT_BYTE: return mlslb::DType::U8;
    case DT_BF16: return mlslb::DType::BF16;
    case DT_FP16: return mlslb::DType::F16;
    case DT_INT32: return mlslb::DType::I32;
  }
  return mlslb::DType::F32;
}
inline mlslb::RedOp to_redop(ReductionType r) {
  switch (r) {
    case RT_SUM: return mlslb::RedOp::SUM;
    case RT_MIN: return mlslb::RedOp::MIN;
    case RT_MAX: return mlslb::RedOp::MAX;
  }
  return mlslb::RedOp::SUM;
}

class SessionImpl;
class OperationImpl;
class DistributionImpl;
class StatisticsImpl;

// Lazily allocated comm buffer from the symmetric heap (reference src/mlsl_impl.hpp:129-172).
struct CommBuf {
  RankContext* ctx = nullptr;
  size_t bytes = 0;
  void* ptr = nullptr;
  void allocate();
  void release();
};

class BlockImpl : public CommBlockInfo {
 public:
  BlockImpl(size_t mbOff, size_t mbCnt, size_t fmOff, size_t fmCnt, size_t fmSz, DataType dt, size_t bufOff)
      : mbOffset(mbOff), mbCount(mbCnt), fmOffset(fmOff), fmCount(fmCnt), fmSize(fmSz), dataType(dt), bufOffset(bufOff) {}
  size_t mbOffset, mbCount, fmOffset, fmCount, fmSize;
  DataType dataType;
  size_t bufOffset;
};

class ActivationImpl : public Activation {
 public:
  ActivationImpl(OperationImpl* op, size_t fmCount, size_t fmSize, DataType dt, bool isInput, size_t index);
  ~ActivationImpl();
  void set_peer(ActivationImpl* other);
  void connect();                 // called on OUTPUT activations at Commit: choose the exchange pattern
  void start(void* buf);
  void start_fused(void* local, void* dst);   // [ext] exchange straight from / into the unpacked tensors
  void* wait();
  void pack(const void* local, void* comm, bool unpack);
  std::string describe() const;

  OperationImpl* op;
  DistributionImpl* dist;
  bool isInput;
  size_t index;
  size_t globalFmCount, globalFmOffset, localFmCount, fmSize;
  DataType dataType;
  bool needReduce = false, needComm = false, peerSet = false;
  ActivationImpl* peer = nullptr;
  CommRequest* req = nullptr;
  size_t sendRegionBytes = 0;     // start(buf) receives into buf + sendRegionBytes when the op is out of place
  CommBuf commBuf;
  std::vector<BlockImpl*> packBlocks, unpackBlocks;
  int commCase = 0;
  void* fusedDst = nullptr;        // start_fused(local, dst): the kernel wrote the consumer's tensor itself, WaitComm returns it
  void* unpackDst = nullptr;       // start_fused fallback: WaitComm unpacks into it after the collective
  size_t msg_bytes() const;
};

class ParameterSetImpl : public ParameterSet {
 public:
  ParameterSetImpl(OperationImpl* op, size_t kernelCount, size_t kernelSize, DataType dt, bool distUpdate,
                   CompressionType ct, size_t index);
  ~ParameterSetImpl();
  void start_gradient(void* buf);
  void* wait_gradient();
  void* test_gradient(bool* done);
  void start_increment(void* buf);
  void* wait_increment();
  void start_fused(void* grad, void* param, DataType paramType, void* master, void* s1, void* s2,
                   const FusedUpdateParams* opt);
  void wait_fused();
  std::string describe() const;

  OperationImpl* op;
  DistributionImpl* dist;
  size_t index;
  size_t globalKernelCount, globalKernelOffset, localKernelCount, ownedKernelCount, ownedKernelOffset, kernelSize;
  DataType dataType;
  bool distributedUpdate, needComm;
  CompressionType compress;
  CommRequest* gradReq = nullptr;
  CommRequest* incReq = nullptr;
  CommRequest* fusedReq = nullptr;
  CommBuf commBuf;
  size_t grad_msg_bytes() const;
  size_t inc_msg_bytes() const;
};

class WindowImpl : public Window {
 public:
  DistributionImpl* dist = nullptr;
  GroupType groupType = GT_GLOBAL;
  ProcessGroup* group = nullptr;
  std::vector<uint64_t> offsets, sizes;   // per group member: heap offset and size of the exposed memory
};

class DistributionImpl : public Distribution {
 public:
  DistributionImpl(RankContext* ctx, size_t dataParts, size_t modelParts, bool replicate, int dataColor,
                   int modelColor);
  DistributionImpl(RankContext* ctx, ProcessGroup* data);   // `data` as the data group, no model parallelism
  ~DistributionImpl();
  ProcessGroup* group(GroupType gt);
  CommRequest* make_request(mlslb::OpKind kind, DataType dt, GroupType gt);
  CommReq* submit(CommRequest* r, void* send, void* recv);

  RankContext* ctx;
  size_t dataParts, modelParts, replicaCount;
  ProcessGroup* dataGroup = nullptr;
  ProcessGroup* modelGroup = nullptr;
  ProcessGroup* replicaGroup = nullptr;
  // Compressed (fp8) all-reduces of the Distribution API keep their request - and with it the error-feedback residual -
  // per (buffers, count, reduction, group), as the reference keys its residual by the buffer address (quant/quant.c:153-167);
  // a fresh one-shot request per call would quantise without ever compensating the error.
  struct CompressKey {
    void* send;
    void* recv;
    size_t count;
    int dt, rt, gt;
    bool operator<(const CompressKey& o) const {
      return std::tie(send, recv, count, dt, rt, gt) < std::tie(o.send, o.recv, o.count, o.dt, o.rt, o.gt);
    }
  };
  std::map<CompressKey, CommRequest*> compressed;
};

struct RegTensor {
  size_t count, size;
  DataType dtype;
  bool distUpdate = false;
  CompressionType compress = CT_NONE;
};

class OperationRegInfoImpl : public OperationRegInfo {
 public:
  explicit OperationRegInfoImpl(OpType t) : opType(t) {}
  OpType opType;
  std::string name;
  std::vector<RegTensor> inputs, outputs, params;
  int refs = 1;   // session handle + one per Operation built from it
};

class OperationImpl : public Operation {
 public:
  OperationImpl(SessionImpl* s, OperationRegInfoImpl* info, DistributionImpl* d, size_t index);
  ~OperationImpl();
  void bind(DistributionImpl* d);
  void commit();
  SessionImpl* session;
  OperationRegInfoImpl* info;
  DistributionImpl* dist = nullptr;
  size_t opIndex;
  int64_t uid;
  OpType opType;
  std::string name;
  size_t localMb = 0, mbOffset = 0;
  std::vector<ActivationImpl*> inputs, outputs;
  std::vector<ParameterSetImpl*> params;
};

// Per-entity counters; entity order inside an operation: inputs, outputs, then (grad, inc) per parameter set
// (the reference's statIdx scheme, src/mlsl_impl_stats.cpp:564-668).
struct EntityStat {
  unsigned long long commCycles = 0, computeCycles = 0, isolationCycles = 0;
  unsigned long long commNs = 0, computeNs = 0;
  unsigned long long devCommNs = 0, devRuns = 0, isolationDevNs = 0;   // device-timed duration of the collective itself
  size_t commBytes = 0, bytesPerIter = 0;
};
struct OpStat {
  std::vector<EntityStat> ent;
};

class StatisticsImpl : public Statistics {
 public:
  explicit StatisticsImpl(SessionImpl* s);
  enum Action { START = 0, WAIT = 1, TEST = 2 };
  enum Kind { INPUT_ACT = 0, OUTPUT_ACT = 1, PARAM_GRAD = 2, PARAM_INC = 3 };
  void initialize();             // size the tables (Commit)
  void collect_isolation();      // timed dry runs of every communication (Commit, when enabled)
  // Bracket an API call: enter() books the time since the previous MLSL call as compute, leave() books the
  // call itself as communication.
  void enter(size_t opIdx, Kind k, size_t entIdx, Action a);
  void leave(size_t opIdx, Kind k, size_t entIdx, Action a, mlslb::CommRequest* req = nullptr);
  static double cycles_per_ns();
  void start();
  void stop();
  void reset();
  void print();
  size_t slot(size_t opIdx, Kind k, size_t entIdx) const;
  SessionImpl* session;
  bool enabled = false, started = false, collecting = false;
  std::vector<OpStat> ops;
  unsigned long long lastCycles = 0, lastNs = 0;
  unsigned long long batches = 0;
};

class SessionImpl : public Session {
 public:
  SessionImpl(RankContext* ctx, PhaseType pt);
  ~SessionImpl();
  RankContext* ctx;
  PhaseType phase;
  size_t globalMb = 0;
  bool committed = false;
  std::vector<OperationImpl*> ops;
  StatisticsImpl* stats;
};

class EnvironmentImpl : public Environment {
 public:
  explicit EnvironmentImpl(RankContext* c) : ctx(c) {}
  RankContext* ctx;
  QuantParams* quantView = nullptr;     // heap copy handed back by GetQuantizationParams
  std::string waitMode, backendDesc;
};

EnvironmentImpl* env_of(RankContext* ctx);

}  // namespace impl
}  // namespace MLSL
