#include "graph.hpp"

#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "log.hpp"
#include "quant.hpp"
#include "sysinfo.hpp"
#include "version.hpp"

namespace MLSL {
namespace impl {

using namespace mlslb;

static const size_t kMaxCount = (size_t)1 << 40;   // sanity bound; the reference is limited to INT_MAX elements

// ============================================================================================================
// CommBuf
// ============================================================================================================
void CommBuf::allocate() {
  if (ptr || bytes == 0) return;
  ptr = ctx->backend->alloc(bytes, 64);
}
void CommBuf::release() {
  if (ptr) ctx->backend->free(ptr);
  ptr = nullptr;
}

// ============================================================================================================
// Distribution
// ============================================================================================================
// Group algebra (reference src/mlsl_impl.hpp:212-266): with D data parts and M model parts, L = D*M consecutive
// ranks form one replica; inside it M consecutive ranks form a model group and stride-M ranks form a data group.
DistributionImpl::DistributionImpl(RankContext* c, size_t dParts, size_t mParts, bool replicate, int dataColor,
                                   int modelColor)
    : ctx(c) {
  ProcessGroup* glob = ctx->global_group;
  const size_t G = (size_t)glob->size();
  const size_t gi = (size_t)glob->idx;
  if (dataColor == -1 && modelColor == -1) {
    MLSLB_ASSERT(dParts > 0 && mParts > 0 && (long long)dParts > 0 && (long long)mParts > 0,
                 "numbers for data and model groups must be positive");
    MLSLB_ASSERT(dParts * mParts <= G, "dataPartitions(%zu) x modelPartitions(%zu) exceeds the process count (%zu)",
                 dParts, mParts, G);
    dataParts = dParts;
    modelParts = mParts;
    const size_t L = dParts * mParts;
    const size_t lid = gi % L, rep = gi / L;
    replicaCount = replicate ? G / L : 1;
    const int mColor = (int)(rep * L + lid / mParts);
    const int dColor = (int)(rep * L + lid % mParts);
    const int rColor = (int)lid;
    if (mParts == 1) modelGroup = ctx->self_group;
    else if (mParts == G) modelGroup = glob;
    else modelGroup = ctx->create_group_by_color(glob, mColor);
    if (dParts == 1) dataGroup = ctx->self_group;
    else if (dParts == G) dataGroup = ctx->env.dup_group ? ctx->create_group_by_color(glob, 1) : glob;
    else dataGroup = ctx->create_group_by_color(glob, dColor);
    if (replicaCount == 1) replicaGroup = ctx->self_group;
    else if (replicaCount == G) replicaGroup = glob;
    else replicaGroup = ctx->create_group_by_color(glob, rColor);
  } else {
    replicaCount = 1;
    modelGroup = ctx->create_group_by_color(glob, modelColor);
    dataGroup = ctx->create_group_by_color(glob, dataColor);
    replicaGroup = ctx->self_group;
    dataParts = (size_t)dataGroup->size();
    modelParts = (size_t)modelGroup->size();
  }
}

DistributionImpl::DistributionImpl(RankContext* c, ProcessGroup* data)
    : ctx(c), dataParts((size_t)data->size()), modelParts(1), replicaCount(1), dataGroup(data),
      modelGroup(c->self_group), replicaGroup(c->self_group) {}

DistributionImpl::~DistributionImpl() {
  for (auto& kv : compressed) delete kv.second;
  compressed.clear();
  auto drop = [&](ProcessGroup* g) {
    if (!(g && g != ctx->global_group && g != ctx->self_group && g != ctx->world_group)) return;
    try {   // destructors must not throw: a poisoned / timed-out job is reported, not escalated to terminate()
      ctx->free_group(g);
    } catch (const std::exception& e) {
      MLSLB_LOG(LOG_ERROR, "while releasing a process group: %s", e.what());
    }
  };
  drop(modelGroup);
  drop(dataGroup);
  drop(replicaGroup);
}

ProcessGroup* DistributionImpl::group(GroupType gt) {
  switch (gt) {
    case GT_DATA: return dataGroup;
    case GT_MODEL: return modelGroup;
    case GT_GLOBAL: return ctx->global_group;
  }
  MLSLB_ASSERT(false, "unexpected group type %d", (int)gt);
  return nullptr;
}

CommRequest* DistributionImpl::make_request(OpKind kind, DataType dt, GroupType gt) {
  CommRequest* r = new CommRequest(ctx, to_dtype(dt), -1, CommDesc::GENERIC);
  r->desc.kind = kind;
  r->desc.group = group(gt);
  r->one_shot = true;
  return r;
}

CommReq* DistributionImpl::submit(CommRequest* r, void* send, void* recv) {
  r->setup();
  ctx->register_request(r);
  r->start(send, recv);
  return (CommReq*)r;
}

// ============================================================================================================
// Activation
// ============================================================================================================
ActivationImpl::ActivationImpl(OperationImpl* o, size_t fmCount, size_t fmSz, DataType dt, bool in, size_t idx)
    : op(o), dist(o->dist), isInput(in), index(idx), globalFmCount(fmCount), fmSize(fmSz), dataType(dt) {
  const size_t M = (size_t)dist->modelGroup->size();
  // An OT_CC output is a full-width partial sum on every model rank; everything else is feature-map partitioned
  // (reference src/mlsl_impl.cpp:43-57).
  if (!isInput && op->opType == OT_CC) {
    localFmCount = globalFmCount;
    globalFmOffset = 0;
    needReduce = M > 1;
  } else {
    // the reference divides silently (src/mlsl_impl.cpp:49) and then exchanges the wrong elements; say so instead
    MLSLB_ASSERT(globalFmCount % M == 0, "%zu feature maps cannot be split over a model group of %zu ranks", globalFmCount, M);
    localFmCount = globalFmCount / M;
    globalFmOffset = localFmCount * (size_t)dist->modelGroup->idx;
    needReduce = false;
  }
  commBuf.ctx = o->session->ctx;
}

ActivationImpl::~ActivationImpl() {
  for (auto b : packBlocks) delete b;
  for (auto b : unpackBlocks) delete b;
  delete req;
  commBuf.release();
}

void ActivationImpl::set_peer(ActivationImpl* other) {
  if (!other) {
    peer = nullptr;
    peerSet = true;
    needComm = false;
    return;
  }
  MLSLB_ASSERT(other->globalFmCount * other->fmSize == globalFmCount * fmSize,
               "prev output activation size must match current input activation size");
  MLSLB_ASSERT(isInput != other->isInput, "input-output doesn't pair");
  MLSLB_ASSERT(dataType == other->dataType, "datatype must match");
  MLSLB_ASSERT(peer == nullptr || peer == other, "peer can be set only once");
  peer = other;
  other->peer = this;
  peerSet = other->peerSet = true;
}

static CommRequest* new_act_request(ActivationImpl* a, CommDesc::CompType ct, OpKind kind, size_t count,
                                    ProcessGroup* g) {
  CommRequest* r = new CommRequest(a->op->session->ctx, to_dtype(a->dataType), a->op->uid, ct);
  r->desc.kind = kind;
  r->desc.count = count;
  r->desc.group = g;
  r->setup();
  return r;
}

// Decide how the tensor travels between the producer's layout and the consumer's layout.  The five supported
// patterns are the reference's (src/mlsl_impl.cpp:155-228); the block lists describe, in (minibatch, feature
// map) coordinates, which rectangle of the local tensor goes to which offset of the comm buffer.
void ActivationImpl::connect() {
  if (isInput || !peer) return;
  ActivationImpl* out = this;
  ActivationImpl* in = peer;
  DistributionImpl* od = out->dist;
  DistributionImpl* id = in->dist;
  RankContext* ctx = op->session->ctx;
  if (!(ctx->global_group->size() > 1 && (out->needReduce || od != id))) return;
  out->needComm = in->needComm = true;
  const size_t oM = (size_t)od->modelGroup->size(), iM = (size_t)id->modelGroup->size();
  const size_t oD = (size_t)od->dataGroup->size(), iD = (size_t)id->dataGroup->size();
  const size_t oMb = out->op->localMb, iMb = in->op->localMb;
  const DataType dt = out->dataType;

  if (out->needReduce && od == id) {
    // case 1: partial sums reduced AND re-partitioned by feature map: ReduceScatter fwd / AllGather bwd
    const size_t M = oM;
    const size_t len = in->localFmCount * oMb * in->fmSize;
    out->commCase = in->commCase = 1;
    out->req = new_act_request(out, CommDesc::FPROP, OpKind::REDUCE_SCATTER, len, id->modelGroup);
    const size_t fmPer = out->localFmCount / M;
    for (size_t i = 0; i < M; ++i)
      out->packBlocks.push_back(new BlockImpl(0, oMb, i * fmPer, fmPer, out->fmSize, dt, i * oMb * fmPer * out->fmSize));
    in->unpackBlocks.push_back(new BlockImpl(0, iMb, 0, in->localFmCount, in->fmSize, dt, 0));
    in->req = new_act_request(in, CommDesc::BPROP, OpKind::ALLGATHER, len, id->modelGroup);
    in->packBlocks.push_back(new BlockImpl(0, iMb, 0, in->localFmCount, in->fmSize, dt,
                                           (size_t)id->modelGroup->idx * iMb * in->localFmCount * in->fmSize));
    for (size_t i = 0; i < M; ++i)
      out->unpackBlocks.push_back(new BlockImpl(0, oMb, i * fmPer, fmPer, out->fmSize, dt, i * oMb * fmPer * out->fmSize));
  } else if (out->needReduce && iM == 1 && oD == iD) {
    // case 2: consumer is pure data parallel over the same data groups: AllReduce fwd, nothing bwd
    out->commCase = in->commCase = 2;
    const size_t len = out->localFmCount * oMb * out->fmSize;
    out->req = new_act_request(out, CommDesc::FPROP, OpKind::ALLREDUCE, len, od->modelGroup);
    out->packBlocks.push_back(new BlockImpl(0, oMb, 0, out->localFmCount, out->fmSize, dt, 0));
    in->unpackBlocks.push_back(new BlockImpl(0, iMb, 0, in->localFmCount, in->fmSize, dt, 0));
    in->req = new_act_request(in, CommDesc::BPROP, OpKind::BARRIER, 0, ctx->self_group);   // no traffic backward
  } else if (out->needReduce && iM == 1 && iD % oD == 0 && iD == oM * oD) {
    // case 3: consumer spreads the SAME samples over data*model ranks: ReduceScatter splitting the minibatch
    out->commCase = in->commCase = 3;
    const size_t M = oM;
    const size_t len = in->localFmCount * iMb * in->fmSize;
    out->req = new_act_request(out, CommDesc::FPROP, OpKind::REDUCE_SCATTER, len, od->modelGroup);
    const size_t mbPer = oMb / M;
    for (size_t i = 0; i < M; ++i)
      out->packBlocks.push_back(new BlockImpl(i * mbPer, mbPer, 0, out->localFmCount, out->fmSize, dt,
                                              i * mbPer * out->localFmCount * out->fmSize));
    in->unpackBlocks.push_back(new BlockImpl(0, iMb, 0, in->localFmCount, in->fmSize, dt, 0));
    in->req = new_act_request(in, CommDesc::BPROP, OpKind::ALLGATHER, len, od->modelGroup);
    in->packBlocks.push_back(new BlockImpl(0, iMb, 0, in->localFmCount, in->fmSize, dt,
                                           (size_t)od->modelGroup->idx * iMb * in->localFmCount * in->fmSize));
    for (size_t i = 0; i < M; ++i)
      out->unpackBlocks.push_back(new BlockImpl(i * mbPer, mbPer, 0, out->localFmCount, out->fmSize, dt,
                                                i * mbPer * out->localFmCount * out->fmSize));
  } else if (!out->needReduce && (oM == 1 || iM == 1)) {
    // cases 4/5: layout change data-parallel <-> model-parallel without reduction: AlltoAll both ways, blocks are
    // (minibatch chunk) x (feature-map chunk) tiles
    out->commCase = in->commCase = (oM == 1) ? 4 : 5;
    ProcessGroup* g = (oM == 1) ? id->modelGroup : od->modelGroup;
    const size_t mb = std::min(oMb, iMb);
    const size_t fmElems = std::min(out->localFmCount * out->fmSize, in->localFmCount * in->fmSize);
    const size_t oFm = fmElems / out->fmSize, iFm = fmElems / in->fmSize;
    const size_t len = mb * fmElems;
    out->req = new_act_request(out, CommDesc::FPROP, OpKind::ALLTOALL, len, g);
    in->req = new_act_request(in, CommDesc::BPROP, OpKind::ALLTOALL, len, g);
    size_t k = 0;
    for (size_t i = 0; i < oMb; i += mb)
      for (size_t j = 0; j < out->localFmCount; j += oFm, ++k) {
        out->packBlocks.push_back(new BlockImpl(i, mb, j, oFm, out->fmSize, dt, k * len));
        out->unpackBlocks.push_back(new BlockImpl(i, mb, j, oFm, out->fmSize, dt, k * len));
      }
    MLSLB_ASSERT(k == (size_t)g->size(), "alltoall: block count (%zu) should equal the group size (%d)", k, g->size());
    k = 0;
    for (size_t i = 0; i < iMb; i += mb)
      for (size_t j = 0; j < in->localFmCount; j += iFm, ++k) {
        in->unpackBlocks.push_back(new BlockImpl(i, mb, j, iFm, in->fmSize, dt, k * len));
        in->packBlocks.push_back(new BlockImpl(i, mb, j, iFm, in->fmSize, dt, k * len));
      }
    MLSLB_ASSERT(k == (size_t)g->size(), "alltoall: block count (%zu) should equal the group size (%d)", k, g->size());
  } else {
    MLSLB_ASSERT(false, "this combination of producer/consumer distributions is not supported yet");
  }
  for (ActivationImpl* a : {out, in}) {
    a->sendRegionBytes = a->req->out_of_place_default() ? a->req->send_bytes() : 0;
    a->commBuf.bytes = a->req->buf_bytes();
  }
}

void ActivationImpl::start(void* buf) {
  StatisticsImpl* st = op->session->stats;
  auto kind = isInput ? StatisticsImpl::INPUT_ACT : StatisticsImpl::OUTPUT_ACT;
  st->enter(op->opIndex, kind, index, StatisticsImpl::START);
  if (needComm && req && req->desc.kind != OpKind::BARRIER)
    req->start(buf, req->out_of_place_default() ? (char*)buf + sendRegionBytes : buf);
  st->leave(op->opIndex, kind, index, StatisticsImpl::START, needComm ? req : nullptr);
}

// [ext] Activation::StartCommFused(local, dst): the pack / exchange / unpack sequence in one call.  For the all-to-all
// patterns (cases 4 and 5) on a backend with strided support this is ONE kernel: every member pulls the rectangle meant
// for it straight out of each peer's unpacked tensor and writes it straight into its own unpacked `dst` (or, without `dst`,
// packed into the receive region): no pack kernel, no unpack kernel, no packed send region.  Everywhere else
// it is Pack + StartComm, and WaitComm unpacks into `dst` before it returns.
void ActivationImpl::start_fused(void* local, void* dst) {
  fusedDst = unpackDst = nullptr;
  if (!needComm || !req || req->desc.kind == OpKind::BARRIER) return;
  RankContext* ctx = op->session->ctx;
  ProcessGroup* g = req->desc.group;
  const bool a2a = (commCase == 4 || commCase == 5) && req->desc.kind == OpKind::ALLTOALL && peer && g &&
                   packBlocks.size() == (size_t)g->size() && peer->unpackBlocks.size() == (size_t)g->size();
  if (!(a2a && ctx->backend->supports_strided_alltoall() && g->size() > 1)) {
    commBuf.allocate();
    pack(local, commBuf.ptr, false);
    unpackDst = dst;
    start(commBuf.ptr);
    return;
  }
  const size_t es = dtype_size(to_dtype(dataType));
  const BlockImpl* mine = packBlocks[(size_t)g->idx];      // the block every peer holds for ME has this geometry in ITS tensor
  CommDesc::Strided& sd = req->desc.strided;
  sd.on = true;
  sd.rows = mine->mbCount;
  sd.row_bytes = mine->fmCount * mine->fmSize * es;
  sd.src_off = (mine->mbOffset * localFmCount + mine->fmOffset) * mine->fmSize * es;
  sd.src_stride = localFmCount * mine->fmSize * es;
  sd.src_total = op->localMb * localFmCount * fmSize * es;
  sd.dst_direct = dst != nullptr;
  sd.dst_off.clear();
  if (dst) {
    sd.dst_stride = peer->localFmCount * peer->fmSize * es;
    sd.dst_total = peer->op->localMb * peer->localFmCount * peer->fmSize * es;
    for (BlockImpl* b : peer->unpackBlocks) sd.dst_off.push_back((b->mbOffset * peer->localFmCount + b->fmOffset) * b->fmSize * es);
  } else {
    commBuf.allocate();
  }
  fusedDst = dst;
  StatisticsImpl* st = op->session->stats;
  auto kind = isInput ? StatisticsImpl::INPUT_ACT : StatisticsImpl::OUTPUT_ACT;
  st->enter(op->opIndex, kind, index, StatisticsImpl::START);
  req->start(local, dst ? dst : (char*)commBuf.ptr + sendRegionBytes);
  st->leave(op->opIndex, kind, index, StatisticsImpl::START, req);
}

void* ActivationImpl::wait() {
  StatisticsImpl* st = op->session->stats;
  auto kind = isInput ? StatisticsImpl::INPUT_ACT : StatisticsImpl::OUTPUT_ACT;
  st->enter(op->opIndex, kind, index, StatisticsImpl::WAIT);
  void* ret = nullptr;
  // the data we consume was sent by the PEER activation (reference src/mlsl_impl.cpp:366-386)
  if (needComm && peer && peer->req && peer->req->desc.kind != OpKind::BARRIER) {
    ret = peer->req->wait();
    peer->req->desc.strided.on = false;             // the next plain StartComm of this request is a packed exchange again
    if (peer->fusedDst) {
      ret = peer->fusedDst;                         // the kernel wrote my unpacked tensor itself
    } else if (peer->unpackDst && ret) {
      pack(peer->unpackDst, ret, true);             // fallback of start_fused: unpack with MY block list
      ret = peer->unpackDst;
    }
    peer->fusedDst = peer->unpackDst = nullptr;
  }
  st->leave(op->opIndex, kind, index, StatisticsImpl::WAIT, needComm && peer ? peer->req : nullptr);
  return ret;
}

void ActivationImpl::pack(const void* local, void* comm, bool unpack) {
  auto& blocks = unpack ? unpackBlocks : packBlocks;
  if (blocks.empty()) return;
  std::vector<BlockDesc> d(blocks.size());
  for (size_t i = 0; i < blocks.size(); ++i)
    d[i] = {blocks[i]->mbOffset, blocks[i]->mbCount, blocks[i]->fmOffset, blocks[i]->fmCount, blocks[i]->fmSize,
            blocks[i]->bufOffset};
  // argument order follows the data flow: pack local->comm, unpack comm->local
  op->session->ctx->backend->pack_blocks(d.data(), d.size(), localFmCount, to_dtype(dataType), unpack ? comm : local,
                                         unpack ? const_cast<void*>(local) : comm, unpack);
}

size_t ActivationImpl::msg_bytes() const { return req ? req->msg_bytes() : 0; }

std::string ActivationImpl::describe() const {
  char buf[512];
  int n = snprintf(buf, sizeof(buf), "global_fm %zu local_fm %zu fm_off %zu fm_size %zu need_comm %d case %d op %s "
                   "pack_blocks %zu unpack_blocks %zu buf_bytes %zu", globalFmCount, localFmCount, globalFmOffset,
                   fmSize, (int)needComm, commCase, req ? opkind_name(req->desc.kind) : "-", packBlocks.size(),
                   unpackBlocks.size(), commBuf.bytes);
  return std::string(buf, (size_t)n);
}

// ============================================================================================================
// ParameterSet
// ============================================================================================================
ParameterSetImpl::ParameterSetImpl(OperationImpl* o, size_t kCount, size_t kSize, DataType dt, bool distUpd,
                                   CompressionType ct, size_t idx)
    : op(o), dist(o->dist), index(idx), globalKernelCount(kCount), kernelSize(kSize), dataType(dt),
      distributedUpdate(distUpd), compress(ct) {
  RankContext* ctx = o->session->ctx;
  const size_t M = (size_t)dist->modelGroup->size();
  const size_t D = (size_t)dist->dataGroup->size();
  MLSLB_ASSERT(globalKernelCount % M == 0, "%zu kernels cannot be split over a model group of %zu ranks", globalKernelCount, M);
  localKernelCount = globalKernelCount / M;
  globalKernelOffset = localKernelCount * (size_t)dist->modelGroup->idx;
  needComm = D > 1;
  if (distributedUpdate) {
    // owned = ceil(local / D); local is padded up so every data rank owns an equal share
    // (reference src/mlsl_impl.cpp:401-406)
    ownedKernelCount = (localKernelCount + D - 1) / D;
    localKernelCount = ownedKernelCount * D;
    ownedKernelOffset = ownedKernelCount * (size_t)dist->dataGroup->idx;
  } else {
    ownedKernelCount = localKernelCount;
    ownedKernelOffset = 0;
  }
  commBuf.ctx = ctx;
  if (needComm) {
    gradReq = new CommRequest(ctx, to_dtype(dt), o->uid, CommDesc::PARAM_GRAD);
    gradReq->desc.kind = distributedUpdate ? OpKind::REDUCE_SCATTER : OpKind::ALLREDUCE;
    gradReq->desc.count = ownedKernelCount * kernelSize;
    gradReq->desc.group = dist->dataGroup;
    gradReq->desc.compress = (ct == CT_QUANTIZATION) && !distributedUpdate;   // reference: AllReduce only
    gradReq->setup();
    if (distributedUpdate) {
      commBuf.bytes = gradReq->recv_bytes();   // the reduced owned shard lands here
      incReq = new CommRequest(ctx, to_dtype(dt), o->uid, CommDesc::PARAM_INC);
      incReq->desc.kind = OpKind::ALLGATHER;
      incReq->desc.count = ownedKernelCount * kernelSize;
      incReq->desc.group = dist->dataGroup;
      incReq->setup();
    }
  }
}

ParameterSetImpl::~ParameterSetImpl() {
  delete gradReq;
  delete incReq;
  delete fusedReq;
  commBuf.release();
}

void ParameterSetImpl::start_gradient(void* buf) {
  StatisticsImpl* st = op->session->stats;
  st->enter(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::START);
  if (needComm) {
    if (distributedUpdate) commBuf.allocate();
    gradReq->start(buf, distributedUpdate ? commBuf.ptr : buf);
  }
  st->leave(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::START, needComm ? gradReq : nullptr);
}

void* ParameterSetImpl::wait_gradient() {
  StatisticsImpl* st = op->session->stats;
  st->enter(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::WAIT);
  void* p = needComm ? gradReq->wait() : nullptr;
  st->leave(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::WAIT, needComm ? gradReq : nullptr);
  return p;
}

void* ParameterSetImpl::test_gradient(bool* done) {
  StatisticsImpl* st = op->session->stats;
  st->enter(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::TEST);
  void* p = nullptr;
  if (needComm) p = gradReq->test(done);
  else *done = true;
  st->leave(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::TEST, needComm ? gradReq : nullptr);
  return p;
}

void ParameterSetImpl::start_increment(void* buf) {
  StatisticsImpl* st = op->session->stats;
  st->enter(op->opIndex, StatisticsImpl::PARAM_INC, index, StatisticsImpl::START);
  // in-place all-gather: rank r's owned shard already sits at buf + r*owned*kernelSize
  if (needComm && distributedUpdate) incReq->start(buf, buf);
  st->leave(op->opIndex, StatisticsImpl::PARAM_INC, index, StatisticsImpl::START, needComm && distributedUpdate ? incReq : nullptr);
}

void* ParameterSetImpl::wait_increment() {
  StatisticsImpl* st = op->session->stats;
  st->enter(op->opIndex, StatisticsImpl::PARAM_INC, index, StatisticsImpl::WAIT);
  void* p = (needComm && distributedUpdate) ? incReq->wait() : nullptr;
  st->leave(op->opIndex, StatisticsImpl::PARAM_INC, index, StatisticsImpl::WAIT, needComm && distributedUpdate ? incReq : nullptr);
  return p;
}

void ParameterSetImpl::start_fused(void* grad, void* param, DataType paramType, void* master, void* s1, void* s2,
                                   const FusedUpdateParams* opt) {
  MLSLB_ASSERT(opt != nullptr, "fused update: optimizer parameters are NULL");
  MLSLB_ASSERT(distributedUpdate, "fused update requires a parameter set registered with distributedUpdate=true");
  RankContext* ctx = op->session->ctx;
  if (!fusedReq) {
    fusedReq = new CommRequest(ctx, to_dtype(dataType), op->uid, CommDesc::PARAM_GRAD);
    fusedReq->desc.kind = OpKind::FUSED_UPDATE;
    fusedReq->desc.count = ownedKernelCount * kernelSize;
    fusedReq->desc.group = dist->dataGroup;
    fusedReq->desc.has_out_dtype = true;
    fusedReq->desc.out_dtype = to_dtype(paramType);
    fusedReq->setup();
  }
  CommDesc::FusedUpdate& f = fusedReq->desc.fused;
  f.optimizer = (int)opt->type;
  f.lr = opt->lr;
  f.momentum = opt->momentum;
  f.beta1 = opt->beta1;
  f.beta2 = opt->beta2;
  f.eps = opt->eps;
  f.weight_decay = opt->weight_decay;
  f.step = opt->step;
  f.param = param;
  f.master = master;
  f.state1 = s1;
  f.state2 = s2;
  fusedReq->desc.scale = opt->grad_scale;
  fusedReq->desc.out_dtype = to_dtype(paramType);
  StatisticsImpl* st = op->session->stats;
  st->enter(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::START);
  fusedReq->start(grad, param);
  st->leave(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::START, fusedReq);
}

void ParameterSetImpl::wait_fused() {
  if (!fusedReq) return;
  StatisticsImpl* st = op->session->stats;
  st->enter(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::WAIT);
  fusedReq->wait();
  st->leave(op->opIndex, StatisticsImpl::PARAM_GRAD, index, StatisticsImpl::WAIT, fusedReq);
}

size_t ParameterSetImpl::grad_msg_bytes() const { return gradReq ? gradReq->msg_bytes() : 0; }
size_t ParameterSetImpl::inc_msg_bytes() const { return incReq ? incReq->msg_bytes() : 0; }

std::string ParameterSetImpl::describe() const {
  char buf[512];
  int n = snprintf(buf, sizeof(buf), "global_kernels %zu local %zu owned %zu owned_off %zu kernel_size %zu dist_update %d "
                   "grad %s inc %s compress %d", globalKernelCount, localKernelCount, ownedKernelCount,
                   ownedKernelOffset, kernelSize, (int)distributedUpdate, gradReq ? opkind_name(gradReq->desc.kind) : "-",
                   incReq ? opkind_name(incReq->desc.kind) : "-", (int)compress);
  return std::string(buf, (size_t)n);
}

// ============================================================================================================
// Operation
// ============================================================================================================
OperationImpl::OperationImpl(SessionImpl* s, OperationRegInfoImpl* i, DistributionImpl* d, size_t idx)
    : session(s), info(i), opIndex(idx) {
  MLSLB_ASSERT(s && i, "session or reg_info is null");
  MLSLB_ASSERT(s->globalMb > 0, "global batch size should be set before operation creation");
  uid = s->ctx->next_op_uid.fetch_add(1);
  info->refs++;
  opType = info->opType;
  name = info->name;
  if (d) bind(d);
}

OperationImpl::~OperationImpl() {
  for (auto a : inputs) delete a;
  for (auto a : outputs) delete a;
  for (auto p : params) delete p;
  if (--info->refs == 0) delete info;
}

void OperationImpl::bind(DistributionImpl* d) {
  MLSLB_ASSERT(dist == nullptr, "distribution can be set only once");
  MLSLB_ASSERT(d != nullptr, "distribution is NULL");
  dist = d;
  const size_t D = (size_t)dist->dataGroup->size();
  MLSLB_ASSERT(session->globalMb % dist->dataParts == 0,
               "global minibatch size (%zu) should be divisible by data partitions (%zu)", session->globalMb,
               dist->dataParts);
  localMb = session->globalMb / D;
  mbOffset = localMb * (size_t)dist->dataGroup->idx;
  for (size_t k = 0; k < info->inputs.size(); ++k)
    inputs.push_back(new ActivationImpl(this, info->inputs[k].count, info->inputs[k].size, info->inputs[k].dtype, true, k));
  for (size_t k = 0; k < info->outputs.size(); ++k)
    outputs.push_back(new ActivationImpl(this, info->outputs[k].count, info->outputs[k].size, info->outputs[k].dtype, false, k));
  for (size_t k = 0; k < info->params.size(); ++k)
    params.push_back(new ParameterSetImpl(this, info->params[k].count, info->params[k].size, info->params[k].dtype,
                                          info->params[k].distUpdate, info->params[k].compress, k));
}

void OperationImpl::commit() {
  MLSLB_ASSERT(dist != nullptr, "operation '%s' has no distribution at Commit", name.c_str());
  for (auto a : inputs)
    if (!a->peerSet) a->set_peer(nullptr);
  for (auto a : outputs) {
    if (!a->peerSet) a->set_peer(nullptr);
    a->connect();
  }
}

// ============================================================================================================
// Session
// ============================================================================================================
SessionImpl::SessionImpl(RankContext* c, PhaseType pt) : ctx(c), phase(pt) { stats = new StatisticsImpl(this); }

SessionImpl::~SessionImpl() {
  for (auto o : ops) delete o;
  delete stats;
}

EnvironmentImpl* env_of(RankContext* ctx) {
  if (!ctx->api_env) {
    ctx->api_env = new EnvironmentImpl(ctx);
    ctx->api_env_free = [](void* p) { delete (EnvironmentImpl*)p; };
  }
  return (EnvironmentImpl*)ctx->api_env;
}

}  // namespace impl

// ==============================================================================================================
// Public facade: every method forwards to the implementation object behind `this`
// ==============================================================================================================
using namespace impl;
#define SELF(T) static_cast<T*>(this)

size_t CommBlockInfo::GetMbOffset() { return SELF(BlockImpl)->mbOffset; }
size_t CommBlockInfo::GetMbCount() { return SELF(BlockImpl)->mbCount; }
size_t CommBlockInfo::GetFmOffset() { return SELF(BlockImpl)->fmOffset; }
size_t CommBlockInfo::GetFmCount() { return SELF(BlockImpl)->fmCount; }
size_t CommBlockInfo::GetFmSize() { return SELF(BlockImpl)->fmSize; }
DataType CommBlockInfo::GetDataType() { return SELF(BlockImpl)->dataType; }
size_t CommBlockInfo::GetBufOffset() { return SELF(BlockImpl)->bufOffset; }

size_t Activation::GetGlobalFmCount() { return SELF(ActivationImpl)->globalFmCount; }
size_t Activation::GetGlobalFmOffset() { return SELF(ActivationImpl)->globalFmOffset; }
size_t Activation::GetLocalFmCount() { return SELF(ActivationImpl)->localFmCount; }
size_t Activation::GetPackBlockCount() { return SELF(ActivationImpl)->packBlocks.size(); }
size_t Activation::GetUnpackBlockCount() { return SELF(ActivationImpl)->unpackBlocks.size(); }
CommBlockInfo* Activation::GetPackBlock(size_t idx) {
  MLSLB_ASSERT(idx < SELF(ActivationImpl)->packBlocks.size(), "invalid pack block idx %zu", idx);
  return SELF(ActivationImpl)->packBlocks[idx];
}
CommBlockInfo* Activation::GetUnpackBlock(size_t idx) {
  MLSLB_ASSERT(idx < SELF(ActivationImpl)->unpackBlocks.size(), "invalid unpack block idx %zu", idx);
  return SELF(ActivationImpl)->unpackBlocks[idx];
}
DataType Activation::GetDataType() { return SELF(ActivationImpl)->dataType; }
size_t Activation::GetFmSize() { return SELF(ActivationImpl)->fmSize; }
void* Activation::GetCommBuf() { return SELF(ActivationImpl)->commBuf.ptr; }
size_t Activation::GetCommBufSize() { return SELF(ActivationImpl)->commBuf.bytes; }
void Activation::StartComm(void* buf) { SELF(ActivationImpl)->start(buf); }
void* Activation::WaitComm() { return SELF(ActivationImpl)->wait(); }
void Activation::StartCommFused(void* localBuf, void* localDst) { SELF(ActivationImpl)->start_fused(localBuf, localDst); }
void Activation::Pack(const void* localBuf, void* commBuf) { SELF(ActivationImpl)->pack(localBuf, commBuf, false); }
void Activation::Unpack(const void* commBuf, void* localBuf) { SELF(ActivationImpl)->pack(localBuf, const_cast<void*>(commBuf), true); }

size_t ParameterSet::GetGlobalKernelCount() { return SELF(ParameterSetImpl)->globalKernelCount; }
size_t ParameterSet::GetGlobalKernelOffset() { return SELF(ParameterSetImpl)->globalKernelOffset; }
size_t ParameterSet::GetLocalKernelCount() { return SELF(ParameterSetImpl)->localKernelCount; }
size_t ParameterSet::GetOwnedKernelCount() { return SELF(ParameterSetImpl)->ownedKernelCount; }
size_t ParameterSet::GetOwnedKernelOffset() { return SELF(ParameterSetImpl)->ownedKernelOffset; }
DataType ParameterSet::GetDataType() { return SELF(ParameterSetImpl)->dataType; }
size_t ParameterSet::GetKernelSize() { return SELF(ParameterSetImpl)->kernelSize; }
bool ParameterSet::IsDistributedUpdate() { return SELF(ParameterSetImpl)->distributedUpdate; }
void ParameterSet::StartGradientComm(void* buf) { SELF(ParameterSetImpl)->start_gradient(buf); }
void ParameterSet::StartIncrementComm(void* buf) { SELF(ParameterSetImpl)->start_increment(buf); }
void* ParameterSet::WaitGradientComm() { return SELF(ParameterSetImpl)->wait_gradient(); }
void* ParameterSet::TestGradientComm(bool* isCompleted) { return SELF(ParameterSetImpl)->test_gradient(isCompleted); }
void* ParameterSet::WaitIncrementComm() { return SELF(ParameterSetImpl)->wait_increment(); }
void ParameterSet::StartFusedUpdate(void* grad, void* param, DataType paramType, void* master, void* state1,
                                    void* state2, const FusedUpdateParams* opt) {
  SELF(ParameterSetImpl)->start_fused(grad, param, paramType, master, state1, state2, opt);
}
void ParameterSet::WaitFusedUpdate() { SELF(ParameterSetImpl)->wait_fused(); }
void ParameterSet::SetGradientScale(float scale) {
  auto p = SELF(ParameterSetImpl);
  if (p->gradReq) p->gradReq->desc.scale = scale;
}

// ---- Distribution ----------------------------------------------------------------------------------------------
size_t Distribution::GetProcessIdx(GroupType gt) { return (size_t)SELF(DistributionImpl)->group(gt)->idx; }
size_t Distribution::GetProcessCount(GroupType gt) { return (size_t)SELF(DistributionImpl)->group(gt)->size(); }

static void check_count(size_t c) { MLSLB_ASSERT(c <= impl::kMaxCount, "element count %zu is out of range", c); }
static void check_root(DistributionImpl* d, GroupType gt, size_t root) {
  MLSLB_ASSERT(root < (size_t)d->group(gt)->size(), "root index %zu is outside the group (size %d)", root,
               d->group(gt)->size());
}

CommReq* Distribution::Bcast(void* buffer, size_t count, DataType dt, size_t rootIdx, GroupType gt) {
  auto d = SELF(DistributionImpl);
  check_count(count);
  check_root(d, gt, rootIdx);
  CommRequest* r = d->make_request(mlslb::OpKind::BCAST, dt, gt);
  r->desc.count = count;
  r->desc.root = rootIdx;
  return d->submit(r, buffer, buffer);
}
CommReq* Distribution::Reduce(void* sendBuffer, void* recvBuffer, size_t count, DataType dt, ReductionType rt,
                              size_t rootIdx, GroupType gt) {
  auto d = SELF(DistributionImpl);
  check_count(count);
  check_root(d, gt, rootIdx);
  CommRequest* r = d->make_request(mlslb::OpKind::REDUCE, dt, gt);
  r->desc.count = count;
  r->desc.root = rootIdx;
  r->desc.rop = to_redop(rt);
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::AllReduce(void* sendBuffer, void* recvBuffer, size_t count, DataType dt, ReductionType rt,
                                 GroupType gt) {
  return AllReduceEx(sendBuffer, recvBuffer, count, dt, rt, gt, 1.0f, CT_NONE);
}
CommReq* Distribution::AllReduceEx(void* sendBuffer, void* recvBuffer, size_t count, DataType dt, ReductionType rt,
                                   GroupType gt, float scale, CompressionType compress) {
  auto d = SELF(DistributionImpl);
  check_count(count);
  if (compress == CT_QUANTIZATION) {
    // persistent per (buffers, count, ...): the error-feedback residual lives in the request's backend state
    DistributionImpl::CompressKey key{sendBuffer, recvBuffer, count, (int)dt, (int)rt, (int)gt};
    auto it = d->compressed.find(key);
    if (it != d->compressed.end() && !it->second->active()) {
      CommRequest* r = it->second;
      r->desc.scale = scale;
      d->ctx->register_request(r);
      r->start(sendBuffer, recvBuffer);
      return (CommReq*)r;
    }
    if (it == d->compressed.end()) {
      if (d->compressed.size() >= 64) {          // bounded: drop idle entries (their residuals start from zero again)
        for (auto e = d->compressed.begin(); e != d->compressed.end();)
          if (!e->second->active()) {
            delete e->second;
            e = d->compressed.erase(e);
          } else {
            ++e;
          }
      }
      CommRequest* r = d->make_request(mlslb::OpKind::ALLREDUCE, dt, gt);
      r->one_shot = false;                       // Environment::Wait / Test leave it alive
      r->desc.count = count;
      r->desc.rop = to_redop(rt);
      r->desc.scale = scale;
      r->desc.compress = true;
      d->compressed[key] = r;
      return d->submit(r, sendBuffer, recvBuffer);
    }
    // same buffers started again while the previous run is still in flight: fall through to a one-shot request
  }
  CommRequest* r = d->make_request(mlslb::OpKind::ALLREDUCE, dt, gt);
  r->desc.count = count;
  r->desc.rop = to_redop(rt);
  r->desc.scale = scale;
  r->desc.compress = compress == CT_QUANTIZATION;
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::AlltoAll(void* sendBuffer, size_t sendCount, void* recvBuffer, DataType dt, GroupType gt) {
  auto d = SELF(DistributionImpl);
  check_count(sendCount);
  CommRequest* r = d->make_request(mlslb::OpKind::ALLTOALL, dt, gt);
  r->desc.count = sendCount;
  return d->submit(r, sendBuffer, recvBuffer);
}
static void fill_v(CommRequest* r, size_t P, size_t* sc, size_t* so, size_t* rc, size_t* ro) {
  MLSLB_ASSERT(sc && so && rc && ro, "count/offset arrays must not be NULL");
  r->desc.send_counts.assign(sc, sc + P);
  r->desc.send_offsets.assign(so, so + P);
  r->desc.recv_counts.assign(rc, rc + P);
  r->desc.recv_offsets.assign(ro, ro + P);
}
CommReq* Distribution::AlltoAllv(void* sendBuffer, size_t* sendCounts, size_t* sendOffsets, void* recvBuffer,
                                 size_t* recvCounts, size_t* recvOffsets, DataType dt, GroupType gt) {
  auto d = SELF(DistributionImpl);
  CommRequest* r = d->make_request(mlslb::OpKind::ALLTOALLV, dt, gt);
  fill_v(r, (size_t)d->group(gt)->size(), sendCounts, sendOffsets, recvCounts, recvOffsets);
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::SendRecvList(void* sendBuffer, size_t* sendCounts, size_t* sendOffsets, void* recvBuffer,
                                    size_t* recvCounts, size_t* recvOffsets, DataType dt, GroupType gt) {
  auto d = SELF(DistributionImpl);
  CommRequest* r = d->make_request(mlslb::OpKind::SENDRECV_LIST, dt, gt);
  fill_v(r, (size_t)d->group(gt)->size(), sendCounts, sendOffsets, recvCounts, recvOffsets);
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::Gather(void* sendBuffer, size_t sendCount, void* recvBuffer, DataType dt, size_t rootIdx,
                              GroupType gt) {
  auto d = SELF(DistributionImpl);
  check_count(sendCount);
  check_root(d, gt, rootIdx);
  CommRequest* r = d->make_request(mlslb::OpKind::GATHER, dt, gt);
  r->desc.count = sendCount;
  r->desc.root = rootIdx;
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::AllGather(void* sendBuffer, size_t sendCount, void* recvBuffer, DataType dt, GroupType gt) {
  auto d = SELF(DistributionImpl);
  check_count(sendCount);
  CommRequest* r = d->make_request(mlslb::OpKind::ALLGATHER, dt, gt);
  r->desc.count = sendCount;
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::AllGatherv(void* sendBuffer, size_t sendCount, void* recvBuffer, size_t* recvCounts,
                                  DataType dt, GroupType gt) {
  auto d = SELF(DistributionImpl);
  check_count(sendCount);
  MLSLB_ASSERT(recvCounts, "recvCounts must not be NULL");
  CommRequest* r = d->make_request(mlslb::OpKind::ALLGATHERV, dt, gt);
  r->desc.count = sendCount;
  size_t P = (size_t)d->group(gt)->size();
  r->desc.recv_counts.assign(recvCounts, recvCounts + P);
  MLSLB_ASSERT(r->desc.recv_counts[(size_t)d->group(gt)->idx] == sendCount,
               "AllGatherv: sendCount (%zu) must equal recvCounts[own index] (%zu)", sendCount,
               r->desc.recv_counts[(size_t)d->group(gt)->idx]);
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::Scatter(void* sendBuffer, void* recvBuffer, size_t recvCount, DataType dt, size_t rootIdx,
                               GroupType gt) {
  auto d = SELF(DistributionImpl);
  check_count(recvCount);
  check_root(d, gt, rootIdx);
  CommRequest* r = d->make_request(mlslb::OpKind::SCATTER, dt, gt);
  r->desc.count = recvCount;
  r->desc.root = rootIdx;
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::ReduceScatter(void* sendBuffer, void* recvBuffer, size_t recvCount, DataType dt,
                                     ReductionType rt, GroupType gt) {
  return ReduceScatterEx(sendBuffer, recvBuffer, recvCount, dt, rt, gt, 1.0f);
}
CommReq* Distribution::ReduceScatterEx(void* sendBuffer, void* recvBuffer, size_t recvCount, DataType dt,
                                       ReductionType rt, GroupType gt, float scale) {
  auto d = SELF(DistributionImpl);
  check_count(recvCount);
  CommRequest* r = d->make_request(mlslb::OpKind::REDUCE_SCATTER, dt, gt);
  r->desc.count = recvCount;
  r->desc.rop = to_redop(rt);
  r->desc.scale = scale;
  return d->submit(r, sendBuffer, recvBuffer);
}
CommReq* Distribution::GemmReduceScatter(const void* a, const void* w, void* out, size_t M, size_t N, size_t K,
                                         DataType outType, GroupType gt) {
  auto d = SELF(DistributionImpl);
  MLSLB_ASSERT(d->ctx->backend->is_device(), "GemmReduceScatter needs the CUDA backend");
  MLSLB_ASSERT(outType == DT_BF16 || outType == DT_FLOAT, "GemmReduceScatter: output must be bf16 or fp32");
  CommRequest* r = d->make_request(mlslb::OpKind::GEMM_RS, DT_BF16, gt);
  r->desc.gemm.M = (int)M;
  r->desc.gemm.N = (int)N;
  r->desc.gemm.K = (int)K;
  r->desc.gemm.a = a;
  r->desc.gemm.w = w;
  r->desc.has_out_dtype = true;
  r->desc.out_dtype = to_dtype(outType);
  return d->submit(r, const_cast<void*>(a), out);
}
CommReq* Distribution::AllGatherGemm(const void* xShard, const void* w, void* gathered, void* out, size_t M, size_t N, size_t K,
                                     DataType outType, GroupType gt) {
  auto d = SELF(DistributionImpl);
  MLSLB_ASSERT(d->ctx->backend->is_device(), "AllGatherGemm needs the CUDA backend");
  MLSLB_ASSERT(outType == DT_BF16 || outType == DT_FLOAT, "AllGatherGemm: output must be bf16 or fp32");
  MLSLB_ASSERT(xShard && w && gathered && out, "AllGatherGemm: NULL buffer");
  CommRequest* r = d->make_request(mlslb::OpKind::AG_GEMM, DT_BF16, gt);
  r->desc.gemm.M = (int)M;
  r->desc.gemm.N = (int)N;
  r->desc.gemm.K = (int)K;
  r->desc.gemm.a = xShard;
  r->desc.gemm.w = w;
  r->desc.gathered = gathered;
  r->desc.has_out_dtype = true;
  r->desc.out_dtype = to_dtype(outType);
  return d->submit(r, const_cast<void*>(xShard), out);
}
// ---- [ext] RMA windows -----------------------------------------------------------------------------------------
Window* Distribution::CreateWindow(void* base, size_t bytes, GroupType gt) {
  auto d = SELF(DistributionImpl);
  ProcessGroup* g = d->group(gt);
  MLSLB_ASSERT(base != nullptr && bytes > 0, "CreateWindow: empty window");
  MLSLB_ASSERT(d->ctx->backend->owns(base, bytes), "CreateWindow: the memory must come from Environment::Alloc");
  auto* w = new WindowImpl();
  w->dist = d;
  w->groupType = gt;
  w->group = g;
  const size_t P = (size_t)g->size();
  struct Msg {
    uint64_t off, bytes;
  } mine{d->ctx->backend->heap_offset(base), (uint64_t)bytes};
  std::vector<Msg> all(P);
  if (P > 1) d->ctx->boot->group_allgather(g->members, g->row, ++g->ctl_seq, &mine, all.data(), sizeof(Msg));
  else all[0] = mine;
  for (size_t i = 0; i < P; ++i) {
    w->offsets.push_back(all[i].off);
    w->sizes.push_back(all[i].bytes);
  }
  return w;
}
void Distribution::FreeWindow(Window* window) {
  if (!window) return;
  auto* w = static_cast<WindowImpl*>(window);
  w->Fence();   // nobody unmaps or reuses the memory while a peer may still access it
  delete w;
}
static void window_range_check(WindowImpl* w, size_t bytes, size_t idx, size_t disp) {
  MLSLB_ASSERT(idx < w->sizes.size(), "window target index %zu out of range (group size %zu)", idx, w->sizes.size());
  MLSLB_ASSERT(disp + bytes <= w->sizes[idx], "window access [%zu, %zu) exceeds the %zu bytes member %zu exposed", disp,
               disp + bytes, (size_t)w->sizes[idx], idx);
}
void Window::Put(const void* origin, size_t bytes, size_t targetIdx, size_t targetDisp) {
  auto* w = SELF(WindowImpl);
  window_range_check(w, bytes, targetIdx, targetDisp);
  mlslb::Backend* b = w->dist->ctx->backend.get();
  b->rma_copy((char*)b->peer_heap_ptr(w->group->members[targetIdx], w->offsets[targetIdx]) + targetDisp, origin, bytes);
}
void Window::Get(void* origin, size_t bytes, size_t targetIdx, size_t targetDisp) {
  auto* w = SELF(WindowImpl);
  window_range_check(w, bytes, targetIdx, targetDisp);
  mlslb::Backend* b = w->dist->ctx->backend.get();
  b->rma_copy(origin, (const char*)b->peer_heap_ptr(w->group->members[targetIdx], w->offsets[targetIdx]) + targetDisp, bytes);
}
void Window::Fence() {
  auto* w = SELF(WindowImpl);
  w->dist->Barrier(w->groupType);   // ordered behind the copies; its handshake makes them visible at the targets
}
size_t Window::GetSize(size_t memberIdx) {
  auto* w = SELF(WindowImpl);
  MLSLB_ASSERT(memberIdx < w->sizes.size(), "window member index out of range");
  return (size_t)w->sizes[memberIdx];
}

void Distribution::Barrier(GroupType gt) {
  auto d = SELF(DistributionImpl);
  CommRequest* r = d->make_request(mlslb::OpKind::BARRIER, DT_BYTE, gt);
  d->submit(r, nullptr, nullptr);
  r->wait();
  d->ctx->remove_request(r);
}

// ---- OperationRegInfo ------------------------------------------------------------------------------------------
void OperationRegInfo::SetName(const char* name) { SELF(OperationRegInfoImpl)->name = name ? name : ""; }
static void check_shape(size_t count, size_t size) {
  MLSLB_ASSERT(count > 0 && size > 0 && count <= impl::kMaxCount && size <= impl::kMaxCount,
               "count and size should be positive (got %zu, %zu)", count, size);
}
size_t OperationRegInfo::AddInput(size_t featureMapCount, size_t featureMapSize, DataType dt) {
  check_shape(featureMapCount, featureMapSize);
  auto& v = SELF(OperationRegInfoImpl)->inputs;
  v.push_back({featureMapCount, featureMapSize, dt, false, CT_NONE});
  return v.size() - 1;
}
size_t OperationRegInfo::AddOutput(size_t featureMapCount, size_t featureMapSize, DataType dt) {
  check_shape(featureMapCount, featureMapSize);
  auto& v = SELF(OperationRegInfoImpl)->outputs;
  v.push_back({featureMapCount, featureMapSize, dt, false, CT_NONE});
  return v.size() - 1;
}
size_t OperationRegInfo::AddParameterSet(size_t kernelCount, size_t kernelSize, DataType dt, bool distributedUpdate,
                                         CompressionType compressType) {
  check_shape(kernelCount, kernelSize);
  auto& v = SELF(OperationRegInfoImpl)->params;
  v.push_back({kernelCount, kernelSize, dt, distributedUpdate, compressType});
  return v.size() - 1;
}
void OperationRegInfo::Validate(Distribution*) {
  OpType t = SELF(OperationRegInfoImpl)->opType;
  bool ok = t == OT_CC || t == OT_BIAS || t == OT_ACT || t == OT_POOL || t == OT_DATA || t == OT_EVAL ||
            t == OT_BCAST || t == OT_CONCAT;
  MLSLB_ASSERT(ok, "operation type %d is not supported yet", (int)t);
}

// ---- Operation ---------------------------------------------------------------------------------------------------
void Operation::SetDistribution(Distribution* dist) { SELF(OperationImpl)->bind(static_cast<DistributionImpl*>(dist)); }
Distribution* Operation::GetDistribution() { return SELF(OperationImpl)->dist; }
Session* Operation::GetSession() { return SELF(OperationImpl)->session; }
OpType Operation::GetOpType() { return SELF(OperationImpl)->opType; }
void Operation::SetPrev(Operation* prev, size_t actIdx, size_t prevOpActIdx) {
  auto me = SELF(OperationImpl);
  MLSLB_ASSERT(actIdx < me->inputs.size(), "invalid input activation idx");
  if (!prev) {
    me->inputs[actIdx]->set_peer(nullptr);
    return;
  }
  auto p = static_cast<OperationImpl*>(prev);
  MLSLB_ASSERT(me->session == p->session, "different sessions");
  MLSLB_ASSERT(prevOpActIdx < p->outputs.size(), "invalid output activation idx");
  p->outputs[prevOpActIdx]->set_peer(me->inputs[actIdx]);
}
void Operation::SetNext(Operation* next, size_t actIdx, size_t nextOpActIdx) {
  auto me = SELF(OperationImpl);
  MLSLB_ASSERT(actIdx < me->outputs.size(), "invalid output activation idx");
  if (!next) {
    me->outputs[actIdx]->set_peer(nullptr);
    return;
  }
  auto n = static_cast<OperationImpl*>(next);
  MLSLB_ASSERT(me->session == n->session, "different sessions");
  MLSLB_ASSERT(nextOpActIdx < n->inputs.size(), "invalid input activation idx");
  me->outputs[actIdx]->set_peer(n->inputs[nextOpActIdx]);
}
const char* Operation::GetName() { return SELF(OperationImpl)->name.c_str(); }
size_t Operation::GetGlobalMinibatchSize() { return SELF(OperationImpl)->session->globalMb; }
size_t Operation::GetLocalMinibatchSize() { return SELF(OperationImpl)->localMb; }
size_t Operation::GetGlobalMinibatchOffset() { return SELF(OperationImpl)->mbOffset; }
size_t Operation::GetInputCount() { return SELF(OperationImpl)->inputs.size(); }
Activation* Operation::GetInput(size_t idx) {
  MLSLB_ASSERT(idx < SELF(OperationImpl)->inputs.size(), "invalid input activation idx %zu", idx);
  return SELF(OperationImpl)->inputs[idx];
}
size_t Operation::GetOutputCount() { return SELF(OperationImpl)->outputs.size(); }
Activation* Operation::GetOutput(size_t idx) {
  MLSLB_ASSERT(idx < SELF(OperationImpl)->outputs.size(), "invalid output activation idx %zu", idx);
  return SELF(OperationImpl)->outputs[idx];
}
bool Operation::HasParameterSets() { return !SELF(OperationImpl)->params.empty(); }
size_t Operation::GetParameterSetCount() { return SELF(OperationImpl)->params.size(); }
ParameterSet* Operation::GetParameterSet(size_t idx) {
  MLSLB_ASSERT(idx < SELF(OperationImpl)->params.size(), "invalid parameter set idx %zu", idx);
  return SELF(OperationImpl)->params[idx];
}

// ---- Session ---------------------------------------------------------------------------------------------------
void Session::SetGlobalMinibatchSize(size_t globalMinibatchSize) {
  auto s = SELF(SessionImpl);
  MLSLB_ASSERT(s->globalMb == 0, "global minibatch size can be set only once");
  MLSLB_ASSERT(globalMinibatchSize > 0, "global minibatch size must be positive");
  s->globalMb = globalMinibatchSize;
}
size_t Session::GetGlobalMinibatchSize() { return SELF(SessionImpl)->globalMb; }
PhaseType Session::GetPhaseType() { return SELF(SessionImpl)->phase; }
OperationRegInfo* Session::CreateOperationRegInfo(OpType opType) { return new OperationRegInfoImpl(opType); }
void Session::DeleteOperationRegInfo(OperationRegInfo* info) {
  auto i = static_cast<OperationRegInfoImpl*>(info);
  if (i && --i->refs == 0) delete i;   // operations built from it keep it alive
}
size_t Session::AddOperation(OperationRegInfo* info, Distribution* dist) {
  auto s = SELF(SessionImpl);
  MLSLB_ASSERT(!s->committed, "operations can not be added after Commit");
  info->Validate(dist);
  size_t idx = s->ops.size();
  s->ops.push_back(new OperationImpl(s, static_cast<OperationRegInfoImpl*>(info), static_cast<DistributionImpl*>(dist), idx));
  return idx;
}
void Session::RemoveOperations() {
  auto s = SELF(SessionImpl);
  for (auto o : s->ops) delete o;
  s->ops.clear();
  s->committed = false;
}
size_t Session::GetOperationCount() { return SELF(SessionImpl)->ops.size(); }
Operation* Session::GetOperation(size_t idx) {
  MLSLB_ASSERT(idx < SELF(SessionImpl)->ops.size(), "invalid operation idx %zu", idx);
  return SELF(SessionImpl)->ops[idx];
}
void Session::Commit() {
  auto s = SELF(SessionImpl);
  MLSLB_ASSERT(!s->committed, "commit should be called only once");
  s->ctx->session_ops_hint = (int)s->ops.size();
  s->stats->initialize();
  for (auto o : s->ops) {
    o->commit();
    for (auto a : o->inputs) a->commBuf.allocate();
    for (auto a : o->outputs) a->commBuf.allocate();
    for (auto p : o->params) p->commBuf.allocate();
    if (mlslb::log_level() >= mlslb::LOG_INFO) {
      MLSLB_LOG(mlslb::LOG_INFO, "operation:%s(%lld) in_acts:%zu out_acts:%zu param_sets:%zu local_mb_size:%zu global_mb_off:%zu",
                o->name.c_str(), (long long)o->uid, o->inputs.size(), o->outputs.size(), o->params.size(), o->localMb,
                o->mbOffset);
      for (auto a : o->inputs) MLSLB_LOG(mlslb::LOG_INFO, "  INPUT_ACT %zu: %s", a->index, a->describe().c_str());
      for (auto a : o->outputs) MLSLB_LOG(mlslb::LOG_INFO, "  OUTPUT_ACT %zu: %s", a->index, a->describe().c_str());
      for (auto p : o->params) MLSLB_LOG(mlslb::LOG_INFO, "  PARAM_SET %zu: %s", p->index, p->describe().c_str());
    }
  }
  s->committed = true;
  s->stats->collect_isolation();
}
Statistics* Session::GetStats() { return SELF(SessionImpl)->stats; }

// ---- Environment -------------------------------------------------------------------------------------------------
Environment& Environment::GetEnv() { return *env_of(mlslb::current_context()); }
int Environment::GetVersion() { return MLSL_VERSION(MLSL_MAJOR_VERSION, MLSL_MINOR_VERSION); }

void Environment::Configure(const char* config) {
  // Only "color=N" exists: split the global group so that ranks with the same colour form independent jobs
  // (reference src/mlsl.cpp:620-647).
  if (!config) return;
  auto e = SELF(EnvironmentImpl);
  MLSLB_ASSERT(e->ctx->initialized, "Configure must be called after Init");
  const char* p = strstr(config, "color=");
  if (!p) return;
  int color = atoi(p + 6);
  mlslb::RankContext* c = e->ctx;
  mlslb::ProcessGroup* old = c->global_group;
  c->global_group = c->create_group_by_color(old, color);
  if (old != c->world_group) c->free_group(old);
}

void Environment::Init(int*, char***) {
  auto e = SELF(EnvironmentImpl);
  MLSLB_ASSERT(!e->ctx->initialized, "MLSL can be initialized only once");
  mlslb::context_init(e->ctx);
  mlslb::RankContext* c = e->ctx;
  if (c->env.auto_config != 0) mlslb::auto_config(c);
  if (c->env.wait_mode == "stream") c->backend->set_wait_mode(true);
  if (c->rank == 0) {
    mlslb::print_env(c->env);
    MLSLB_LOG(mlslb::LOG_INFO, "%s", MLSLB_PACKAGE_VERSION);
    MLSLB_LOG(mlslb::LOG_INFO, "MLSL API: %d.%d, backend: %s, servers: %d", MLSL_MAJOR_VERSION, MLSL_MINOR_VERSION,
              c->backend->describe().c_str(), c->progress->servers());
  }
}

void Environment::Finalize() {
  auto e = SELF(EnvironmentImpl);
  mlslb::RankContext* c = e->ctx;
  if (!c->initialized) {
    MLSLB_LOG(mlslb::LOG_INFO, "MLSL isn't initialized, skip finalization");
    return;
  }
  if (c->init_pid != (int)getpid()) {
    MLSLB_LOG(mlslb::LOG_INFO, "different pids: init_pid %d, current_pid %d, skip finalization", c->init_pid, (int)getpid());
    return;
  }
  size_t leaked;
  {
    std::lock_guard<std::mutex> g(c->req_mu);
    leaked = c->inflight.size();
  }
  if (leaked) MLSLB_LOG(mlslb::LOG_INFO, "there are %zu incompleted requests", leaked);
  mlslb::context_finalize(c);
  if (e->quantView) {
    free(e->quantView->lib_path);
    free(e->quantView->quant_buffer_func_name);
    free(e->quantView->dequant_buffer_func_name);
    free(e->quantView->reduce_sum_func_name);
    delete e->quantView;
    e->quantView = nullptr;
  }
}

bool Environment::IsInitialized() { return SELF(EnvironmentImpl)->ctx->initialized; }
static mlslb::RankContext* live(Environment* env) {
  auto c = static_cast<EnvironmentImpl*>(env)->ctx;
  MLSLB_ASSERT(c->initialized, "MLSL is not initialized");
  return c;
}
size_t Environment::GetProcessIdx() { return (size_t)live(this)->global_group->idx; }
size_t Environment::GetProcessCount() { return (size_t)live(this)->global_group->size(); }
Session* Environment::CreateSession(PhaseType phaseType) { return new SessionImpl(live(this), phaseType); }
void Environment::DeleteSession(Session* session) { delete static_cast<SessionImpl*>(session); }
Distribution* Environment::CreateDistribution(size_t dataPartitions, size_t modelPartitions) {
  return new DistributionImpl(live(this), dataPartitions, modelPartitions, true, -1, -1);
}
Distribution* Environment::CreateDistributionWithColors(int dataColor, int modelColor) {
  MLSLB_ASSERT(dataColor >= 0 && modelColor >= 0, "colors must be non-negative");
  return new DistributionImpl(live(this), 0, 0, false, dataColor, modelColor);
}
void Environment::GetGroupState(unsigned long long* rowsInUse, unsigned long long* ticketMark) {
  mlslb::RankContext* c = live(this);
  MLSLB_ASSERT(rowsInUse != nullptr && ticketMark != nullptr, "output pointers are NULL");
  *rowsInUse = c->row_used;
  *ticketMark = std::max(c->seq_hwm, c->global_group->hwm());
}
Distribution* Environment::CreateDistributionFromRanks(const size_t* ranks, size_t count, unsigned long long rowsInUse,
                                                       unsigned long long ticketMark) {
  mlslb::RankContext* c = live(this);
  MLSLB_ASSERT(ranks != nullptr && count > 0, "the rank list is empty");
  std::vector<int> members(count);
  for (size_t i = 0; i < count; ++i) {
    MLSLB_ASSERT(ranks[i] < (size_t)c->global_group->size(), "rank %zu is outside the global group (%d processes)",
                 ranks[i], c->global_group->size());
    members[i] = c->global_group->members[ranks[i]];
  }
  return new DistributionImpl(c, c->create_group_from_members(members, rowsInUse, ticketMark));
}
void Environment::DeleteDistribution(Distribution* distribution) { delete static_cast<DistributionImpl*>(distribution); }
void Environment::Wait(CommReq* req) {
  MLSLB_ASSERT(req != nullptr, "request is NULL");
  auto r = reinterpret_cast<mlslb::CommRequest*>(req);
  r->wait();
  live(this)->remove_request(r);
}
void Environment::Test(CommReq* req, bool* isCompleted) {
  MLSLB_ASSERT(req != nullptr && isCompleted != nullptr, "request or completion flag is NULL");
  auto r = reinterpret_cast<mlslb::CommRequest*>(req);
  r->test(isCompleted);
  if (*isCompleted) live(this)->remove_request(r);
}
void* Environment::Alloc(size_t size, size_t alignment) { return live(this)->backend->alloc(size, alignment); }
void Environment::Free(void* ptr) { live(this)->backend->free(ptr); }

void Environment::SetQuantizationParams(QuantParams* params) {
  auto e = SELF(EnvironmentImpl);
  mlslb::RankContext* c = live(this);
  MLSLB_ASSERT(params != nullptr, "quantization parameters are NULL");
  MLSLB_ASSERT(!c->quant.set, "quantization parameters can be set only once");
  auto dupz = [](const char* s) { return s ? std::string(s) : std::string(); };
  c->quant.set = true;
  c->quant.lib_path = dupz(params->lib_path);
  c->quant.quant_name = dupz(params->quant_buffer_func_name);
  c->quant.dequant_name = dupz(params->dequant_buffer_func_name);
  c->quant.reduce_name = dupz(params->reduce_sum_func_name);
  c->quant.block_size = params->block_size ? params->block_size : mlslb::kQuantBlockBytes;
  c->quant.elem_in_block = params->elem_in_block ? params->elem_in_block : (size_t)mlslb::kQuantBlock;
  e->quantView = new QuantParams();
  e->quantView->lib_path = strdup(c->quant.lib_path.c_str());
  e->quantView->quant_buffer_func_name = strdup(c->quant.quant_name.c_str());
  e->quantView->dequant_buffer_func_name = strdup(c->quant.dequant_name.c_str());
  e->quantView->reduce_sum_func_name = strdup(c->quant.reduce_name.c_str());
  e->quantView->block_size = c->quant.block_size;
  e->quantView->elem_in_block = c->quant.elem_in_block;
  if (!c->quant.lib_path.empty() && c->backend->is_device())
    MLSLB_LOG(mlslb::LOG_INFO,
              "quantization library %s: host functions cannot run inside the device kernels - the CUDA backend uses its "
              "built-in fused fp8 block format (the host backend calls the library)", c->quant.lib_path.c_str());
}
QuantParams* Environment::GetQuantizationParams() { return SELF(EnvironmentImpl)->quantView; }

void Environment::SetStream(void* s) { live(this)->backend->set_user_stream(s); }
void* Environment::GetStream() { return live(this)->backend->user_stream(); }
void Environment::SetWaitMode(const char* mode) {
  MLSLB_ASSERT(mode && (!strcmp(mode, "host") || !strcmp(mode, "stream")), "wait mode must be 'host' or 'stream'");
  live(this)->backend->set_wait_mode(!strcmp(mode, "stream"));
}
size_t Environment::GetLaunchOrder(long long* uids, size_t capacity) {
  std::vector<int64_t> v = live(this)->progress->recent_launches();
  const size_t n = std::min(capacity, v.size());
  for (size_t i = 0; i < n; ++i) uids[i] = v[v.size() - n + i];
  return n;
}
void Environment::SetTuning(const char* key, long value) {
  MLSLB_ASSERT(key && tune_set(live(this)->env.tune, key, value), "unknown tuning key '%s'", key ? key : "(null)");
}
long Environment::GetTuning(const char* key) {
  long v = 0;
  MLSLB_ASSERT(key && tune_get(live(this)->env.tune, key, &v), "unknown tuning key '%s'", key ? key : "(null)");
  return v;
}
const char* Environment::GetBackendName() { return live(this)->backend->name(); }
const char* Environment::DescribeBackend() {
  auto e = SELF(EnvironmentImpl);
  e->backendDesc = live(this)->backend->describe();
  return e->backendDesc.c_str();
}
bool Environment::IsDeviceBackend() { return live(this)->backend->is_device(); }
void Environment::SuspendServers() { live(this)->progress->suspend(); }
void Environment::ResumeServers() { live(this)->progress->resume(); }

}  // namespace MLSL
