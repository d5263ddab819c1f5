// Communication statistics (reference src/mlsl_impl_stats.cpp:40-668): every Start/Wait/Test of the graph API is
// bracketed; the time since the previous MLSL call is booked as compute of the calling operation, the call itself
// as communication of the entity.  "Isolation" numbers are measured at Commit by running every communication of
// the session stats_iters times (first stats_skip discarded).  Differences from the reference: counters exist in
// TSC cycles (API compatibility) and nanoseconds; only rank 0 touches mlsl_stats.log (the reference lets every rank
// truncate it); on the CUDA backend the timed calls include the device execution because Wait blocks on the
// completion event.
#include <cstdio>
#include <cstring>

#include "graph.hpp"
#include "log.hpp"

namespace MLSL {
namespace impl {

using namespace mlslb;

StatisticsImpl::StatisticsImpl(SessionImpl* s) : session(s) { enabled = s->ctx->env.stats; }

size_t StatisticsImpl::slot(size_t opIdx, Kind k, size_t entIdx) const {
  OperationImpl* o = session->ops[opIdx];
  switch (k) {
    case INPUT_ACT: return entIdx;
    case OUTPUT_ACT: return o->inputs.size() + entIdx;
    case PARAM_GRAD: return o->inputs.size() + o->outputs.size() + 2 * entIdx;
    case PARAM_INC: return o->inputs.size() + o->outputs.size() + 2 * entIdx + 1;
  }
  return 0;
}

void StatisticsImpl::initialize() {
  ops.clear();
  ops.resize(session->ops.size());
  for (size_t i = 0; i < session->ops.size(); ++i) {
    OperationImpl* o = session->ops[i];
    ops[i].ent.assign(o->inputs.size() + o->outputs.size() + 2 * o->params.size(), EntityStat());
  }
}

void StatisticsImpl::start() {
  if (!enabled) return;
  started = true;
  lastCycles = cycles_now();
  lastNs = now_ns();
}
void StatisticsImpl::stop() { started = false; }

void StatisticsImpl::reset() {
  for (auto& o : ops)
    for (auto& e : o.ent) {
      e.commCycles = e.computeCycles = e.commNs = e.computeNs = e.devCommNs = e.devRuns = 0;
      e.commBytes = 0;
    }
  batches = 0;
  lastCycles = cycles_now();
  lastNs = now_ns();
}

// Start/Wait of one transfer are issued by different activations; book both on the entity that STARTED it so a
// row of the report describes one message (reference src/mlsl_impl_stats.cpp:575-621).
static bool resolve(StatisticsImpl* st, size_t& opIdx, StatisticsImpl::Kind& k, size_t& entIdx, StatisticsImpl::Action a) {
  if (a == StatisticsImpl::START || k == StatisticsImpl::PARAM_GRAD || k == StatisticsImpl::PARAM_INC) return true;
  OperationImpl* o = st->session->ops[opIdx];
  ActivationImpl* act = (k == StatisticsImpl::INPUT_ACT) ? o->inputs[entIdx] : o->outputs[entIdx];
  ActivationImpl* peer = act->peer;
  if (!peer) return false;
  opIdx = peer->op->opIndex;
  entIdx = peer->index;
  k = peer->isInput ? StatisticsImpl::INPUT_ACT : StatisticsImpl::OUTPUT_ACT;
  return true;
}

void StatisticsImpl::enter(size_t opIdx, Kind k, size_t entIdx, Action a) {
  if (!enabled || !started || collecting) return;
  if (!resolve(this, opIdx, k, entIdx, a)) return;
  unsigned long long c = cycles_now(), n = now_ns();
  EntityStat& e = ops[opIdx].ent[slot(opIdx, k, entIdx)];
  e.computeCycles += c - lastCycles;
  e.computeNs += n - lastNs;
  if (k == PARAM_GRAD && a == WAIT && entIdx == 0) batches++;
  lastCycles = c;
  lastNs = n;
}

void StatisticsImpl::leave(size_t opIdx, Kind k, size_t entIdx, Action a, CommRequest* req) {
  if (!enabled || !started || collecting) return;
  if (!resolve(this, opIdx, k, entIdx, a)) return;
  unsigned long long c = cycles_now(), n = now_ns();
  EntityStat& e = ops[opIdx].ent[slot(opIdx, k, entIdx)];
  e.commCycles += c - lastCycles;
  e.commNs += n - lastNs;
  // Device timestamps: what the collective took ON THE DEVICE, from the event pair the backend records
  // around the kernel.  With stream-ordered waits the host calls above only bracket a launch, so the device duration is
  // what the comm counters have to carry; with host-blocking waits the bracket already contains it.
  if (req) {
    RankContext* ctx = session->ctx;
    ctx->backend->harvest_device_time(*req);
    if (unsigned long long d = req->take_device_ns()) {
      e.devCommNs += d;
      e.devRuns++;
      if (ctx->backend->stream_ordered_wait()) {
        e.commNs += d;
        e.commCycles += (unsigned long long)((double)d * cycles_per_ns());
      }
    }
  }
  if (a == START) e.commBytes += e.bytesPerIter;
  lastCycles = c;
  lastNs = n;
}

double StatisticsImpl::cycles_per_ns() {
  static double ratio = 0.0;
  if (ratio == 0.0) {
    const unsigned long long c0 = cycles_now(), n0 = now_ns();
    while (now_ns() - n0 < 2000000ull) {}
    ratio = (double)(cycles_now() - c0) / (double)(now_ns() - n0);
    if (!(ratio > 0.0)) ratio = 1.0;
  }
  return ratio;
}

void StatisticsImpl::collect_isolation() {
  // message sizes are always recorded (GetCommSize works even with MLSL_STATS=0 dry runs disabled)
  for (size_t i = 0; i < session->ops.size(); ++i) {
    OperationImpl* o = session->ops[i];
    for (auto a : o->inputs) ops[i].ent[slot(i, INPUT_ACT, a->index)].bytesPerIter = a->needComm ? a->msg_bytes() : 0;
    for (auto a : o->outputs) ops[i].ent[slot(i, OUTPUT_ACT, a->index)].bytesPerIter = a->needComm ? a->msg_bytes() : 0;
    for (auto p : o->params) {
      ops[i].ent[slot(i, PARAM_GRAD, p->index)].bytesPerIter = p->grad_msg_bytes();
      ops[i].ent[slot(i, PARAM_INC, p->index)].bytesPerIter = p->inc_msg_bytes();
    }
  }
  if (!enabled) return;
  RankContext* ctx = session->ctx;
  const int iters = ctx->env.stats_iters, skip = ctx->env.stats_skip;
  collecting = true;
  const bool was_stream = ctx->backend->stream_ordered_wait();
  if (was_stream) ctx->backend->set_wait_mode(false);      // dry runs block the host: the bracket holds the whole collective
  size_t maxParam = 64;
  for (auto o : session->ops)
    for (auto p : o->params)
      maxParam = std::max(maxParam, p->localKernelCount * p->kernelSize * dtype_size(to_dtype(p->dataType)));
  void* scratch = ctx->backend->alloc(maxParam, 4096);
  auto timed = [&](EntityStat& e, const std::function<void()>& fn) {
    unsigned long long acc = 0;
    for (int it = 0; it < iters; ++it) {
      if (it == skip) acc = 0;
      unsigned long long t0 = cycles_now();
      fn();
      acc += cycles_now() - t0;
    }
    e.isolationCycles = acc;
  };
  for (size_t i = 0; i < session->ops.size(); ++i) {
    OperationImpl* o = session->ops[i];
    for (auto a : o->outputs)
      if (a->needComm && a->peer)
        timed(ops[i].ent[slot(i, OUTPUT_ACT, a->index)], [&] {
          a->start(a->commBuf.ptr);
          a->peer->wait();
        });
    for (auto a : o->inputs)
      if (a->needComm && a->peer && a->req && a->req->desc.kind != OpKind::BARRIER)
        timed(ops[i].ent[slot(i, INPUT_ACT, a->index)], [&] {
          a->start(a->commBuf.ptr);
          a->peer->wait();
        });
    for (auto p : o->params) {
      if (!p->needComm) continue;
      timed(ops[i].ent[slot(i, PARAM_GRAD, p->index)], [&] {
        p->start_gradient(scratch);
        p->wait_gradient();
      });
      if (p->distributedUpdate)
        timed(ops[i].ent[slot(i, PARAM_INC, p->index)], [&] {
          p->start_increment(scratch);
          p->wait_increment();
        });
    }
  }
  ctx->backend->free(scratch);
  if (was_stream) ctx->backend->set_wait_mode(true);
  collecting = false;
  reset();
}

static const char* kind_name(size_t slotIdx, OperationImpl* o) {
  if (slotIdx < o->inputs.size()) return "IA";
  if (slotIdx < o->inputs.size() + o->outputs.size()) return "OA";
  return ((slotIdx - o->inputs.size() - o->outputs.size()) & 1) ? "INC" : "GRAD";
}

void StatisticsImpl::print() {
  if (!enabled) return;
  RankContext* ctx = session->ctx;
  if (ctx->rank != 0) return;
  FILE* f = fopen("mlsl_stats.log", "w");
  FILE* outs[2] = {stdout, f};
  const int iters = std::max(1, ctx->env.stats_iters - ctx->env.stats_skip);
  const unsigned long long nb = batches ? batches : 1;
  const size_t mb = session->globalMb ? session->globalMb : 1;
  for (FILE* o : outs) {
    if (!o) continue;
    fprintf(o, "MLSL statistics (batches %llu, global minibatch %zu)\n", batches, session->globalMb);
    fprintf(o, "%-24s %-5s %12s %16s %16s %16s %14s %14s\n", "operation", "ent", "KB/iter", "isol Kcyc/img", "comm Kcyc/img",
            "comp Kcyc/img", "comm us/iter", "device us/run");
    for (size_t i = 0; i < ops.size(); ++i) {
      OperationImpl* op = session->ops[i];
      for (size_t s = 0; s < ops[i].ent.size(); ++s) {
        const EntityStat& e = ops[i].ent[s];
        if (!e.bytesPerIter && !e.commCycles) continue;
        fprintf(o, "%-24s %-5s %12.1f %16.2f %16.2f %16.2f %14.1f %14.1f\n", op->name.c_str(), kind_name(s, op),
                e.bytesPerIter / 1024.0, e.isolationCycles / 1000.0 / iters / mb, e.commCycles / 1000.0 / nb / mb,
                e.computeCycles / 1000.0 / nb / mb, e.commNs / 1000.0 / nb, e.devRuns ? e.devCommNs / 1000.0 / e.devRuns : 0.0);
      }
    }
    fprintf(o, "TOTAL: comm size %zu bytes, isolation %llu cycles, comm %llu cycles, compute %llu cycles\n",
            GetTotalCommSize(), GetTotalIsolationCommCycles(), GetTotalCommCycles(), GetTotalComputeCycles());
    fflush(o);
  }
  if (f) fclose(f);
}

}  // namespace impl

using namespace impl;
#define SELF(T) static_cast<T*>(this)
static const OpStat& opstat(Statistics* s, size_t opIdx) {
  auto st = static_cast<StatisticsImpl*>(s);
  MLSLB_ASSERT(opIdx < st->ops.size(), "invalid operation idx %zu", opIdx);
  return st->ops[opIdx];
}
void Statistics::Start() { SELF(StatisticsImpl)->start(); }
void Statistics::Stop() { SELF(StatisticsImpl)->stop(); }
void Statistics::Reset() { SELF(StatisticsImpl)->reset(); }
bool Statistics::IsStarted() { return SELF(StatisticsImpl)->started; }
bool Statistics::IsEnabled() { return SELF(StatisticsImpl)->enabled; }
void Statistics::Print() { SELF(StatisticsImpl)->print(); }
unsigned long long Statistics::GetIsolationCommCycles(size_t opIdx) {
  unsigned long long t = 0;
  for (auto& e : opstat(this, opIdx).ent) t += e.isolationCycles;
  return t;
}
size_t Statistics::GetCommSize(size_t opIdx) {
  size_t t = 0;
  for (auto& e : opstat(this, opIdx).ent) t += e.commBytes;
  return t;
}
unsigned long long Statistics::GetCommCycles(size_t opIdx) {
  unsigned long long t = 0;
  for (auto& e : opstat(this, opIdx).ent) t += e.commCycles;
  return t;
}
unsigned long long Statistics::GetComputeCycles(size_t opIdx) {
  unsigned long long t = 0;
  for (auto& e : opstat(this, opIdx).ent) t += e.computeCycles;
  return t;
}
unsigned long long Statistics::GetCommNanos(size_t opIdx) {
  unsigned long long t = 0;
  for (auto& e : opstat(this, opIdx).ent) t += e.commNs;
  return t;
}
unsigned long long Statistics::GetDeviceCommNanos(size_t opIdx) {
  unsigned long long t = 0;
  for (auto& e : opstat(this, opIdx).ent) t += e.devCommNs;
  return t;
}
unsigned long long Statistics::GetComputeNanos(size_t opIdx) {
  unsigned long long t = 0;
  for (auto& e : opstat(this, opIdx).ent) t += e.computeNs;
  return t;
}
unsigned long long Statistics::GetTotalIsolationCommCycles() {
  unsigned long long t = 0;
  for (size_t i = 0; i < SELF(StatisticsImpl)->ops.size(); ++i) t += GetIsolationCommCycles(i);
  return t;
}
size_t Statistics::GetTotalCommSize() {
  size_t t = 0;
  for (size_t i = 0; i < SELF(StatisticsImpl)->ops.size(); ++i) t += GetCommSize(i);
  return t;
}
unsigned long long Statistics::GetTotalCommCycles() {
  unsigned long long t = 0;
  for (size_t i = 0; i < SELF(StatisticsImpl)->ops.size(); ++i) t += GetCommCycles(i);
  return t;
}
unsigned long long Statistics::GetTotalComputeCycles() {
  unsigned long long t = 0;
  for (size_t i = 0; i < SELF(StatisticsImpl)->ops.size(); ++i) t += GetComputeCycles(i);
  return t;
}

}  // namespace MLSL
