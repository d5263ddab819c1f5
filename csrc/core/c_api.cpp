// C binding: thin handle wrappers over the C++ facade (reference src/c_bind.cpp:31-651 does the same with a
// TRY_CATCH_RETURN macro).  Handles are the object addresses.  Failures are reported as CMLSL_FAILURE with the
// message available from mlsl_last_error() when assertions are in "throw" mode (mlsl_set_assert_throws(1) or
// MLSL_ASSERT_MODE=throw); in the default mode a failed assertion terminates the job like the reference.
#include <exception>
#include <string>

#include "../../include/mlsl.h"
#include "../../include/mlsl.hpp"
#include "log.hpp"
#include "runtime.hpp"

namespace {
thread_local std::string g_last_error;

template <typename T>
inline T* H(mlsl_handle_t h) {
  if (!h) throw mlslb::Error("null handle");
  return reinterpret_cast<T*>(static_cast<uintptr_t>(h));
}
template <typename T>
inline mlsl_handle_t U(T* p) {
  return static_cast<mlsl_handle_t>(reinterpret_cast<uintptr_t>(p));
}
template <typename T>
inline T* need(T* p) {
  if (!p) throw mlslb::Error("null output pointer");
  return p;
}
}  // namespace

#define C_GUARD(...)                                  \
  try {                                               \
    __VA_ARGS__;                                      \
    return CMLSL_SUCCESS;                             \
  } catch (const std::exception& e) {                 \
    g_last_error = e.what();                          \
    return CMLSL_FAILURE;                             \
  } catch (...) {                                     \
    g_last_error = "unknown exception";               \
    return CMLSL_FAILURE;                             \
  }

using MLSL::Activation;
using MLSL::CommBlockInfo;
using MLSL::CommReq;
using MLSL::Distribution;
using MLSL::Environment;
using MLSL::Operation;
using MLSL::OperationRegInfo;
using MLSL::ParameterSet;
using MLSL::Session;
using MLSL::Statistics;

static inline MLSL::DataType DT(mlsl_data_type d) { return (MLSL::DataType)(int)d; }
static inline MLSL::GroupType GT(mlsl_group_type g) { return (MLSL::GroupType)(int)g; }
static inline MLSL::ReductionType RT(mlsl_reduction_type r) { return (MLSL::ReductionType)(int)r; }

extern "C" {

const char* mlsl_last_error(void) { return g_last_error.c_str(); }

// ---- CommBlockInfo ----
int mlsl_comm_block_info_get_mb_offset(mlsl_comm_block_info b, size_t* v) { C_GUARD(*need(v) = H<CommBlockInfo>(b)->GetMbOffset()) }
int mlsl_comm_block_info_get_mb_count(mlsl_comm_block_info b, size_t* v) { C_GUARD(*need(v) = H<CommBlockInfo>(b)->GetMbCount()) }
int mlsl_comm_block_info_get_fm_offset(mlsl_comm_block_info b, size_t* v) { C_GUARD(*need(v) = H<CommBlockInfo>(b)->GetFmOffset()) }
int mlsl_comm_block_info_get_fm_count(mlsl_comm_block_info b, size_t* v) { C_GUARD(*need(v) = H<CommBlockInfo>(b)->GetFmCount()) }
int mlsl_comm_block_info_get_fm_size(mlsl_comm_block_info b, size_t* v) { C_GUARD(*need(v) = H<CommBlockInfo>(b)->GetFmSize()) }
int mlsl_comm_block_info_get_data_type(mlsl_comm_block_info b, mlsl_data_type* v) {
  C_GUARD(*need(v) = (mlsl_data_type)(int)H<CommBlockInfo>(b)->GetDataType())
}
int mlsl_comm_block_info_get_buf_offset(mlsl_comm_block_info b, size_t* v) { C_GUARD(*need(v) = H<CommBlockInfo>(b)->GetBufOffset()) }

// ---- Activation ----
int mlsl_activation_get_global_fm_count(mlsl_activation a, size_t* v) { C_GUARD(*need(v) = H<Activation>(a)->GetGlobalFmCount()) }
int mlsl_activation_get_global_fm_offset(mlsl_activation a, size_t* v) { C_GUARD(*need(v) = H<Activation>(a)->GetGlobalFmOffset()) }
int mlsl_activation_get_local_fm_count(mlsl_activation a, size_t* v) { C_GUARD(*need(v) = H<Activation>(a)->GetLocalFmCount()) }
int mlsl_activation_get_pack_block_count(mlsl_activation a, size_t* v) { C_GUARD(*need(v) = H<Activation>(a)->GetPackBlockCount()) }
int mlsl_activation_get_unpack_block_count(mlsl_activation a, size_t* v) { C_GUARD(*need(v) = H<Activation>(a)->GetUnpackBlockCount()) }
int mlsl_activation_get_pack_block(mlsl_activation a, size_t i, mlsl_comm_block_info* v) {
  C_GUARD(*need(v) = U(H<Activation>(a)->GetPackBlock(i)))
}
int mlsl_activation_get_unpack_block(mlsl_activation a, size_t i, mlsl_comm_block_info* v) {
  C_GUARD(*need(v) = U(H<Activation>(a)->GetUnpackBlock(i)))
}
int mlsl_activation_get_data_type(mlsl_activation a, mlsl_data_type* v) {
  C_GUARD(*need(v) = (mlsl_data_type)(int)H<Activation>(a)->GetDataType())
}
int mlsl_activation_get_fm_size(mlsl_activation a, size_t* v) { C_GUARD(*need(v) = H<Activation>(a)->GetFmSize()) }
int mlsl_activation_get_comm_buf(mlsl_activation a, void** v) { C_GUARD(*need(v) = H<Activation>(a)->GetCommBuf()) }
int mlsl_activation_get_comm_buf_size(mlsl_activation a, size_t* v) { C_GUARD(*need(v) = H<Activation>(a)->GetCommBufSize()) }
int mlsl_activation_start_comm(mlsl_activation a, void* buf) { C_GUARD(H<Activation>(a)->StartComm(buf)) }
int mlsl_activation_start_comm_fused(mlsl_activation a, void* local, void* dst) { C_GUARD(H<Activation>(a)->StartCommFused(local, dst)) }
int mlsl_activation_wait_comm(mlsl_activation a, void** v) { C_GUARD(*need(v) = H<Activation>(a)->WaitComm()) }
int mlsl_activation_pack(mlsl_activation a, const void* local_buf, void* comm_buf) { C_GUARD(H<Activation>(a)->Pack(local_buf, comm_buf)) }
int mlsl_activation_unpack(mlsl_activation a, const void* comm_buf, void* local_buf) { C_GUARD(H<Activation>(a)->Unpack(comm_buf, local_buf)) }

// ---- ParameterSet ----
int mlsl_parameter_set_get_global_kernel_count(mlsl_parameter_set p, size_t* v) { C_GUARD(*need(v) = H<ParameterSet>(p)->GetGlobalKernelCount()) }
int mlsl_parameter_set_get_global_kernel_offset(mlsl_parameter_set p, size_t* v) { C_GUARD(*need(v) = H<ParameterSet>(p)->GetGlobalKernelOffset()) }
int mlsl_parameter_set_get_local_kernel_count(mlsl_parameter_set p, size_t* v) { C_GUARD(*need(v) = H<ParameterSet>(p)->GetLocalKernelCount()) }
int mlsl_parameter_set_get_owned_kernel_count(mlsl_parameter_set p, size_t* v) { C_GUARD(*need(v) = H<ParameterSet>(p)->GetOwnedKernelCount()) }
int mlsl_parameter_set_get_owned_kernel_offset(mlsl_parameter_set p, size_t* v) { C_GUARD(*need(v) = H<ParameterSet>(p)->GetOwnedKernelOffset()) }
int mlsl_parameter_set_get_data_type(mlsl_parameter_set p, mlsl_data_type* v) {
  C_GUARD(*need(v) = (mlsl_data_type)(int)H<ParameterSet>(p)->GetDataType())
}
int mlsl_parameter_set_get_kernel_size(mlsl_parameter_set p, size_t* v) { C_GUARD(*need(v) = H<ParameterSet>(p)->GetKernelSize()) }
int mlsl_parameter_set_is_distributed_update(mlsl_parameter_set p, int* v) { C_GUARD(*need(v) = H<ParameterSet>(p)->IsDistributedUpdate() ? 1 : 0) }
int mlsl_parameter_set_start_gradient_comm(mlsl_parameter_set p, void* buf) { C_GUARD(H<ParameterSet>(p)->StartGradientComm(buf)) }
int mlsl_parameter_set_start_increment_comm(mlsl_parameter_set p, void* buf) { C_GUARD(H<ParameterSet>(p)->StartIncrementComm(buf)) }
int mlsl_parameter_set_wait_gradient_comm(mlsl_parameter_set p, void** v) { C_GUARD(*need(v) = H<ParameterSet>(p)->WaitGradientComm()) }
int mlsl_parameter_set_test_gradient_comm(mlsl_parameter_set p, int* done, void** v) {
  C_GUARD(bool d = false; *need(v) = H<ParameterSet>(p)->TestGradientComm(&d); *need(done) = d ? 1 : 0)
}
int mlsl_parameter_set_wait_increment_comm(mlsl_parameter_set p, void** v) { C_GUARD(*need(v) = H<ParameterSet>(p)->WaitIncrementComm()) }
int mlsl_parameter_set_start_fused_update(mlsl_parameter_set p, void* grad, void* param, mlsl_data_type pt, void* master,
                                          void* s1, void* s2, const mlsl_fused_update_params* o) {
  C_GUARD(MLSL::FusedUpdateParams f; if (!o) throw mlslb::Error("null optimizer params");
          f.type = (MLSL::OptimizerType)o->type; f.lr = o->lr; f.momentum = o->momentum; f.beta1 = o->beta1;
          f.beta2 = o->beta2; f.eps = o->eps; f.weight_decay = o->weight_decay; f.step = o->step;
          f.grad_scale = o->grad_scale; H<ParameterSet>(p)->StartFusedUpdate(grad, param, DT(pt), master, s1, s2, &f))
}
int mlsl_parameter_set_wait_fused_update(mlsl_parameter_set p) { C_GUARD(H<ParameterSet>(p)->WaitFusedUpdate()) }
int mlsl_parameter_set_set_gradient_scale(mlsl_parameter_set p, float scale) { C_GUARD(H<ParameterSet>(p)->SetGradientScale(scale)) }

// ---- Distribution ----
int mlsl_distribution_get_process_count(mlsl_distribution d, mlsl_group_type g, size_t* v) { C_GUARD(*need(v) = H<Distribution>(d)->GetProcessCount(GT(g))) }
int mlsl_distribution_get_process_idx(mlsl_distribution d, mlsl_group_type g, size_t* v) { C_GUARD(*need(v) = H<Distribution>(d)->GetProcessIdx(GT(g))) }
int mlsl_distribution_bcast(mlsl_distribution d, void* buf, size_t n, mlsl_data_type t, size_t root, mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->Bcast(buf, n, DT(t), root, GT(g))))
}
int mlsl_distribution_reduce(mlsl_distribution d, void* s, void* rv, size_t n, mlsl_data_type t, mlsl_reduction_type op,
                             size_t root, mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->Reduce(s, rv, n, DT(t), RT(op), root, GT(g))))
}
int mlsl_distribution_all_reduce(mlsl_distribution d, void* s, void* rv, size_t n, mlsl_data_type t, mlsl_reduction_type op,
                                 mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->AllReduce(s, rv, n, DT(t), RT(op), GT(g))))
}
int mlsl_distribution_all_reduce_ex(mlsl_distribution d, void* s, void* rv, size_t n, mlsl_data_type t, mlsl_reduction_type op,
                                    mlsl_group_type g, float scale, mlsl_compression_type c, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->AllReduceEx(s, rv, n, DT(t), RT(op), GT(g), scale, (MLSL::CompressionType)(int)c)))
}
// [ext] AllReduceEx followed by Environment::Wait in one call (bindings: half the call overhead of the blocking form)
int mlsl_distribution_all_reduce_ex_wait(mlsl_distribution d, mlsl_environment e, void* s, void* rv, size_t n, mlsl_data_type t,
                                         mlsl_reduction_type op, mlsl_group_type g, float scale, mlsl_compression_type c) {
  C_GUARD(H<Environment>(e)->Wait(H<Distribution>(d)->AllReduceEx(s, rv, n, DT(t), RT(op), GT(g), scale, (MLSL::CompressionType)(int)c)))
}
int mlsl_distribution_all_to_all(mlsl_distribution d, void* s, size_t n, void* rv, mlsl_data_type t, mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->AlltoAll(s, n, rv, DT(t), GT(g))))
}
int mlsl_distribution_all_to_allv(mlsl_distribution d, void* s, size_t* sc, size_t* so, void* rv, size_t* rc, size_t* ro,
                                  mlsl_data_type t, mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->AlltoAllv(s, sc, so, rv, rc, ro, DT(t), GT(g))))
}
int mlsl_distribution_send_recv_list(mlsl_distribution d, void* s, size_t* sc, size_t* so, void* rv, size_t* rc, size_t* ro,
                                     mlsl_data_type t, mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->SendRecvList(s, sc, so, rv, rc, ro, DT(t), GT(g))))
}
int mlsl_distribution_gather(mlsl_distribution d, void* s, size_t n, void* rv, mlsl_data_type t, size_t root, mlsl_group_type g,
                             mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->Gather(s, n, rv, DT(t), root, GT(g))))
}
int mlsl_distribution_all_gather(mlsl_distribution d, void* s, size_t n, void* rv, mlsl_data_type t, mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->AllGather(s, n, rv, DT(t), GT(g))))
}
int mlsl_distribution_all_gatherv(mlsl_distribution d, void* s, size_t n, void* rv, size_t* rc, mlsl_data_type t, mlsl_group_type g,
                                  mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->AllGatherv(s, n, rv, rc, DT(t), GT(g))))
}
int mlsl_distribution_scatter(mlsl_distribution d, void* s, void* rv, size_t n, mlsl_data_type t, size_t root, mlsl_group_type g,
                              mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->Scatter(s, rv, n, DT(t), root, GT(g))))
}
int mlsl_distribution_reduce_scatter(mlsl_distribution d, void* s, void* rv, size_t n, mlsl_data_type t, mlsl_reduction_type op,
                                     mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->ReduceScatter(s, rv, n, DT(t), RT(op), GT(g))))
}
int mlsl_distribution_reduce_scatter_ex(mlsl_distribution d, void* s, void* rv, size_t n, mlsl_data_type t, mlsl_reduction_type op,
                                        mlsl_group_type g, float scale, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->ReduceScatterEx(s, rv, n, DT(t), RT(op), GT(g), scale)))
}
int mlsl_distribution_gemm_reduce_scatter(mlsl_distribution d, const void* a, const void* w, void* out, size_t m, size_t n, size_t k,
                                          mlsl_data_type ot, mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->GemmReduceScatter(a, w, out, m, n, k, DT(ot), GT(g))))
}
int mlsl_distribution_all_gather_gemm(mlsl_distribution d, const void* x, const void* w, void* gathered, void* out, size_t m, size_t n,
                                       size_t k, mlsl_data_type ot, mlsl_group_type g, mlsl_comm_req* r) {
  C_GUARD(*need(r) = U(H<Distribution>(d)->AllGatherGemm(x, w, gathered, out, m, n, k, DT(ot), GT(g))))
}
int mlsl_distribution_barrier(mlsl_distribution d, mlsl_group_type g) { C_GUARD(H<Distribution>(d)->Barrier(GT(g))) }
int mlsl_distribution_create_window(mlsl_distribution d, void* base, size_t bytes, mlsl_group_type g, mlsl_window* w) {
  C_GUARD(*need(w) = (mlsl_window)H<Distribution>(d)->CreateWindow(base, bytes, GT(g)))
}
int mlsl_distribution_free_window(mlsl_distribution d, mlsl_window w) { C_GUARD(H<Distribution>(d)->FreeWindow((MLSL::Window*)w)) }
int mlsl_window_put(mlsl_window w, const void* o, size_t n, size_t t, size_t disp) { C_GUARD(((MLSL::Window*)w)->Put(o, n, t, disp)) }
int mlsl_window_get(mlsl_window w, void* o, size_t n, size_t t, size_t disp) { C_GUARD(((MLSL::Window*)w)->Get(o, n, t, disp)) }
int mlsl_window_fence(mlsl_window w) { C_GUARD(((MLSL::Window*)w)->Fence()) }
int mlsl_window_get_size(mlsl_window w, size_t i, size_t* bytes) { C_GUARD(*need(bytes) = ((MLSL::Window*)w)->GetSize(i)) }

// ---- OperationRegInfo ----
int mlsl_operation_reg_info_set_name(mlsl_operation_reg_info i, const char* name) { C_GUARD(H<OperationRegInfo>(i)->SetName(name)) }
int mlsl_operation_reg_info_add_input(mlsl_operation_reg_info i, size_t c, size_t s, mlsl_data_type t) { C_GUARD(H<OperationRegInfo>(i)->AddInput(c, s, DT(t))) }
int mlsl_operation_reg_info_add_output(mlsl_operation_reg_info i, size_t c, size_t s, mlsl_data_type t) { C_GUARD(H<OperationRegInfo>(i)->AddOutput(c, s, DT(t))) }
int mlsl_operation_reg_info_add_parameter_set(mlsl_operation_reg_info i, size_t c, size_t s, mlsl_data_type t, int du) {
  C_GUARD(H<OperationRegInfo>(i)->AddParameterSet(c, s, DT(t), du != 0))
}
int mlsl_operation_reg_info_add_parameter_set_with_compress(mlsl_operation_reg_info i, size_t c, size_t s, mlsl_data_type t, int du,
                                                            mlsl_compression_type ct) {
  C_GUARD(H<OperationRegInfo>(i)->AddParameterSet(c, s, DT(t), du != 0, (MLSL::CompressionType)(int)ct))
}
int mlsl_operation_reg_info_validate(mlsl_operation_reg_info i, mlsl_distribution d) {
  C_GUARD(H<OperationRegInfo>(i)->Validate(d ? H<Distribution>(d) : nullptr))
}

// ---- Operation ----
int mlsl_operation_set_distribution(mlsl_operation o, mlsl_distribution d) { C_GUARD(H<Operation>(o)->SetDistribution(H<Distribution>(d))) }
int mlsl_operation_get_distribution(mlsl_operation o, mlsl_distribution* d) { C_GUARD(*need(d) = U(H<Operation>(o)->GetDistribution())) }
int mlsl_operation_get_session(mlsl_operation o, mlsl_session* s) { C_GUARD(*need(s) = U(H<Operation>(o)->GetSession())) }
int mlsl_operation_get_op_type(mlsl_operation o, mlsl_op_type* t) { C_GUARD(*need(t) = (mlsl_op_type)(int)H<Operation>(o)->GetOpType()) }
int mlsl_operation_set_prev(mlsl_operation o, mlsl_operation prev, size_t a, size_t pa) {
  C_GUARD(H<Operation>(o)->SetPrev(prev ? H<Operation>(prev) : nullptr, a, pa))
}
int mlsl_operation_set_next(mlsl_operation o, mlsl_operation next, size_t a, size_t na) {
  C_GUARD(H<Operation>(o)->SetNext(next ? H<Operation>(next) : nullptr, a, na))
}
int mlsl_operation_get_name(mlsl_operation o, const char** n) { C_GUARD(*need(n) = H<Operation>(o)->GetName()) }
int mlsl_operation_get_global_minibatch_size(mlsl_operation o, size_t* v) { C_GUARD(*need(v) = H<Operation>(o)->GetGlobalMinibatchSize()) }
int mlsl_operation_get_local_minibatch_size(mlsl_operation o, size_t* v) { C_GUARD(*need(v) = H<Operation>(o)->GetLocalMinibatchSize()) }
int mlsl_operation_get_global_minibatch_offset(mlsl_operation o, size_t* v) { C_GUARD(*need(v) = H<Operation>(o)->GetGlobalMinibatchOffset()) }
int mlsl_operation_get_input_count(mlsl_operation o, size_t* v) { C_GUARD(*need(v) = H<Operation>(o)->GetInputCount()) }
int mlsl_operation_get_input(mlsl_operation o, size_t i, mlsl_activation* a) { C_GUARD(*need(a) = U(H<Operation>(o)->GetInput(i))) }
int mlsl_operation_get_output_count(mlsl_operation o, size_t* v) { C_GUARD(*need(v) = H<Operation>(o)->GetOutputCount()) }
int mlsl_operation_get_output(mlsl_operation o, size_t i, mlsl_activation* a) { C_GUARD(*need(a) = U(H<Operation>(o)->GetOutput(i))) }
int mlsl_operation_has_parameter_sets(mlsl_operation o, int* v) { C_GUARD(*need(v) = H<Operation>(o)->HasParameterSets() ? 1 : 0) }
int mlsl_operation_get_parameter_set_count(mlsl_operation o, size_t* v) { C_GUARD(*need(v) = H<Operation>(o)->GetParameterSetCount()) }
int mlsl_operation_get_parameter_set(mlsl_operation o, size_t i, mlsl_parameter_set* p) { C_GUARD(*need(p) = U(H<Operation>(o)->GetParameterSet(i))) }

// ---- Statistics ----
int mlsl_statistics_start(mlsl_statistics s) { C_GUARD(H<Statistics>(s)->Start()) }
int mlsl_statistics_stop(mlsl_statistics s) { C_GUARD(H<Statistics>(s)->Stop()) }
int mlsl_statistics_reset(mlsl_statistics s) { C_GUARD(H<Statistics>(s)->Reset()) }
int mlsl_statistics_print(mlsl_statistics s) { C_GUARD(H<Statistics>(s)->Print()) }
int mlsl_statistics_is_started(mlsl_statistics s, int* v) { C_GUARD(*need(v) = H<Statistics>(s)->IsStarted() ? 1 : 0) }
int mlsl_statistics_is_enabled(mlsl_statistics s, int* v) { C_GUARD(*need(v) = H<Statistics>(s)->IsEnabled() ? 1 : 0) }
int mlsl_statistics_get_isolation_comm_cycles(mlsl_statistics s, size_t i, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetIsolationCommCycles(i)) }
int mlsl_statistics_get_comm_size(mlsl_statistics s, size_t i, size_t* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetCommSize(i)) }
int mlsl_statistics_get_comm_cycles(mlsl_statistics s, size_t i, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetCommCycles(i)) }
int mlsl_statistics_get_compute_cycles(mlsl_statistics s, size_t i, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetComputeCycles(i)) }
int mlsl_statistics_get_total_isolation_comm_cycles(mlsl_statistics s, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetTotalIsolationCommCycles()) }
int mlsl_statistics_get_total_comm_size(mlsl_statistics s, size_t* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetTotalCommSize()) }
int mlsl_statistics_get_total_comm_cycles(mlsl_statistics s, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetTotalCommCycles()) }
int mlsl_statistics_get_total_compute_cycles(mlsl_statistics s, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetTotalComputeCycles()) }
int mlsl_statistics_get_comm_nanos(mlsl_statistics s, size_t i, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetCommNanos(i)) }
int mlsl_statistics_get_device_comm_nanos(mlsl_statistics s, size_t i, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetDeviceCommNanos(i)) }
int mlsl_statistics_get_compute_nanos(mlsl_statistics s, size_t i, unsigned long long* v) { C_GUARD(*need(v) = H<Statistics>(s)->GetComputeNanos(i)) }

// ---- Session ----
int mlsl_session_set_global_minibatch_size(mlsl_session s, size_t n) { C_GUARD(H<Session>(s)->SetGlobalMinibatchSize(n)) }
int mlsl_session_get_global_minibatch_size(mlsl_session s, size_t* v) { C_GUARD(*need(v) = H<Session>(s)->GetGlobalMinibatchSize()) }
int mlsl_session_get_phase_type(mlsl_session s, mlsl_phase_type* v) { C_GUARD(*need(v) = (mlsl_phase_type)(int)H<Session>(s)->GetPhaseType()) }
int mlsl_session_create_operation_reg_info(mlsl_session s, mlsl_op_type t, mlsl_operation_reg_info* i) {
  C_GUARD(*need(i) = U(H<Session>(s)->CreateOperationRegInfo((MLSL::OpType)(int)t)))
}
int mlsl_session_delete_operation_reg_info(mlsl_session s, mlsl_operation_reg_info i) { C_GUARD(H<Session>(s)->DeleteOperationRegInfo(H<OperationRegInfo>(i))) }
int mlsl_session_add_operation_with_distribution(mlsl_session s, mlsl_operation_reg_info i, mlsl_distribution d, size_t* idx) {
  C_GUARD(*need(idx) = H<Session>(s)->AddOperation(H<OperationRegInfo>(i), d ? H<Distribution>(d) : nullptr))
}
int mlsl_session_add_operation(mlsl_session s, mlsl_operation_reg_info i, size_t* idx) {
  C_GUARD(*need(idx) = H<Session>(s)->AddOperation(H<OperationRegInfo>(i), nullptr))
}
int mlsl_session_remove_operations(mlsl_session s) { C_GUARD(H<Session>(s)->RemoveOperations()) }
int mlsl_session_get_operation_count(mlsl_session s, size_t* v) { C_GUARD(*need(v) = H<Session>(s)->GetOperationCount()) }
int mlsl_session_get_operation(mlsl_session s, size_t i, mlsl_operation* o) { C_GUARD(*need(o) = U(H<Session>(s)->GetOperation(i))) }
int mlsl_session_commit(mlsl_session s) { C_GUARD(H<Session>(s)->Commit()) }
int mlsl_session_get_stats(mlsl_session s, mlsl_statistics* v) { C_GUARD(*need(v) = U(H<Session>(s)->GetStats())) }

// ---- Environment ----
int mlsl_environment_get_env(mlsl_environment* e) { C_GUARD(*need(e) = U(&Environment::GetEnv())) }
int mlsl_environment_get_version(int* v) { C_GUARD(*need(v) = Environment::GetVersion()) }
int mlsl_environment_configure(mlsl_environment e, const char* c) { C_GUARD(H<Environment>(e)->Configure(c)) }
int mlsl_environment_init(mlsl_environment e, int* argc, char** argv[]) { C_GUARD(H<Environment>(e)->Init(argc, argv)) }
int mlsl_environment_finalize(mlsl_environment e) { C_GUARD(H<Environment>(e)->Finalize()) }
int mlsl_environment_is_initialized(mlsl_environment e, int* v) { C_GUARD(*need(v) = H<Environment>(e)->IsInitialized() ? 1 : 0) }
int mlsl_environment_get_process_idx(mlsl_environment e, size_t* v) { C_GUARD(*need(v) = H<Environment>(e)->GetProcessIdx()) }
int mlsl_environment_get_process_count(mlsl_environment e, size_t* v) { C_GUARD(*need(v) = H<Environment>(e)->GetProcessCount()) }
int mlsl_environment_create_session(mlsl_environment e, mlsl_phase_type p, mlsl_session* s) {
  C_GUARD(*need(s) = U(H<Environment>(e)->CreateSession((MLSL::PhaseType)(int)p)))
}
int mlsl_environment_delete_session(mlsl_environment e, mlsl_session s) { C_GUARD(H<Environment>(e)->DeleteSession(H<Session>(s))) }
int mlsl_environment_create_distribution(mlsl_environment e, size_t dp, size_t mp, mlsl_distribution* d) {
  C_GUARD(*need(d) = U(H<Environment>(e)->CreateDistribution(dp, mp)))
}
int mlsl_environment_get_group_state(mlsl_environment e, unsigned long long* rows, unsigned long long* mark) {
  C_GUARD(H<Environment>(e)->GetGroupState(need(rows), need(mark)))
}
int mlsl_environment_create_distribution_from_ranks(mlsl_environment e, const size_t* ranks, size_t count,
                                                    unsigned long long rows, unsigned long long mark,
                                                    mlsl_distribution* d) {
  C_GUARD(*need(d) = U(H<Environment>(e)->CreateDistributionFromRanks(ranks, count, rows, mark)))
}
int mlsl_environment_create_distribution_with_colors(mlsl_environment e, int dc, int mc, mlsl_distribution* d) {
  C_GUARD(*need(d) = U(H<Environment>(e)->CreateDistributionWithColors(dc, mc)))
}
int mlsl_environment_delete_distribution(mlsl_environment e, mlsl_distribution d) { C_GUARD(H<Environment>(e)->DeleteDistribution(H<Distribution>(d))) }
int mlsl_environment_wait(mlsl_environment e, mlsl_comm_req r) { C_GUARD(H<Environment>(e)->Wait(H<CommReq>(r))) }
int mlsl_environment_test(mlsl_environment e, mlsl_comm_req r, int* done) {
  C_GUARD(bool d = false; H<Environment>(e)->Test(H<CommReq>(r), &d); *need(done) = d ? 1 : 0)
}
int mlsl_environment_alloc(mlsl_environment e, size_t size, size_t alignment, void** p) { C_GUARD(*need(p) = H<Environment>(e)->Alloc(size, alignment)) }
int mlsl_environment_free(mlsl_environment e, void* p) { C_GUARD(H<Environment>(e)->Free(p)) }
int mlsl_environment_set_quantization_params(mlsl_environment e, mlsl_quant_params* q) {
  C_GUARD(H<Environment>(e)->SetQuantizationParams(reinterpret_cast<MLSL::QuantParams*>(need(q))))
}
int mlsl_environment_get_quantization_params(mlsl_environment e, mlsl_quant_params* q) {
  C_GUARD(MLSL::QuantParams* v = H<Environment>(e)->GetQuantizationParams(); if (!v) throw mlslb::Error("quantization parameters are not set");
          *reinterpret_cast<MLSL::QuantParams*>(need(q)) = *v)
}
int mlsl_environment_set_stream(mlsl_environment e, void* s) { C_GUARD(H<Environment>(e)->SetStream(s)) }
int mlsl_environment_get_stream(mlsl_environment e, void** s) { C_GUARD(*need(s) = H<Environment>(e)->GetStream()) }
int mlsl_environment_set_wait_mode(mlsl_environment e, const char* m) { C_GUARD(H<Environment>(e)->SetWaitMode(m)) }
int mlsl_environment_get_launch_order(mlsl_environment e, long long* uids, size_t cap, size_t* n) { C_GUARD(*need(n) = H<Environment>(e)->GetLaunchOrder(uids, cap)) }
int mlsl_environment_set_tuning(mlsl_environment e, const char* k, long long v) { C_GUARD(H<Environment>(e)->SetTuning(k, (long)v)) }
int mlsl_environment_get_tuning(mlsl_environment e, const char* k, long long* v) { C_GUARD(*need(v) = H<Environment>(e)->GetTuning(k)) }
int mlsl_environment_get_backend_name(mlsl_environment e, const char** n) { C_GUARD(*need(n) = H<Environment>(e)->GetBackendName()) }
int mlsl_environment_describe_backend(mlsl_environment e, const char** n) { C_GUARD(*need(n) = H<Environment>(e)->DescribeBackend()) }
int mlsl_environment_is_device_backend(mlsl_environment e, int* v) { C_GUARD(*need(v) = H<Environment>(e)->IsDeviceBackend() ? 1 : 0) }
int mlsl_environment_suspend_servers(mlsl_environment e) { C_GUARD(H<Environment>(e)->SuspendServers()) }
int mlsl_environment_resume_servers(mlsl_environment e) { C_GUARD(H<Environment>(e)->ResumeServers()) }

// ---- in-process worlds / misc ----
int mlsl_inproc_world_create(int nranks, int* world_id) { C_GUARD(*need(world_id) = mlslb::inproc_world_create(nranks)) }
int mlsl_inproc_world_destroy(int world_id) { C_GUARD(mlslb::inproc_world_destroy(world_id)) }
int mlsl_inproc_bind_thread(int world_id, int rank) { C_GUARD(mlslb::inproc_bind_thread(world_id, rank)) }
int mlsl_inproc_unbind_thread(void) { C_GUARD(mlslb::inproc_unbind_thread()) }
// ---- file-IO offload ----
static mlslb::RankContext* io_ctx() {
  mlslb::RankContext* c = mlslb::current_context();
  if (!c->initialized) throw mlslb::Error("MLSL is not initialized");
  return c;
}
int mlsl_io_open(mlsl_environment, const char* path, mlsl_handle_t* f) { C_GUARD(*need(f) = U(mlslb::io_open(io_ctx(), path))) }
int mlsl_io_size(mlsl_handle_t f, size_t* n) { C_GUARD(*need(n) = mlslb::io_size(H<mlslb::IoFile>(f))) }
int mlsl_io_read_nb(mlsl_handle_t f, void* dst, size_t bytes, long long off, mlsl_handle_t* r) {
  C_GUARD(*need(r) = U(mlslb::io_read_nb(H<mlslb::IoFile>(f), dst, bytes, off)))
}
int mlsl_io_open_read_close_nb(mlsl_environment, const char* path, void* dst, size_t bytes, long long off, mlsl_handle_t* r) {
  C_GUARD(*need(r) = U(mlslb::io_open_read_close_nb(io_ctx(), path, dst, bytes, off)))
}
int mlsl_io_test(mlsl_handle_t r, int* done, size_t* n) { C_GUARD(size_t v = 0; *need(done) = mlslb::io_test(H<mlslb::IoRequest>(r), &v) ? 1 : 0; if (n) *n = v) }
int mlsl_io_wait(mlsl_handle_t r, size_t* n) { C_GUARD(size_t v = mlslb::io_wait(H<mlslb::IoRequest>(r)); if (n) *n = v) }
int mlsl_io_close(mlsl_handle_t f) { C_GUARD(mlslb::io_close(H<mlslb::IoFile>(f))) }
int mlsl_set_assert_throws(int on) { C_GUARD(mlslb::set_assert_throws(on != 0)) }
int mlsl_cuda_available(int* available) { C_GUARD(*need(available) = mlslb::cuda_backend_available() ? 1 : 0) }

}  // extern "C"

// ---- torch.cuda.memory.CUDAPluggableAllocator entry points -------------------------------------------------------------
// `with mlsl_b200.heap_pool():` routes the torch allocations made inside it (DDP / FSDP buckets, optimizer flats, activations
// that are exchanged) to the symmetric heap of the calling rank, so collectives on those tensors are zero-copy - the GPU
// answer to the reference's "stage only what is not in the shared heap" (src/comm_ep.cpp:363-566, EPLIB_memory_is_shmem).
// The signatures are the ones torch expects: (size, device, stream).  A block freed after Finalize is simply dropped.
extern "C" void* mlsl_heap_malloc(ssize_t size, int /*device*/, void* /*stream*/) {
  try {
    MLSL::Environment& e = MLSL::Environment::GetEnv();
    if (!e.IsInitialized()) return nullptr;
    return e.Alloc(size > 0 ? (size_t)size : 1, 512);
  } catch (const std::exception&) {
    return nullptr;
  }
}
extern "C" void mlsl_heap_free(void* ptr, ssize_t /*size*/, int /*device*/, void* /*stream*/) {
  try {
    MLSL::Environment& e = MLSL::Environment::GetEnv();
    if (ptr && e.IsInitialized()) e.Free(ptr);
  } catch (const std::exception&) {
  }
}
