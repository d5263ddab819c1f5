"""Python object model over the native runtime.

Same surface as the reference's Python binding (reference include/mlsl/mlsl.py:31-556: MLSL, Session, Distribution,
OperationRegInfo, Operation, Activation, ParameterSet, CommBlockInfo, Statistics and the enum classes, snake_case
method names), plus what that binding lacks (DataType.BYTE, compression on add_parameter_set, all_gatherv,
create_distribution_with_colors, Statistics.print) and the Blackwell extensions.  Buffers may be raw addresses,
objects exposing `data_ptr()` (torch tensors), `__cuda_array_interface__`, or the buffer protocol (numpy).
"""
import ctypes
import threading

from . import _lib
from ._lib import H, c_int, c_size_t, c_ull, c_void_p, check


class DataType:
    FLOAT, DOUBLE, BYTE, BF16, FP16, INT32 = range(6)
    SIZES = {0: 4, 1: 8, 2: 1, 3: 2, 4: 2, 5: 4}


class PhaseType:
    TRAIN, TEST = 0, 1


class GroupType:
    DATA, MODEL, GLOBAL = 0, 1, 2


class ReductionType:
    SUM, MIN, MAX = 0, 1, 2


class OperationType:
    CC, BIAS, ACT, POOL, SPLIT, CONCAT, BCAST, REDUCE, DATA, EVAL = range(10)


class CompressionType:
    NONE, QUANTIZATION = 0, 1


class OptimizerType:
    SGD, ADAMW = 0, 1


def buffer_address(buf):
    """Address of a communication buffer: int, torch tensor, CUDA-array-interface object, ctypes or buffer object."""
    if buf is None:
        return None
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr()
    if hasattr(buf, "__cuda_array_interface__"):
        return buf.__cuda_array_interface__["data"][0]
    if hasattr(buf, "__array_interface__"):
        return buf.__array_interface__["data"][0]
    if isinstance(buf, ctypes.c_void_p):
        return buf.value
    return ctypes.addressof(ctypes.c_char.from_buffer(buf))


def _size_array(values):
    arr = (c_size_t * len(values))(*[int(v) for v in values])
    return arr


# Installed by mlsl_b200.comm.init(): binds the library to torch's CURRENT stream before every call that starts or
# waits for communication, so the raw object API is stream-ordered exactly like the tensor-level functions.
_stream_hook = None


def _pre():
    if _stream_hook is not None:
        _stream_hook()


_FN = {}     # native entry points by name (resolved once)
_INIT_COUNT = {}     # (environment handle, thread) -> init() calls not yet matched by finalize()


class _Handle:
    __slots__ = ("handle",)

    def __init__(self, handle):
        self.handle = handle

    def get_handle(self):
        return self.handle

    def _get(self, fname, ctype=c_size_t, *args):
        out = ctype()
        fn = _FN.get(fname)
        if fn is None:
            fn = _FN[fname] = getattr(_lib.lib(), fname)      # ctypes attribute lookups are not free on a hot path
        check(fn(self.handle, *args, ctypes.byref(out)))
        return out.value

    def _call(self, fname, *args):
        fn = _FN.get(fname)
        if fn is None:
            fn = _FN[fname] = getattr(_lib.lib(), fname)
        check(fn(self.handle, *args))


def _getters(cls, prefix, names, ctype=c_size_t):
    for n in names:
        def make(cname):
            return lambda self: self._get(cname, ctype)
        setattr(cls, "get_" + n, make("%s_get_%s" % (prefix, n)))


class CommBlockInfo(_Handle):
    pass


_getters(CommBlockInfo, "mlsl_comm_block_info", ["mb_offset", "mb_count", "fm_offset", "fm_count", "fm_size", "buf_offset"])
_getters(CommBlockInfo, "mlsl_comm_block_info", ["data_type"], c_int)


class Activation(_Handle):
    def get_pack_block(self, idx):
        return CommBlockInfo(self._get("mlsl_activation_get_pack_block", H, idx))

    def get_unpack_block(self, idx):
        return CommBlockInfo(self._get("mlsl_activation_get_unpack_block", H, idx))

    def get_comm_buf(self):
        return self._get("mlsl_activation_get_comm_buf", c_void_p)

    def start_comm(self, buf):
        _pre()
        self._call("mlsl_activation_start_comm", buffer_address(buf))

    def start_comm_fused(self, local_buf, local_dst=None):
        """Pack + exchange + unpack in one call from the UNPACKED local tensor; with `local_dst` the consumer's unpacked tensor
        is filled directly (the peer's wait_comm returns its address).  All-to-all patterns on the CUDA backend: one kernel."""
        _pre()
        self._call("mlsl_activation_start_comm_fused", buffer_address(local_buf),
                   buffer_address(local_dst) if local_dst is not None else None)

    def wait_comm(self):
        _pre()
        return self._get("mlsl_activation_wait_comm", c_void_p)

    def pack(self, local_buf, comm_buf):
        _pre()
        self._call("mlsl_activation_pack", buffer_address(local_buf), buffer_address(comm_buf))

    def unpack(self, comm_buf, local_buf):
        _pre()
        self._call("mlsl_activation_unpack", buffer_address(comm_buf), buffer_address(local_buf))


_getters(Activation, "mlsl_activation", ["global_fm_count", "global_fm_offset", "local_fm_count", "pack_block_count",
                                        "unpack_block_count", "fm_size", "comm_buf_size"])
_getters(Activation, "mlsl_activation", ["data_type"], c_int)


class ParameterSet(_Handle):
    def is_distributed_update(self):
        return bool(self._get("mlsl_parameter_set_is_distributed_update", c_int))

    def start_gradient_comm(self, buf):
        _pre()
        self._call("mlsl_parameter_set_start_gradient_comm", buffer_address(buf))

    def wait_gradient_comm(self):
        _pre()
        return self._get("mlsl_parameter_set_wait_gradient_comm", c_void_p)

    def test_gradient_comm(self):
        _pre()
        done, ret = c_int(), c_void_p()
        check(_lib.lib().mlsl_parameter_set_test_gradient_comm(self.handle, ctypes.byref(done), ctypes.byref(ret)))
        return ret.value, bool(done.value)

    def start_increment_comm(self, buf):
        _pre()
        self._call("mlsl_parameter_set_start_increment_comm", buffer_address(buf))

    def wait_increment_comm(self):
        _pre()
        return self._get("mlsl_parameter_set_wait_increment_comm", c_void_p)

    def start_fused_update(self, grad, param, param_type, master, state1, state2, opt_type=OptimizerType.SGD, lr=0.0,
                           momentum=0.0, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1, grad_scale=1.0):
        _pre()
        p = _lib.FusedUpdateParams(opt_type, lr, momentum, beta1, beta2, eps, weight_decay, step, grad_scale)
        self._call("mlsl_parameter_set_start_fused_update", buffer_address(grad), buffer_address(param), param_type,
                   buffer_address(master), buffer_address(state1), buffer_address(state2), ctypes.byref(p))

    def wait_fused_update(self):
        _pre()
        self._call("mlsl_parameter_set_wait_fused_update")

    def set_gradient_scale(self, scale):
        self._call("mlsl_parameter_set_set_gradient_scale", float(scale))


_getters(ParameterSet, "mlsl_parameter_set", ["global_kernel_count", "global_kernel_offset", "local_kernel_count",
                                             "owned_kernel_count", "owned_kernel_offset", "kernel_size"])
_getters(ParameterSet, "mlsl_parameter_set", ["data_type"], c_int)


class Distribution(_Handle):
    # instances keep a __dict__ (like before); `_shape` caches (group type) -> (process count, process index), which are
    # fixed for the life of a distribution

    def _group_shape(self, group_type):
        try:
            return self._shape[group_type]
        except AttributeError:
            self._shape = {}
        except KeyError:
            pass
        v = self._shape[group_type] = (self._get("mlsl_distribution_get_process_count", c_size_t, group_type),
                                       self._get("mlsl_distribution_get_process_idx", c_size_t, group_type))
        return v

    def get_process_count(self, group_type):
        return self._group_shape(group_type)[0]

    def get_process_idx(self, group_type):
        return self._group_shape(group_type)[1]

    def _req(self, fname, *args):
        _pre()
        req = H()
        fn = _FN.get(fname)
        if fn is None:
            fn = _FN[fname] = getattr(_lib.lib(), fname)      # ctypes attribute lookups are not free on a hot path
        check(fn(self.handle, *args, ctypes.byref(req)))
        return req.value

    def bcast(self, buf, count, data_type, root_idx, group_type):
        return self._req("mlsl_distribution_bcast", buffer_address(buf), count, data_type, root_idx, group_type)

    def reduce(self, send_buf, recv_buf, count, data_type, red_type, root_idx, group_type):
        return self._req("mlsl_distribution_reduce", buffer_address(send_buf), buffer_address(recv_buf), count,
                         data_type, red_type, root_idx, group_type)

    def all_reduce(self, send_buf, recv_buf, count, data_type, red_type, group_type):
        return self._req("mlsl_distribution_all_reduce", buffer_address(send_buf), buffer_address(recv_buf), count,
                         data_type, red_type, group_type)

    def all_reduce_blocking(self, env, send_buf, recv_buf, count, data_type, red_type, group_type, scale=1.0,
                            compress=CompressionType.NONE):
        """AllReduceEx + Environment::Wait in ONE native call (the tensor-level blocking path: half the binding overhead)."""
        _pre()
        fn = _FN.get("mlsl_distribution_all_reduce_ex_wait")
        if fn is None:
            fn = _FN["mlsl_distribution_all_reduce_ex_wait"] = _lib.lib().mlsl_distribution_all_reduce_ex_wait
        check(fn(self.handle, env.handle, buffer_address(send_buf), buffer_address(recv_buf), count, data_type, red_type,
                 group_type, scale, compress))

    def all_reduce_ex(self, send_buf, recv_buf, count, data_type, red_type, group_type, scale=1.0,
                      compress=CompressionType.NONE):
        return self._req("mlsl_distribution_all_reduce_ex", buffer_address(send_buf), buffer_address(recv_buf), count,
                         data_type, red_type, group_type, scale, compress)

    def all_to_all(self, send_buf, send_count, recv_buf, data_type, group_type):
        return self._req("mlsl_distribution_all_to_all", buffer_address(send_buf), send_count, buffer_address(recv_buf),
                         data_type, group_type)

    def all_to_allv(self, send_buf, send_counts, send_offsets, recv_buf, recv_counts, recv_offsets, data_type, group_type):
        return self._req("mlsl_distribution_all_to_allv", buffer_address(send_buf), _size_array(send_counts),
                         _size_array(send_offsets), buffer_address(recv_buf), _size_array(recv_counts),
                         _size_array(recv_offsets), data_type, group_type)

    def send_recv_list(self, send_buf, send_counts, send_offsets, recv_buf, recv_counts, recv_offsets, data_type, group_type):
        return self._req("mlsl_distribution_send_recv_list", buffer_address(send_buf), _size_array(send_counts),
                         _size_array(send_offsets), buffer_address(recv_buf), _size_array(recv_counts),
                         _size_array(recv_offsets), data_type, group_type)

    def gather(self, send_buf, send_count, recv_buf, data_type, root_idx, group_type):
        return self._req("mlsl_distribution_gather", buffer_address(send_buf), send_count, buffer_address(recv_buf),
                         data_type, root_idx, group_type)

    def all_gather(self, send_buf, send_count, recv_buf, data_type, group_type):
        return self._req("mlsl_distribution_all_gather", buffer_address(send_buf), send_count, buffer_address(recv_buf),
                         data_type, group_type)

    def all_gatherv(self, send_buf, send_count, recv_buf, recv_counts, data_type, group_type):
        return self._req("mlsl_distribution_all_gatherv", buffer_address(send_buf), send_count, buffer_address(recv_buf),
                         _size_array(recv_counts), data_type, group_type)

    def scatter(self, send_buf, recv_buf, recv_count, data_type, root_idx, group_type):
        return self._req("mlsl_distribution_scatter", buffer_address(send_buf), buffer_address(recv_buf), recv_count,
                         data_type, root_idx, group_type)

    def reduce_scatter(self, send_buf, recv_buf, recv_count, data_type, red_type, group_type, scale=1.0):
        return self._req("mlsl_distribution_reduce_scatter_ex", buffer_address(send_buf), buffer_address(recv_buf),
                         recv_count, data_type, red_type, group_type, scale)

    def all_gather_gemm(self, x_shard, w, gathered, out, m, n, k, out_type, group_type):
        return self._req("mlsl_distribution_all_gather_gemm", buffer_address(x_shard), buffer_address(w),
                         buffer_address(gathered), buffer_address(out), m, n, k, out_type, group_type)

    def gemm_reduce_scatter(self, a, w, out, m, n, k, out_type, group_type):
        return self._req("mlsl_distribution_gemm_reduce_scatter", buffer_address(a), buffer_address(w), buffer_address(out),
                         m, n, k, out_type, group_type)

    def create_window(self, tensor, group_type=GroupType.GLOBAL):
        """Collective: expose `tensor` (heap memory: mlsl_b200.alloc_tensor / Environment.alloc) to the group."""
        _pre()
        h = self._get("mlsl_distribution_create_window", H, buffer_address(tensor), tensor.numel() * tensor.element_size(),
                      group_type)
        return Window(h, self, tensor)

    def barrier(self, group_type):
        _pre()
        self._call("mlsl_distribution_barrier", group_type)


class Window(_Handle):
    """[ext] RMA window over memory every member of a group exposed (Distribution.create_window).  `put` / `get`
    address a member by group index and an ELEMENT displacement of the tensor's dtype; both are ordered like the caller's
    other work and complete at the next `fence()` (collective)."""

    def __init__(self, handle, dist, keepalive):
        super().__init__(handle)
        self._dist, self._keep = dist, keepalive

    def put(self, tensor, target_idx, target_disp=0):
        _pre()
        self._call("mlsl_window_put", buffer_address(tensor), tensor.numel() * tensor.element_size(), target_idx,
                   target_disp * tensor.element_size())

    def get(self, tensor, target_idx, target_disp=0):
        _pre()
        self._call("mlsl_window_get", buffer_address(tensor), tensor.numel() * tensor.element_size(), target_idx,
                   target_disp * tensor.element_size())

    def fence(self):
        _pre()
        self._call("mlsl_window_fence")

    def get_size(self, member_idx):
        return self._get("mlsl_window_get_size", c_size_t, member_idx)

    def free(self):
        if self.handle:
            check(_lib.lib().mlsl_distribution_free_window(self._dist.handle, self.handle))
            self.handle, self._keep = None, None


class OperationRegInfo(_Handle):
    def set_name(self, name):
        self._call("mlsl_operation_reg_info_set_name", name.encode())

    def add_input(self, fm_count, fm_size, data_type):
        self._call("mlsl_operation_reg_info_add_input", fm_count, fm_size, data_type)

    def add_output(self, fm_count, fm_size, data_type):
        self._call("mlsl_operation_reg_info_add_output", fm_count, fm_size, data_type)

    def add_parameter_set(self, kernel_count, kernel_size, data_type, dist_update=False, compress=CompressionType.NONE):
        self._call("mlsl_operation_reg_info_add_parameter_set_with_compress", kernel_count, kernel_size, data_type,
                   int(bool(dist_update)), compress)

    def validate(self, dist=None):
        self._call("mlsl_operation_reg_info_validate", dist.handle if dist else 0)


class Operation(_Handle):
    def set_distribution(self, dist):
        self._call("mlsl_operation_set_distribution", dist.handle)

    def get_distribution(self):
        return Distribution(self._get("mlsl_operation_get_distribution", H))

    def get_session(self):
        return Session(self._get("mlsl_operation_get_session", H))

    def get_op_type(self):
        return self._get("mlsl_operation_get_op_type", c_int)

    def set_prev(self, prev_op, act_idx, prev_op_act_idx):
        self._call("mlsl_operation_set_prev", prev_op.handle if prev_op else 0, act_idx, prev_op_act_idx)

    def set_next(self, next_op, act_idx, next_op_act_idx):
        self._call("mlsl_operation_set_next", next_op.handle if next_op else 0, act_idx, next_op_act_idx)

    def get_name(self):
        return self._get("mlsl_operation_get_name", ctypes.c_char_p).decode()

    def get_input(self, idx):
        return Activation(self._get("mlsl_operation_get_input", H, idx))

    def get_output(self, idx):
        return Activation(self._get("mlsl_operation_get_output", H, idx))

    def has_parameter_sets(self):
        return bool(self._get("mlsl_operation_has_parameter_sets", c_int))

    def get_parameter_set(self, idx):
        return ParameterSet(self._get("mlsl_operation_get_parameter_set", H, idx))


_getters(Operation, "mlsl_operation", ["global_minibatch_size", "local_minibatch_size", "global_minibatch_offset",
                                      "input_count", "output_count", "parameter_set_count"])


class Statistics(_Handle):
    def start(self):
        self._call("mlsl_statistics_start")

    def stop(self):
        self._call("mlsl_statistics_stop")

    def reset(self):
        self._call("mlsl_statistics_reset")

    def print(self):
        self._call("mlsl_statistics_print")

    dump = print  # name used by the reference's Python binding

    def is_started(self):
        return bool(self._get("mlsl_statistics_is_started", c_int))

    def is_enabled(self):
        return bool(self._get("mlsl_statistics_is_enabled", c_int))

    def get_isolation_comm_cycles(self, op_idx):
        return self._get("mlsl_statistics_get_isolation_comm_cycles", c_ull, op_idx)

    def get_comm_size(self, op_idx):
        return self._get("mlsl_statistics_get_comm_size", c_size_t, op_idx)

    def get_comm_cycles(self, op_idx):
        return self._get("mlsl_statistics_get_comm_cycles", c_ull, op_idx)

    def get_compute_cycles(self, op_idx):
        return self._get("mlsl_statistics_get_compute_cycles", c_ull, op_idx)

    def get_comm_nanos(self, op_idx):
        return self._get("mlsl_statistics_get_comm_nanos", c_ull, op_idx)

    def get_device_comm_nanos(self, op_idx):
        """Duration of the operation's collectives measured on the device (event pair around each kernel)."""
        return self._get("mlsl_statistics_get_device_comm_nanos", c_ull, op_idx)

    def get_compute_nanos(self, op_idx):
        return self._get("mlsl_statistics_get_compute_nanos", c_ull, op_idx)


_getters(Statistics, "mlsl_statistics", ["total_isolation_comm_cycles", "total_comm_cycles", "total_compute_cycles"], c_ull)
_getters(Statistics, "mlsl_statistics", ["total_comm_size"])


class Session(_Handle):
    def set_global_minibatch_size(self, batch_size):
        self._call("mlsl_session_set_global_minibatch_size", batch_size)

    def get_global_minibatch_size(self):
        return self._get("mlsl_session_get_global_minibatch_size")

    def get_phase_type(self):
        return self._get("mlsl_session_get_phase_type", c_int)

    def create_operation_reg_info(self, op_type):
        return OperationRegInfo(self._get("mlsl_session_create_operation_reg_info", H, op_type))

    def delete_operation_reg_info(self, reg_info):
        self._call("mlsl_session_delete_operation_reg_info", reg_info.handle)

    def add_operation_with_distribution(self, reg_info, dist):
        return self._get("mlsl_session_add_operation_with_distribution", c_size_t, reg_info.handle, dist.handle)

    def add_operation(self, reg_info, dist=None):
        if dist is not None:
            return self.add_operation_with_distribution(reg_info, dist)
        return self._get("mlsl_session_add_operation", c_size_t, reg_info.handle)

    def remove_operations(self):
        self._call("mlsl_session_remove_operations")

    def get_operation_count(self):
        return self._get("mlsl_session_get_operation_count")

    def get_operation(self, op_idx):
        return Operation(self._get("mlsl_session_get_operation", H, op_idx))

    def commit(self):
        self._call("mlsl_session_commit")

    def get_stats(self):
        return Statistics(self._get("mlsl_session_get_stats", H))


class MLSL(_Handle):
    """The Environment.  One per process - or one per thread for in-process virtual ranks."""

    def __init__(self):
        env = H()
        check(_lib.lib().mlsl_environment_get_env(ctypes.byref(env)))
        super().__init__(env.value)

    def get_version(self):
        v = c_int()
        check(_lib.lib().mlsl_environment_get_version(ctypes.byref(v)))
        return v.value

    def configure(self, config):
        self._call("mlsl_environment_configure", config.encode())

    def init(self):
        # Native programs get the host backend unless they ask for MLSL_BACKEND=cuda (Environment::Alloc must keep
        # returning CPU-addressable memory for sources written against the reference).  From Python the tensors say where
        # the data lives, so a visible GPU selects the CUDA backend.
        import os
        if "MLSL_BACKEND" not in os.environ:
            try:
                import torch
                use_cuda = torch.cuda.is_available() and cuda_available()
            except Exception:  # noqa: BLE001
                use_cuda = False
            os.environ["MLSL_BACKEND"] = "cuda" if use_cuda else "host"
        # init / finalize nest like in the reference's binding (reference include/mlsl/mlsl.py:680-703): only the first
        # init and the matching last finalize reach the library
        key = (self.handle, threading.get_ident())
        n = _INIT_COUNT.get(key, 0)
        if n > 0 and self.is_initialized():
            _INIT_COUNT[key] = n + 1
            return
        _INIT_COUNT[key] = 1              # (a left-over count of a run that never finalized does not nest)
        self._call("mlsl_environment_init", None, None)

    def finalize(self):
        key = (self.handle, threading.get_ident())
        n = _INIT_COUNT.get(key, 0)
        if n > 1:
            _INIT_COUNT[key] = n - 1
            return
        _INIT_COUNT.pop(key, None)
        self._call("mlsl_environment_finalize")

    def is_initialized(self):
        return bool(self._get("mlsl_environment_is_initialized", c_int))

    def get_process_idx(self):
        return self._get("mlsl_environment_get_process_idx")

    def get_process_count(self):
        return self._get("mlsl_environment_get_process_count")

    def create_session(self, phase_type=PhaseType.TRAIN):
        return Session(self._get("mlsl_environment_create_session", H, phase_type))

    def delete_session(self, session):
        self._call("mlsl_environment_delete_session", session.handle)

    def create_distribution(self, data_parts, model_parts):
        return Distribution(self._get("mlsl_environment_create_distribution", H, data_parts, model_parts))

    def create_distribution_with_colors(self, data_color, model_color):
        return Distribution(self._get("mlsl_environment_create_distribution_with_colors", H, data_color, model_color))

    def get_group_state(self):
        """(rows in use, ticket mark) of this rank: the two words the members of a new group exchange before
        create_distribution_from_ranks."""
        rows, mark = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        self._call("mlsl_environment_get_group_state", ctypes.byref(rows), ctypes.byref(mark))
        return rows.value, mark.value

    def create_distribution_from_ranks(self, ranks, rows_in_use, ticket_mark):
        """Members-only creation of a distribution whose data group is `ranks` (global process indices, the same list
        in the same order on every member).  `rows_in_use` is the OR and `ticket_mark` the maximum of the members'
        get_group_state() words, exchanged over the caller's own rendezvous (mlsl_b200.torch_backend uses the
        torch.distributed store)."""
        return Distribution(self._get("mlsl_environment_create_distribution_from_ranks", H, _size_array(ranks),
                                      len(ranks), rows_in_use, ticket_mark))

    def delete_distribution(self, dist):
        self._call("mlsl_environment_delete_distribution", dist.handle)

    def wait(self, req):
        _pre()
        self._call("mlsl_environment_wait", req)

    def test(self, req):
        _pre()
        return bool(self._get("mlsl_environment_test", c_int, req))

    def alloc(self, size, alignment=64):
        return self._get("mlsl_environment_alloc", c_void_p, size, alignment)

    def free(self, ptr):
        self._call("mlsl_environment_free", buffer_address(ptr))

    def set_quantization_params(self, lib_path="", quant_name="", dequant_name="", reduce_name="", block_size=0,
                                elem_in_block=0):
        q = _lib.QuantParams(lib_path.encode(), quant_name.encode(), dequant_name.encode(), reduce_name.encode(),
                             block_size, elem_in_block)
        self._call("mlsl_environment_set_quantization_params", ctypes.byref(q))

    def get_quantization_params(self):
        q = _lib.QuantParams()
        self._call("mlsl_environment_get_quantization_params", ctypes.byref(q))
        return {"lib_path": q.lib_path, "block_size": q.block_size, "elem_in_block": q.elem_in_block}

    # ---- extensions ----
    def set_stream(self, stream):
        """stream: torch.cuda.Stream, raw cudaStream_t integer, or None for the backend's own stream."""
        ptr = getattr(stream, "cuda_stream", stream)
        self._call("mlsl_environment_set_stream", ptr)

    def get_stream(self):
        return self._get("mlsl_environment_get_stream", c_void_p)

    def set_wait_mode(self, mode):
        self._call("mlsl_environment_set_wait_mode", mode.encode())

    def get_launch_order(self, capacity=256):
        """Operation uids of the most recently launched collectives, oldest first (progress threads only)."""
        buf = (ctypes.c_longlong * capacity)()
        n = c_size_t()
        check(_lib.lib().mlsl_environment_get_launch_order(self.handle, buf, capacity, ctypes.byref(n)))
        return [int(buf[i]) for i in range(n.value)]

    def set_tuning(self, key, value):
        """Device-path tuning knob by name (see MLSL_LOG_LEVEL=1 for the list); same change on every rank."""
        self._call("mlsl_environment_set_tuning", key.encode(), int(value))

    def get_tuning(self, key):
        return int(self._get("mlsl_environment_get_tuning", ctypes.c_longlong, key.encode()))

    def get_backend_name(self):
        return self._get("mlsl_environment_get_backend_name", ctypes.c_char_p).decode()

    def describe_backend(self):
        return self._get("mlsl_environment_describe_backend", ctypes.c_char_p).decode()

    def is_device_backend(self):
        return bool(self._get("mlsl_environment_is_device_backend", c_int))

    def suspend_servers(self):
        self._call("mlsl_environment_suspend_servers")

    def resume_servers(self):
        self._call("mlsl_environment_resume_servers")


def cuda_available():
    v = c_int()
    check(_lib.lib().mlsl_cuda_available(ctypes.byref(v)))
    return bool(v.value)


class InprocWorld:
    """N virtual ranks inside this process (threads).  Used by the tests and for single-GPU loopback.

    Loop-back ranks on the CUDA backend are threads of ONE CUDA context whose kernels wait for each other, so nothing
    may synchronise that context behind their back: set CUDA_MODULE_LOADING=EAGER (lazy loading synchronises on the
    first launch of every kernel) and CUDA_DEVICE_MAX_CONNECTIONS=32 before CUDA is initialised.  Both are set here
    when the process has not chosen otherwise, which only helps if no CUDA call was made yet."""

    def __init__(self, nranks):
        import os
        os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
        wid = c_int()
        check(_lib.lib().mlsl_inproc_world_create(nranks, ctypes.byref(wid)))
        self.world_id, self.nranks = wid.value, nranks

    def run(self, fn, timeout=300):
        """Run fn(rank) on every virtual rank; re-raises the first failure; returns the list of results."""
        results, errors = [None] * self.nranks, [None] * self.nranks

        def body(r):
            try:
                check(_lib.lib().mlsl_inproc_bind_thread(self.world_id, r))
                try:
                    results[r] = fn(r)
                finally:
                    _lib.lib().mlsl_inproc_unbind_thread()
            except BaseException as e:  # noqa: BLE001 - reported to the caller
                errors[r] = e

        threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(self.nranks)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout)
            if t.is_alive():
                raise TimeoutError("in-process rank did not finish within %s s" % timeout)
        failed = [e for e in errors if e is not None]
        if failed:
            # the root cause first: peers of a failing rank only report "poisoned" / watchdog follow-up errors
            root = [e for e in failed if "poisoned" not in str(e) and "watchdog" not in str(e)]
            dog = [e for e in failed if "watchdog" in str(e)]        # names the signal word a kernel gave up on
            raise (root or dog or failed)[0]
        return results

    def close(self):
        _lib.lib().mlsl_inproc_world_destroy(self.world_id)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


_mlsl_obj = None


def close():
    """Finalize the environment that MLSL_ALLOW_REINIT=1 created at import (reference include/mlsl/mlsl.py:1216-1225)."""
    global _mlsl_obj
    if _mlsl_obj is not None:
        n = _INIT_COUNT.get((_mlsl_obj.handle, threading.get_ident()), 0)
        if n != 1:
            raise RuntimeError("Unexpected reference count for the MLSL object: %d" % n)
        _mlsl_obj.finalize()
        _mlsl_obj = None


def _auto_init():
    # MLSL_ALLOW_REINIT=1: the module owns one initialisation for the life of the process, so that user code may call
    # init() / finalize() any number of times (reference include/mlsl/mlsl.py:1227-1229)
    import os
    global _mlsl_obj
    if os.getenv("MLSL_ALLOW_REINIT") == "1" and _mlsl_obj is None:
        _mlsl_obj = MLSL()
        _mlsl_obj.init()


_auto_init()
