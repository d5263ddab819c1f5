"""ctypes loader for the native runtime (mlsl_b200/lib/libmlsl_b200.so).

The reference's Python binding is also a ctypes layer over its C API (reference include/mlsl/mlsl.py:753-776 locates
libmlsl.so through MLSL_ROOT / LD_LIBRARY_PATH).  Ours loads the in-tree library next to this file, building it on
first use when the toolchain is present, and fails loudly otherwise - there is no pure-Python fallback.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "lib", "libmlsl_b200.so")

_lock = threading.Lock()
_lib = None

c_size_t = ctypes.c_size_t
c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ull = ctypes.c_ulonglong
c_float = ctypes.c_float
c_char_p = ctypes.c_char_p
H = c_ull  # opaque handle


class MLSLError(RuntimeError):
    pass


class FusedUpdateParams(ctypes.Structure):
    _fields_ = [
        ("type", c_int),
        ("lr", c_float),
        ("momentum", c_float),
        ("beta1", c_float),
        ("beta2", c_float),
        ("eps", c_float),
        ("weight_decay", c_float),
        ("step", ctypes.c_longlong),
        ("grad_scale", c_float),
    ]


class QuantParams(ctypes.Structure):
    _fields_ = [
        ("lib_path", c_char_p),
        ("quant_buffer_func_name", c_char_p),
        ("dequant_buffer_func_name", c_char_p),
        ("reduce_sum_func_name", c_char_p),
        ("block_size", c_size_t),
        ("elem_in_block", c_size_t),
    ]


def build(force=False, no_cuda=False):
    """Compile the native library in-tree (make).  Returns the library path."""
    if os.path.exists(LIB_PATH) and not force:
        return LIB_PATH
    cmd = ["make", "-C", _REPO, "-j8"]
    if no_cuda:
        cmd.append("NO_CUDA=1")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0 or not os.path.exists(LIB_PATH):
        raise MLSLError("building libmlsl_b200.so failed:\n" + res.stdout[-4000:])
    return LIB_PATH


def lib():
    """The loaded native library (ctypes.CDLL)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                build()
            l = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
            l.mlsl_last_error.restype = c_char_p
            l.mlsl_set_assert_throws(1)  # failures become exceptions instead of killing the interpreter
            _declare(l)
            _lib = l
    return _lib


def _declare(l):
    P = ctypes.POINTER
    sigs = {
        "mlsl_environment_get_env": [P(H)],
        "mlsl_environment_get_version": [P(c_int)],
        "mlsl_environment_configure": [H, c_char_p],
        "mlsl_environment_init": [H, c_void_p, c_void_p],
        "mlsl_environment_finalize": [H],
        "mlsl_environment_is_initialized": [H, P(c_int)],
        "mlsl_environment_get_process_idx": [H, P(c_size_t)],
        "mlsl_environment_get_process_count": [H, P(c_size_t)],
        "mlsl_environment_create_session": [H, c_int, P(H)],
        "mlsl_environment_delete_session": [H, H],
        "mlsl_environment_create_distribution": [H, c_size_t, c_size_t, P(H)],
        "mlsl_environment_create_distribution_with_colors": [H, c_int, c_int, P(H)],
        "mlsl_environment_delete_distribution": [H, H],
        "mlsl_environment_wait": [H, H],
        "mlsl_environment_test": [H, H, P(c_int)],
        "mlsl_environment_alloc": [H, c_size_t, c_size_t, P(c_void_p)],
        "mlsl_environment_free": [H, c_void_p],
        "mlsl_environment_set_quantization_params": [H, P(QuantParams)],
        "mlsl_environment_get_quantization_params": [H, P(QuantParams)],
        "mlsl_environment_set_stream": [H, c_void_p],
        "mlsl_environment_get_stream": [H, P(c_void_p)],
        "mlsl_environment_set_wait_mode": [H, c_char_p],
        "mlsl_environment_set_tuning": [H, c_char_p, ctypes.c_longlong],
        "mlsl_environment_get_launch_order": [H, ctypes.POINTER(ctypes.c_longlong), c_size_t, P(c_size_t)],
        "mlsl_environment_get_tuning": [H, c_char_p, ctypes.POINTER(ctypes.c_longlong)],
        "mlsl_environment_get_backend_name": [H, P(c_char_p)],
        "mlsl_environment_is_device_backend": [H, P(c_int)],
        "mlsl_environment_describe_backend": [H, P(c_char_p)],
        "mlsl_environment_suspend_servers": [H],
        "mlsl_environment_resume_servers": [H],
        "mlsl_environment_get_group_state": [H, P(ctypes.c_ulonglong), P(ctypes.c_ulonglong)],
        "mlsl_environment_create_distribution_from_ranks": [H, P(c_size_t), c_size_t, ctypes.c_ulonglong,
                                                            ctypes.c_ulonglong, P(H)],
        "mlsl_distribution_get_process_count": [H, c_int, P(c_size_t)],
        "mlsl_distribution_get_process_idx": [H, c_int, P(c_size_t)],
        "mlsl_distribution_bcast": [H, c_void_p, c_size_t, c_int, c_size_t, c_int, P(H)],
        "mlsl_distribution_reduce": [H, c_void_p, c_void_p, c_size_t, c_int, c_int, c_size_t, c_int, P(H)],
        "mlsl_distribution_all_reduce": [H, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, P(H)],
        "mlsl_distribution_all_reduce_ex": [H, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_float, c_int, P(H)],
        "mlsl_distribution_all_reduce_ex_wait": [H, H, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_float, c_int],
        "mlsl_distribution_all_to_all": [H, c_void_p, c_size_t, c_void_p, c_int, c_int, P(H)],
        "mlsl_distribution_all_to_allv": [H, c_void_p, P(c_size_t), P(c_size_t), c_void_p, P(c_size_t), P(c_size_t), c_int, c_int, P(H)],
        "mlsl_distribution_send_recv_list": [H, c_void_p, P(c_size_t), P(c_size_t), c_void_p, P(c_size_t), P(c_size_t), c_int, c_int, P(H)],
        "mlsl_distribution_gather": [H, c_void_p, c_size_t, c_void_p, c_int, c_size_t, c_int, P(H)],
        "mlsl_distribution_all_gather": [H, c_void_p, c_size_t, c_void_p, c_int, c_int, P(H)],
        "mlsl_distribution_all_gatherv": [H, c_void_p, c_size_t, c_void_p, P(c_size_t), c_int, c_int, P(H)],
        "mlsl_distribution_scatter": [H, c_void_p, c_void_p, c_size_t, c_int, c_size_t, c_int, P(H)],
        "mlsl_distribution_reduce_scatter": [H, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, P(H)],
        "mlsl_distribution_reduce_scatter_ex": [H, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_float, P(H)],
        "mlsl_distribution_barrier": [H, c_int],
        "mlsl_distribution_create_window": [H, c_void_p, c_size_t, c_int, P(H)],
        "mlsl_distribution_free_window": [H, H],
        "mlsl_window_put": [H, c_void_p, c_size_t, c_size_t, c_size_t],
        "mlsl_window_get": [H, c_void_p, c_size_t, c_size_t, c_size_t],
        "mlsl_window_fence": [H],
        "mlsl_window_get_size": [H, c_size_t, P(c_size_t)],
        "mlsl_distribution_gemm_reduce_scatter": [H, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_int, P(H)],
        "mlsl_distribution_all_gather_gemm": [H, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_int, P(H)],
        "mlsl_session_set_global_minibatch_size": [H, c_size_t],
        "mlsl_session_get_global_minibatch_size": [H, P(c_size_t)],
        "mlsl_session_get_phase_type": [H, P(c_int)],
        "mlsl_session_create_operation_reg_info": [H, c_int, P(H)],
        "mlsl_session_delete_operation_reg_info": [H, H],
        "mlsl_session_add_operation_with_distribution": [H, H, H, P(c_size_t)],
        "mlsl_session_add_operation": [H, H, P(c_size_t)],
        "mlsl_session_remove_operations": [H],
        "mlsl_session_get_operation_count": [H, P(c_size_t)],
        "mlsl_session_get_operation": [H, c_size_t, P(H)],
        "mlsl_session_commit": [H],
        "mlsl_session_get_stats": [H, P(H)],
        "mlsl_operation_reg_info_set_name": [H, c_char_p],
        "mlsl_operation_reg_info_add_input": [H, c_size_t, c_size_t, c_int],
        "mlsl_operation_reg_info_add_output": [H, c_size_t, c_size_t, c_int],
        "mlsl_operation_reg_info_add_parameter_set": [H, c_size_t, c_size_t, c_int, c_int],
        "mlsl_operation_reg_info_add_parameter_set_with_compress": [H, c_size_t, c_size_t, c_int, c_int, c_int],
        "mlsl_operation_reg_info_validate": [H, H],
        "mlsl_operation_set_distribution": [H, H],
        "mlsl_operation_get_distribution": [H, P(H)],
        "mlsl_operation_get_session": [H, P(H)],
        "mlsl_operation_get_op_type": [H, P(c_int)],
        "mlsl_operation_set_prev": [H, H, c_size_t, c_size_t],
        "mlsl_operation_set_next": [H, H, c_size_t, c_size_t],
        "mlsl_operation_get_name": [H, P(c_char_p)],
        "mlsl_operation_get_global_minibatch_size": [H, P(c_size_t)],
        "mlsl_operation_get_local_minibatch_size": [H, P(c_size_t)],
        "mlsl_operation_get_global_minibatch_offset": [H, P(c_size_t)],
        "mlsl_operation_get_input_count": [H, P(c_size_t)],
        "mlsl_operation_get_input": [H, c_size_t, P(H)],
        "mlsl_operation_get_output_count": [H, P(c_size_t)],
        "mlsl_operation_get_output": [H, c_size_t, P(H)],
        "mlsl_operation_has_parameter_sets": [H, P(c_int)],
        "mlsl_operation_get_parameter_set_count": [H, P(c_size_t)],
        "mlsl_operation_get_parameter_set": [H, c_size_t, P(H)],
        "mlsl_activation_get_global_fm_count": [H, P(c_size_t)],
        "mlsl_activation_get_global_fm_offset": [H, P(c_size_t)],
        "mlsl_activation_get_local_fm_count": [H, P(c_size_t)],
        "mlsl_activation_get_pack_block_count": [H, P(c_size_t)],
        "mlsl_activation_get_unpack_block_count": [H, P(c_size_t)],
        "mlsl_activation_get_pack_block": [H, c_size_t, P(H)],
        "mlsl_activation_get_unpack_block": [H, c_size_t, P(H)],
        "mlsl_activation_get_data_type": [H, P(c_int)],
        "mlsl_activation_get_fm_size": [H, P(c_size_t)],
        "mlsl_activation_get_comm_buf": [H, P(c_void_p)],
        "mlsl_activation_get_comm_buf_size": [H, P(c_size_t)],
        "mlsl_activation_start_comm": [H, c_void_p],
        "mlsl_activation_start_comm_fused": [H, c_void_p, c_void_p],
        "mlsl_activation_wait_comm": [H, P(c_void_p)],
        "mlsl_activation_pack": [H, c_void_p, c_void_p],
        "mlsl_activation_unpack": [H, c_void_p, c_void_p],
        "mlsl_comm_block_info_get_mb_offset": [H, P(c_size_t)],
        "mlsl_comm_block_info_get_mb_count": [H, P(c_size_t)],
        "mlsl_comm_block_info_get_fm_offset": [H, P(c_size_t)],
        "mlsl_comm_block_info_get_fm_count": [H, P(c_size_t)],
        "mlsl_comm_block_info_get_fm_size": [H, P(c_size_t)],
        "mlsl_comm_block_info_get_data_type": [H, P(c_int)],
        "mlsl_comm_block_info_get_buf_offset": [H, P(c_size_t)],
        "mlsl_parameter_set_get_global_kernel_count": [H, P(c_size_t)],
        "mlsl_parameter_set_get_global_kernel_offset": [H, P(c_size_t)],
        "mlsl_parameter_set_get_local_kernel_count": [H, P(c_size_t)],
        "mlsl_parameter_set_get_owned_kernel_count": [H, P(c_size_t)],
        "mlsl_parameter_set_get_owned_kernel_offset": [H, P(c_size_t)],
        "mlsl_parameter_set_get_data_type": [H, P(c_int)],
        "mlsl_parameter_set_get_kernel_size": [H, P(c_size_t)],
        "mlsl_parameter_set_is_distributed_update": [H, P(c_int)],
        "mlsl_parameter_set_start_gradient_comm": [H, c_void_p],
        "mlsl_parameter_set_start_increment_comm": [H, c_void_p],
        "mlsl_parameter_set_wait_gradient_comm": [H, P(c_void_p)],
        "mlsl_parameter_set_test_gradient_comm": [H, P(c_int), P(c_void_p)],
        "mlsl_parameter_set_wait_increment_comm": [H, P(c_void_p)],
        "mlsl_parameter_set_start_fused_update": [H, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, P(FusedUpdateParams)],
        "mlsl_parameter_set_wait_fused_update": [H],
        "mlsl_parameter_set_set_gradient_scale": [H, c_float],
        "mlsl_statistics_start": [H],
        "mlsl_statistics_stop": [H],
        "mlsl_statistics_reset": [H],
        "mlsl_statistics_print": [H],
        "mlsl_statistics_is_started": [H, P(c_int)],
        "mlsl_statistics_is_enabled": [H, P(c_int)],
        "mlsl_statistics_get_isolation_comm_cycles": [H, c_size_t, P(c_ull)],
        "mlsl_statistics_get_comm_size": [H, c_size_t, P(c_size_t)],
        "mlsl_statistics_get_comm_cycles": [H, c_size_t, P(c_ull)],
        "mlsl_statistics_get_compute_cycles": [H, c_size_t, P(c_ull)],
        "mlsl_statistics_get_total_isolation_comm_cycles": [H, P(c_ull)],
        "mlsl_statistics_get_total_comm_size": [H, P(c_size_t)],
        "mlsl_statistics_get_total_comm_cycles": [H, P(c_ull)],
        "mlsl_statistics_get_total_compute_cycles": [H, P(c_ull)],
        "mlsl_statistics_get_comm_nanos": [H, c_size_t, P(c_ull)],
        "mlsl_statistics_get_device_comm_nanos": [H, c_size_t, P(c_ull)],
        "mlsl_statistics_get_compute_nanos": [H, c_size_t, P(c_ull)],
        "mlsl_inproc_world_create": [c_int, P(c_int)],
        "mlsl_inproc_world_destroy": [c_int],
        "mlsl_inproc_bind_thread": [c_int, c_int],
        "mlsl_inproc_unbind_thread": [],
        "mlsl_io_open": [H, c_char_p, P(H)],
        "mlsl_io_size": [H, P(c_size_t)],
        "mlsl_io_read_nb": [H, c_void_p, c_size_t, ctypes.c_longlong, P(H)],
        "mlsl_io_open_read_close_nb": [H, c_char_p, c_void_p, c_size_t, ctypes.c_longlong, P(H)],
        "mlsl_io_test": [H, P(c_int), P(c_size_t)],
        "mlsl_io_wait": [H, P(c_size_t)],
        "mlsl_io_close": [H],
        "mlsl_set_assert_throws": [c_int],
        "mlsl_cuda_available": [P(c_int)],
    }
    for name, argtypes in sigs.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = c_int


def check(rc):
    if rc != 0:
        msg = lib().mlsl_last_error()
        raise MLSLError(msg.decode() if msg else "MLSL call failed")
