"""Asynchronous file reads on the runtime's IO thread (the reference's EPLIB_fopen / fread_nb / forc_nb / fwait):
data loading that does not occupy the training thread; the destination may be a host tensor or a CUDA tensor."""
import ctypes

from .. import _lib, comm
from .._lib import H, c_int, c_size_t, check
from ..api import buffer_address


class AsyncRead:
    def __init__(self, req, keep):
        self._req, self._keep = req, keep

    def done(self):
        d, n = c_int(), c_size_t()
        check(_lib.lib().mlsl_io_test(self._req, ctypes.byref(d), ctypes.byref(n)))
        return bool(d.value)

    def wait(self):
        n = c_size_t()
        check(_lib.lib().mlsl_io_wait(self._req, ctypes.byref(n)))
        self._req = None
        return n.value


class File:
    def __init__(self, path):
        h = H()
        check(_lib.lib().mlsl_io_open(comm.env().handle, path.encode(), ctypes.byref(h)))
        self._h = h.value

    def size(self):
        n = c_size_t()
        check(_lib.lib().mlsl_io_size(self._h, ctypes.byref(n)))
        return n.value

    def read_nb(self, tensor, offset=0, nbytes=None):
        nbytes = tensor.numel() * tensor.element_size() if nbytes is None else nbytes
        r = H()
        check(_lib.lib().mlsl_io_read_nb(self._h, buffer_address(tensor), nbytes, offset, ctypes.byref(r)))
        return AsyncRead(r.value, tensor)

    def close(self):
        if self._h:
            check(_lib.lib().mlsl_io_close(self._h))
            self._h = 0


def read_file_nb(path, tensor, offset=0, nbytes=None):
    """open + read + close as one non-blocking request."""
    nbytes = tensor.numel() * tensor.element_size() if nbytes is None else nbytes
    r = H()
    check(_lib.lib().mlsl_io_open_read_close_nb(comm.env().handle, path.encode(), buffer_address(tensor), nbytes, offset,
                                                 ctypes.byref(r)))
    return AsyncRead(r.value, tensor)
