import os
import sys

import pytest


def free_port():
    """A TCP port for a rendezvous on 127.0.0.1: below the kernel's ephemeral range (a port that the outgoing connections of
    an earlier job still hold would make rank 0's listen fail and its peers retry until they time out) and free right now."""
    import random
    import socket
    rng = random.Random()
    for _ in range(200):
        port = rng.randrange(20000, 32000)
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
            try:
                s.bind(("127.0.0.1", port))
                return port
            except OSError:
                continue
    raise RuntimeError("no free port between 20000 and 32000")

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # loopback ranks share one GPU (see mlsl_b200/__init__.py)
# Loop-back ranks are threads of ONE CUDA context.  With lazy module loading the first launch of any kernel (ours or
# torch's) synchronises the context while holding its lock; if a peer rank's kernel is spinning for a third rank whose
# launch now waits for that lock, the job dead-locks until the watchdog fires.  Load every kernel up front instead.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    # MLSL_TEST_DUMP_AFTER=<sec>: dump every thread's Python stack after <sec> seconds of a test (hang diagnosis)
    if os.environ.get("MLSL_TEST_DUMP_AFTER"):
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["MLSL_TEST_DUMP_AFTER"]), repeat=True, file=sys.stderr)

        def _native_dumps():
            import threading
            import time

            def loop():
                from mlsl_b200 import _lib
                while True:
                    time.sleep(float(os.environ["MLSL_TEST_DUMP_AFTER"]))
                    _lib.lib().mlsl_debug_dump_stacks()
            threading.Thread(target=loop, daemon=True).start()
        _native_dumps()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on hosts without a usable CUDA device."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this host")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    """Build the in-tree native library once; the tests never run against a Python fallback."""
    from mlsl_b200 import _lib
    _lib.build()
    yield


def run_ranks(nranks, fn, backend="host", env=None, timeout=300, wait_mode=None):
    """Run fn(rank, mlsl) on `nranks` in-process virtual ranks of the given backend; returns the per-rank results.

    wait_mode="host" (CUDA backend): Wait blocks the calling thread in cudaEventSynchronize / cudaStreamSynchronize until
    the collective is done, so the `.cpu()` / `.item()` that follows never sits in a pageable copy behind a kernel that
    still spins for a peer (such a copy holds a driver lock that blocks the peers' next calls: DESIGN 5b, class 4).
    Tests that start collectives from autograd hooks keep the stream-ordered default (one shared autograd thread)."""
    import mlsl_b200 as mlsl

    old = {}
    new = {"MLSL_BACKEND": backend}
    if backend == "cuda":
        new["MLSL_DEVICE"] = "0"       # loop-back ranks all live on GPU 0, also on a box with several GPUs
    new.update(env or {})
    for k, v in new.items():
        old[k] = os.environ.get(k)
        os.environ[k] = str(v)
    try:
        with mlsl.InprocWorld(nranks) as world:
            def body(r):
                mlsl.bind_thread_state()
                if backend == "cuda":
                    import torch
                    torch.cuda.set_device(0)
                    stream = torch.cuda.Stream()
                    ctx = torch.cuda.stream(stream)
                    ctx.__enter__()
                    # torch creates one cuBLAS handle per thread on the first matmul; cublasCreate synchronises the
                    # device, which dead-locks behind a peer rank's spinning kernel (scripts/probe_blocking_torch.py):
                    # create it now, before the library's init barrier, i.e. before any collective can be in flight
                    for wdt in (torch.float32, torch.bfloat16):
                        wa = torch.ones(16, 16, device="cuda", dtype=wdt)
                        torch.mm(wa, wa)
                    stream.synchronize()
                mlsl.init(wait_mode=wait_mode) if wait_mode else mlsl.init()
                try:
                    res = fn(r, mlsl)
                except BaseException:
                    import traceback
                    sys.stderr.write("rank %d failed:\n%s\n" % (r, traceback.format_exc()))
                    sys.stderr.flush()
                    try:
                        mlsl.finalize()
                    except BaseException:  # noqa: BLE001 - the first error is the one that matters
                        pass
                    raise
                if backend == "cuda":
                    import torch
                    torch.cuda.current_stream().synchronize()
                mlsl.finalize()
                if backend == "cuda":
                    ctx.__exit__(None, None, None)
                return res
            return world.run(body, timeout=timeout)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
