"""Behavioural checks of the graph / environment API on the host backend: the rules of SURVEY 7.4 and the error
paths the reference asserts on, plus features its own tests never touch (colours, Configure, statistics, pointer
checker, priority lanes, server suspend/resume, golden layouts)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, run_ranks


def _session_two_layers(mlsl, world, model_parts, dist_update, mb=16):
    from mlsl_b200.api import DataType, OperationType
    e = mlsl.env()
    sess = e.create_session()
    sess.set_global_minibatch_size(mb)
    dist = e.create_distribution(world // model_parts, model_parts)
    ops = []
    for l, (ifm, ofm) in enumerate([(128, 256), (256, 256)]):
        ri = sess.create_operation_reg_info(OperationType.CC)
        ri.set_name("layer_%d" % l)
        ri.add_input(ifm, 144, DataType.FLOAT)
        ri.add_output(ofm, 144, DataType.FLOAT)
        ri.add_parameter_set(ifm * ofm, 9, DataType.FLOAT, dist_update)
        ops.append(sess.get_operation(sess.add_operation(ri, dist)))
        sess.delete_operation_reg_info(ri)   # reg info may be dropped right after AddOperation (ref-counted)
    ops[1].set_prev(ops[0], 0, 0)
    sess.commit()
    return sess, dist, ops


def test_golden_layout_2x2_distributed_update():
    """The layout the reference prints for 4 ranks, 2x2, distUpdate=1, rank 0 (SURVEY 2.6 golden dump)."""
    def body(r, mlsl):
        sess, dist, ops = _session_two_layers(mlsl, 4, 2, True)
        oa, ia, ps = ops[0].get_output(0), ops[1].get_input(0), ops[0].get_parameter_set(0)
        blocks = [(b.get_mb_offset(), b.get_mb_count(), b.get_fm_offset(), b.get_fm_count(), b.get_fm_size(), b.get_buf_offset())
                  for b in (oa.get_pack_block(i) for i in range(oa.get_pack_block_count()))]
        res = dict(out_global=oa.get_global_fm_count(), out_local=oa.get_local_fm_count(), blocks=blocks,
                   out_buf=oa.get_comm_buf_size(), in_local=ia.get_local_fm_count(), in_pack=ia.get_pack_block_count(),
                   in_unpack=ia.get_unpack_block_count(), in_buf=ia.get_comm_buf_size(),
                   kernels=(ps.get_global_kernel_count(), ps.get_local_kernel_count(), ps.get_owned_kernel_count(),
                            ps.get_owned_kernel_offset()), lmb=ops[0].get_local_minibatch_size(),
                   mb_off=ops[0].get_global_minibatch_offset())
        mlsl.env().delete_session(sess)
        mlsl.env().delete_distribution(dist)
        return res

    outs = run_ranks(4, body)
    r0 = outs[0]
    assert (r0["out_global"], r0["out_local"]) == (256, 256)
    assert r0["blocks"] == [(0, 8, 0, 128, 144, 0), (0, 8, 128, 128, 144, 147456)]
    assert r0["in_local"] == 128 and r0["in_pack"] == 1 and r0["in_unpack"] == 1
    # reduce-scatter runs out of place here: send region (the reference's tmp_size 1179648) + the receive shard
    assert r0["out_buf"] == 1179648 + 147456 * 4
    assert r0["in_buf"] == 147456 * 2 * 4
    assert r0["kernels"] == (32768, 16384, 8192, 0)
    assert outs[2]["kernels"][3] == 8192 and outs[2]["mb_off"] == 8 and r0["lmb"] == 8


def test_owned_count_is_padded_for_distributed_update():
    def body(r, mlsl):
        from mlsl_b200.api import DataType, OperationType
        e = mlsl.env()
        sess = e.create_session()
        sess.set_global_minibatch_size(3)
        dist = e.create_distribution(3, 1)
        ri = sess.create_operation_reg_info(OperationType.CC)
        ri.add_input(4, 1, DataType.FLOAT)
        ri.add_output(4, 1, DataType.FLOAT)
        ri.add_parameter_set(10, 7, DataType.DOUBLE, True)     # 10 kernels over 3 ranks -> owned 4, local padded to 12
        op = sess.get_operation(sess.add_operation(ri, dist))
        sess.commit()
        ps = op.get_parameter_set(0)
        out = (ps.get_local_kernel_count(), ps.get_owned_kernel_count(), ps.get_owned_kernel_offset(), ps.is_distributed_update())
        e.delete_session(sess)
        e.delete_distribution(dist)
        return out

    outs = run_ranks(3, body)
    assert [o[:3] for o in outs] == [(12, 4, 0), (12, 4, 4), (12, 4, 8)] and all(o[3] for o in outs)


def test_error_paths_raise():
    def body(r, mlsl):
        from mlsl_b200 import MLSLError
        from mlsl_b200.api import DataType, OperationType
        e = mlsl.env()
        errs = []

        def expect(fn, what):
            try:
                fn()
                errs.append("no error for " + what)
            except MLSLError as ex:
                assert "assertion" in str(ex).lower() or "null" in str(ex).lower(), str(ex)

        # the library itself refuses a second Init; the Python object counts init() / finalize() like the reference's binding
        expect(lambda: e._call("mlsl_environment_init", None, None), "double init")
        e.init()
        e.finalize()
        assert e.is_initialized()
        sess = e.create_session()
        expect(lambda: sess.set_global_minibatch_size(0), "zero minibatch")
        sess.set_global_minibatch_size(8)
        expect(lambda: sess.set_global_minibatch_size(8), "minibatch set twice")
        expect(lambda: e.create_distribution(0, 1), "zero partitions")
        expect(lambda: e.create_distribution(3, 3), "more partitions than ranks")
        d = mlsl.world_distribution()
        expect(lambda: d.bcast(torch.zeros(4), 4, DataType.FLOAT, 99, 0), "root outside group")
        ri = sess.create_operation_reg_info(OperationType.SPLIT)
        expect(lambda: sess.add_operation(ri, d), "unsupported op type")
        ri2 = sess.create_operation_reg_info(OperationType.CC)
        expect(lambda: ri2.add_input(0, 4, DataType.FLOAT), "zero feature maps")
        ri2.add_input(4, 4, DataType.FLOAT)
        ri2.add_output(4, 4, DataType.FLOAT)
        op = sess.get_operation(sess.add_operation(ri2, d))
        expect(lambda: op.set_distribution(d), "distribution set twice")
        expect(lambda: op.get_input(5), "input index out of range")
        sess.commit()
        expect(lambda: sess.commit(), "commit twice")
        e.delete_session(sess)
        return errs

    assert run_ranks(2, body) == [[], []]


def test_distribution_with_colors_and_configure():
    def body(r, mlsl):
        e = mlsl.env()
        # rows of a 2x3 grid as data groups, columns as model groups
        d = e.create_distribution_with_colors(r // 3, r % 3)
        info = (d.get_process_count(0), d.get_process_idx(0), d.get_process_count(1), d.get_process_idx(1))
        t = torch.full((8,), float(r))
        mlsl.allreduce(t, group="data", distribution=d)
        u = torch.full((8,), float(r))
        mlsl.allreduce(u, group="model", distribution=d)
        e.delete_distribution(d)
        # Configure("color=N"): the global group itself is split
        e.configure("color=%d" % (r % 2))
        g = (e.get_process_count(), e.get_process_idx())
        d2 = e.create_distribution(e.get_process_count(), 1)
        v = torch.full((4,), float(r))
        mlsl.allreduce(v, group="global", distribution=d2)
        e.delete_distribution(d2)
        return info, float(t[0]), float(u[0]), g, float(v[0])

    outs = run_ranks(6, body)
    for r, (info, t, u, g, v) in enumerate(outs):
        assert info == (3, r % 3, 2, r // 3)
        assert t == sum(q for q in range(6) if q // 3 == r // 3)
        assert u == sum(q for q in range(6) if q % 3 == r % 3)
        assert g == (3, r // 2)
        assert v == sum(q for q in range(6) if q % 2 == r % 2)


def test_statistics_counters():
    def body(r, mlsl):
        sess, dist, ops = _session_two_layers(mlsl, 2, 1, False, mb=4)
        st = sess.get_stats()
        assert st.is_enabled() and not st.is_started()
        iso = st.get_total_isolation_comm_cycles()
        st.start()
        ps = ops[0].get_parameter_set(0)
        n = ps.get_local_kernel_count() * ps.get_kernel_size()
        g = mlsl.alloc_tensor(n, torch.float32)
        for _ in range(3):
            ps.start_gradient_comm(g)
            ps.wait_gradient_comm()
        st.stop()
        res = (iso, st.get_comm_size(0), st.get_comm_cycles(0), st.get_compute_cycles(0), st.get_comm_size(1),
               st.get_total_comm_size(), st.get_comm_nanos(0), n)
        st.print()
        st.reset()
        assert st.get_total_comm_cycles() == 0
        mlsl.env().delete_session(sess)
        mlsl.env().delete_distribution(dist)
        return res

    outs = run_ranks(2, body, env={"MLSL_STATS": "1", "MLSL_STATS_ITERS": "3", "MLSL_STATS_SKIP": "1"})
    for iso, size0, comm0, comp0, size1, total, ns0, n in outs:
        assert iso > 0 and comm0 > 0 and ns0 > 0
        assert size0 == 3 * n * 4 and size1 == 0 and total == size0
    if os.path.exists("mlsl_stats.log"):
        os.remove("mlsl_stats.log")


def test_pointer_checker_rejects_foreign_buffers():
    def body(r, mlsl):
        from mlsl_b200 import MLSLError
        ok = mlsl.alloc_tensor(16, torch.float32)
        mlsl.allreduce(ok)
        try:
            mlsl.allreduce(torch.zeros(16))
            return "foreign buffer accepted"
        except MLSLError as ex:
            return "pointer check" in str(ex)

    assert run_ranks(2, body, env={"MLSL_POINTER_CHECK": "1"}) == [True, True]


@pytest.mark.parametrize("servers", ["0", "2"])
def test_servers_priority_lane_and_suspend_resume(servers):
    """Inline execution (no servers) and two progress threads with the priority lane: same results."""
    def body(r, mlsl):
        sess, dist, ops = _session_two_layers(mlsl, 2, 1, False, mb=4)
        e = mlsl.env()
        grads = []
        for op in ops:
            ps = op.get_parameter_set(0)
            g = mlsl.alloc_tensor(ps.get_local_kernel_count() * ps.get_kernel_size(), torch.float32)
            g.fill_(r + 1.0)
            grads.append((ps, g))
        if servers != "0":
            e.suspend_servers()
        for ps, g in reversed(grads):        # backward order: last layer first
            ps.start_gradient_comm(g)
        if servers != "0":
            e.resume_servers()
        for ps, g in grads:
            ps.wait_gradient_comm()
        ok = all(float(g[0]) == 3.0 and float(g[-1]) == 3.0 for _, g in grads)
        e.delete_session(sess)
        e.delete_distribution(dist)
        return ok

    env = {"MLSL_NUM_SERVERS": servers, "MLSL_MSG_PRIORITY": "1", "MLSL_MSG_PRIORITY_THRESHOLD": "1000"}
    assert run_ranks(2, body, env=env) == [True, True]


def test_message_priority_newest_first_same_order_on_every_rank():
    """MLSL_MSG_PRIORITY=1 with a progress thread: big gradient all-reduces that queue up go out newest first (the first
    layers' gradients, produced last, overtake the bulk - reference eplib/allreduce_pr.c:76-79), small messages go at
    once, and every rank launches the row's collectives in the SAME order (the leader's order log)."""
    world, layers = 4, 6

    def body(r, mlsl):
        from mlsl_b200.api import DataType, OperationType
        e = mlsl.env()
        sess = e.create_session()
        sess.set_global_minibatch_size(world)
        dist = e.create_distribution(world, 1)
        ops = []
        for l in range(layers):
            ri = sess.create_operation_reg_info(OperationType.CC)
            ri.add_input(8, 1, DataType.FLOAT)
            ri.add_output(8, 1, DataType.FLOAT)
            ri.add_parameter_set(4096 if l != 2 else 16, 1, DataType.FLOAT, False)      # layer 2: a small message
            ops.append(sess.get_operation(sess.add_operation(ri, dist)))
        sess.commit()
        grads = []
        for op in ops:
            ps = op.get_parameter_set(0)
            g = mlsl.alloc_tensor(ps.get_local_kernel_count() * ps.get_kernel_size(), torch.float32)
            g.fill_(r + 1.0)
            grads.append((ps, g))
        before = len(e.get_launch_order())
        e.suspend_servers()                       # let the whole backward pass queue up, as a busy device would
        for ps, g in reversed(grads):             # backward order: last layer first
            ps.start_gradient_comm(g)
        e.resume_servers()
        for ps, g in grads:
            ps.wait_gradient_comm()
        order = e.get_launch_order()[before:]
        ok = all(float(g[0]) == sum(range(1, world + 1)) for _, g in grads)
        e.delete_session(sess)
        e.delete_distribution(dist)
        return ok, order

    env = {"MLSL_NUM_SERVERS": "1", "MLSL_MSG_PRIORITY": "1", "MLSL_MSG_PRIORITY_THRESHOLD": "1000", "MLSL_MSG_PRIORITY_MODE": "1"}
    outs = run_ranks(world, body, env=env)
    assert all(ok for ok, _ in outs)
    orders = [o for _, o in outs]
    assert all(o == orders[0] for o in orders), orders            # identical on every rank
    uids = orders[0]
    assert len(uids) == layers
    big = [u for u in uids if u != uids[0]] if False else uids
    # the small message (layer 2, third from the top of the model) went first; among the big ones the newest queued =
    # the FIRST layer's gradient (started last) leads, the last layer's (started first) trails
    small_uid = sorted(uids)[2]
    assert uids[0] == small_uid, uids
    rest = [u for u in uids if u != small_uid]
    assert rest[0] == min(rest) and rest[-1] == max(rest), uids


def test_test_returns_pointer_again_after_completion():
    def body(r, mlsl):
        sess, dist, ops = _session_two_layers(mlsl, 2, 1, False, mb=4)
        ps = ops[0].get_parameter_set(0)
        g = mlsl.alloc_tensor(ps.get_local_kernel_count() * ps.get_kernel_size(), torch.float32)
        ps.start_gradient_comm(g)
        ptr, done = None, False
        while not done:
            ptr, done = ps.test_gradient_comm()
        again, done2 = ps.test_gradient_comm()
        mlsl.env().delete_session(sess)
        mlsl.env().delete_distribution(dist)
        return ptr == g.data_ptr() and again == ptr and done2

    assert run_ranks(2, body) == [True, True]


def test_quantization_params_roundtrip_and_version():
    def body(r, mlsl):
        e = mlsl.env()
        e.set_quantization_params("libdl_comp.so", "q", "d", "rs", 0, 0)
        q = e.get_quantization_params()
        return q["block_size"], q["elem_in_block"], e.get_version()

    assert run_ranks(1, body) == [(132, 128, (1 << 16) | 0)]


# ---- native test programs through the launcher (multi-process, POSIX shared memory) --------------------------------
def _bin(name):
    p = os.path.join(ROOT, "bin", name)
    if not os.path.exists(p):
        subprocess.run(["make", "-C", ROOT, "-j8"], check=True, stdout=subprocess.DEVNULL)
    return p


@pytest.mark.parametrize("args", [["1", "0", "0", "0"], ["2", "1", "1", "1"], ["4", "1", "0", "0"], ["1", "0", "0", "0", "1"]])
def test_cpp_functional_test_multiprocess(args):
    env = dict(os.environ, MLSL_BACKEND="host", MLSL_HEAP_SIZE_GB="0.25")
    res = subprocess.run([_bin("mlslrun"), "-n", "4", "--timeout", "120", _bin("mlsl_functional_test")] + args,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=180)
    assert res.returncode == 0, res.stdout[-2000:]
    assert ": FAILED" not in res.stdout and res.stdout.count("summary: ") == 4 and "0 FAILED" in res.stdout


def test_c_binding_and_sample_multiprocess():
    env = dict(os.environ, MLSL_BACKEND="host", MLSL_HEAP_SIZE_GB="0.25")
    for prog, n in (("cmlsl_smoke_test", 3), ("mlsl_sample", 2), ("mlsl_example", 4)):
        res = subprocess.run([_bin("mlslrun"), "-n", str(n), "--timeout", "60", _bin(prog)], stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, env=env, timeout=120)
        assert res.returncode == 0 and "FAILED" not in res.stdout, res.stdout[-2000:]


@pytest.mark.parametrize("n,servers", [(1, "0"), (3, "0"), (4, "2")])
def test_eplib_entry_points_from_c(n, servers, tmp_path):
    """include/eplib.h: the reference's stand-alone endpoint library surface (init / teardown, the allocation family,
    EPLIB_memory_is_shmem, suspend / execute, file reads on a progress thread) on this runtime (reference eplib/eplib.h)."""
    env = dict(os.environ, MLSL_BACKEND="host", MLSL_HEAP_SIZE_GB="0.25", MLSL_NUM_SERVERS=servers)
    res = subprocess.run([_bin("mlslrun"), "-n", str(n), "--timeout", "60", _bin("cmlsl_eplib_test"), str(tmp_path / "blob.bin")],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.count("PASSED") == n and "FAILED" not in res.stdout, res.stdout[-2000:]


def test_poison_makes_peers_fail_fast():
    """A rank that dies with a fatal signal poisons the shared control block: the survivor errors out at once instead
    of waiting for the watchdog (the reference can only _exit the failing process)."""
    code = r'''
import os, sys, signal, time
sys.path.insert(0, %r)
import torch, mlsl_b200 as mlsl
mlsl.init()
t = torch.ones(8)
mlsl.allreduce(t)
if mlsl.rank() == 1:
    time.sleep(0.3)          # let the peer leave the first collective
    os.kill(os.getpid(), signal.SIGSEGV)
time.sleep(1.0)
t0 = time.time()
try:
    mlsl.allreduce(t)
    print("survived")
except Exception as e:
    print("FAILFAST %%.1f %%s" %% (time.time() - t0, "poisoned" in str(e)))
''' % ROOT
    procs = []
    job = "pz%d" % os.getpid()
    for r in range(2):
        env = dict(os.environ, MLSL_BACKEND="host", MLSL_RANK=str(r), MLSL_WORLD_SIZE="2", MLSL_JOB_ID=job,
                   MLSL_HEAP_SIZE_GB="0.1", MLSL_WATCHDOG_SEC="60")
        procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
    out0 = procs[0].communicate(timeout=120)[0]
    procs[1].communicate(timeout=120)
    assert "FAILFAST" in out0 and "True" in out0, out0
    assert float(out0.split("FAILFAST")[1].split()[0]) < 30


def test_file_io_offload(tmp_path):
    import numpy as np
    data = np.arange(100000, dtype=np.float32)
    path = str(tmp_path / "blob.bin")
    data.tofile(path)

    def body(r, mlsl):
        from mlsl_b200.utils.fileio import File, read_file_nb
        f = File(path)
        assert f.size() == data.nbytes
        a = torch.zeros(1000)
        b = torch.zeros(500)
        ra = f.read_nb(a, offset=4 * 1000 * r)
        rb = read_file_nb(path, b, offset=4 * 50000)
        na, nb = ra.wait(), rb.wait()
        f.close()
        return na, nb, a.clone(), b.clone()

    for r, (na, nb, a, b) in enumerate(run_ranks(2, body)):
        assert (na, nb) == (4000, 2000)
        assert torch.equal(a, torch.arange(1000 * r, 1000 * r + 1000, dtype=torch.float32))
        assert torch.equal(b, torch.arange(50000, 50500, dtype=torch.float32))


def test_failing_inproc_rank_tears_down_cleanly():
    """A rank that raises in the middle of a job: its peers fail fast (poison), every context is torn down without
    crashing the process, the root cause is what the caller sees, and the next world works."""
    import torch
    from conftest import run_ranks

    def body(r, mlsl):
        x = torch.ones(1000)
        mlsl.allreduce(x)
        if r == 1:
            raise ValueError("boom")
        mlsl.allreduce(x)
        return x[0].item()

    with pytest.raises(ValueError, match="boom"):
        run_ranks(2, body, backend="host", env={"MLSL_WATCHDOG_SEC": "3"})

    def body2(r, mlsl):
        x = torch.ones(10)
        mlsl.allreduce(x)
        return x[0].item()

    assert run_ranks(2, body2, backend="host") == [2.0, 2.0]


def test_heap_tensor_lifetime_explicit_free_then_recycled_address():
    """free_tensor() followed by an allocation that recycles the address: the old tensor object dying later must not
    release the new owner's block (the carrier, not the address, owns the block)."""
    import torch
    from conftest import run_ranks

    def body(r, mlsl):
        from mlsl_b200 import comm
        st = comm._state()
        x = mlsl.alloc_tensor(1000, torch.float32)
        p0 = x.data_ptr()
        mlsl.free_tensor(x)
        x2 = mlsl.alloc_tensor(1000, torch.float32)       # recycles p0 while the old `x` object is still alive
        same = x2.data_ptr() == p0
        del x                                             # old object dies now
        y = mlsl.alloc_tensor(1000, torch.float32)
        return same, y.data_ptr() != x2.data_ptr(), x2.data_ptr() in st["live"]

    for same, distinct, live in run_ranks(1, body, backend="host"):
        assert distinct and live


def test_host_heap_grows_on_demand():
    """MLSL_HEAP_SIZE_GB far too small for what the program allocates: the heap registers further shared regions
    (each at least twice the last) and peers attach them lazily when a collective first points into one - the
    reference's EPLIB heap expansion (eplib/memory.c:396-410)."""
    code = r'''
import sys
sys.path.insert(0, %r)
import torch, mlsl_b200 as mlsl
mlsl.init()
r, W = mlsl.rank(), mlsl.world_size()
n = 3 << 20                                   # 12 MiB each, the heap starts with 8 MiB
ts = []
for i in range(4):
    t = mlsl.alloc_tensor(n, torch.float32)
    t.fill_(float(r + 1 + i))
    ts.append(t)
ptrs = [t.data_ptr() for t in ts]
ok = True
for i, t in enumerate(ts):
    out = mlsl.alloc_tensor(n, torch.float32)
    mlsl.allreduce(t, out=out)               # zero copy: peers read this rank's expansion regions directly
    want = sum(p + 1 + i for p in range(W))
    ok = ok and bool((out == want).all()) and float(t[0]) == r + 1 + i
    mlsl.free_tensor(out)
sh = mlsl.reduce_scatter(ts[3])              # library-allocated result in a grown region
ok = ok and bool((sh == sum(p + 1 + 3 for p in range(W))).all())
for t in ts:
    mlsl.free_tensor(t)
t = mlsl.alloc_tensor(n, torch.float32)      # freed space is reused, no further growth needed
ok = ok and t.data_ptr() in ptrs
print("GROW %%s" %% ("OK" if ok else "BAD"), flush=True)
del t, ts, sh, out
mlsl.finalize()
''' % ROOT
    env = dict(os.environ, MLSL_BACKEND="host", MLSL_HEAP_SIZE_GB=str(8.0 / 1024), MLSL_WATCHDOG_SEC="60",
               MLSL_JOB_ID="grow%d" % os.getpid(), MLSL_LOG_LEVEL="1")
    p = subprocess.run([os.path.join(ROOT, "bin", "mlslrun"), "-n", "3", sys.executable, "-c", code], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    assert p.returncode == 0 and p.stdout.count("GROW OK") == 3, p.stdout[-3000:]
    assert "host heap grown" in p.stdout
    leftovers = [f for f in os.listdir("/dev/shm") if "grow%d" % os.getpid() in f]
    assert not leftovers, leftovers


def test_make_install_layout_is_usable(tmp_path):
    """`make install PREFIX=...` + `source intel64/bin/mlslvars.sh`: a C++ program builds with -lmlsl_b200 and runs under
    mlslrun, the Python package imports from the prefix (the reference's packaged layout, scripts/mlslvars.sh)."""
    prefix = str(tmp_path / "inst")
    r = subprocess.run(["make", "-C", ROOT, "install", "PREFIX=" + prefix], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    src = tmp_path / "t.cpp"
    src.write_text('#include <mlsl.hpp>\n#include <cstdio>\nint main(int c, char** v) { MLSL::Environment& e = '
                   'MLSL::Environment::GetEnv(); e.Init(&c, &v); printf("rank %zu of %zu\\n", e.GetProcessIdx(), '
                   'e.GetProcessCount()); e.Finalize(); return 0; }\n')
    script = ("source %s/intel64/bin/mlslvars.sh && cd %s && g++ -std=c++17 t.cpp -o t -lmlsl_b200 && mlslrun -n 2 ./t && "
              "python -c 'import mlsl_b200, os; print(os.path.dirname(mlsl_b200.__file__))'") % (prefix, tmp_path)
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "LD_LIBRARY_PATH")}
    env.update(MLSL_BACKEND="host", MLSL_JOB_ID="inst%d" % os.getpid(), MLSL_HEAP_SIZE_GB="0.05")
    r = subprocess.run(["bash", "-c", script], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "rank 0 of 2" in r.stdout and "rank 1 of 2" in r.stdout and prefix + "/python/mlsl_b200" in r.stdout


def test_rma_window_multiprocess_shared_memory():
    """Windows across real processes: offsets are translated into every peer's mapping of the owner's region."""
    code = r'''
import sys
sys.path.insert(0, %r)
import torch, mlsl_b200 as mlsl
from mlsl_b200.api import GroupType
mlsl.init()
r, W = mlsl.rank(), mlsl.world_size()
mem = mlsl.alloc_tensor(32, torch.float32)
mem.fill_(float(r))
win = mlsl.world_distribution().create_window(mem, GroupType.GLOBAL)
win.fence()
win.put(torch.full((8,), 50.0 + r), (r + 1) %% W, target_disp=8)
win.fence()
got = torch.zeros(16)
win.get(got, (r + 2) %% W, target_disp=0)
win.fence()
ok = bool((mem[8:16] == 50.0 + (r - 1) %% W).all()) and float(mem[0]) == r
o = (r + 2) %% W
ok = ok and bool((got[:8] == o).all()) and bool((got[8:] == 50.0 + (o - 1) %% W).all())
print("RMA %%s" %% ("OK" if ok else "BAD %%s %%s" %% (mem, got)), flush=True)
win.free()
del mem
mlsl.finalize()
''' % ROOT
    env = dict(os.environ, MLSL_BACKEND="host", MLSL_HEAP_SIZE_GB="0.05", MLSL_WATCHDOG_SEC="60", MLSL_JOB_ID="rma%d" % os.getpid())
    p = subprocess.run([os.path.join(ROOT, "bin", "mlslrun"), "-n", "3", sys.executable, "-c", code], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.count("RMA OK") == 3, p.stdout[-2000:]


@pytest.mark.parametrize("args", [["1"], ["4"], ["2", "1"], ["2", "1", "1"]])
def test_c_functional_test_multiprocess(args):
    """The two-layer scenario through the C binding (data / model / hybrid parallel, distributed update, Test polling)."""
    env = dict(os.environ, MLSL_BACKEND="host", MLSL_HEAP_SIZE_GB="0.25")
    res = subprocess.run([_bin("mlslrun"), "-n", "4", "--timeout", "90", _bin("cmlsl_functional_test"), *args],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=150)
    assert res.returncode == 0 and ": FAILED" not in res.stdout and res.stdout.count("0 FAILED") == 4, res.stdout[-2000:]


def test_trace_file_is_valid_chrome_trace(tmp_path):
    """MLSL_TRACE_FILE=<prefix>: every completed request is one slice of <prefix>.<rank>.json (Chrome trace format)."""
    import json
    prefix = str(tmp_path / "trace")

    def body(r, mlsl):
        import torch
        for n in (8, 1000):
            mlsl.allreduce(torch.ones(n))
        mlsl.allgather(torch.ones(16))
        mlsl.barrier()
        return True

    from conftest import run_ranks as rr
    assert rr(2, body, env={"MLSL_TRACE_FILE": prefix}) == [True, True]
    for rank in range(2):
        doc = json.load(open("%s.%d.json" % (prefix, rank)))
        ev = [e for e in doc["traceEvents"] if e["ph"] == "X"]
        names = [e["name"] for e in ev]
        assert names.count("AllReduce") == 2 and "AllGather" in names and "Barrier" in names
        assert all(e["dur"] >= 0 and e["pid"] == rank for e in ev)
        assert sorted(e["args"]["bytes"] for e in ev if e["name"] == "AllReduce") == [32, 4000]


def test_out_of_order_starts_on_different_groups_need_servers():
    """Ranks that start collectives of two groups in opposite orders: fine with one progress server per group in flight
    (rows are spread over the servers), a dead-lock the watchdog reports when collectives run inside Start()."""
    import torch

    def body(r, mlsl):
        a, b = torch.ones(1000), torch.ones(1000) * 2
        first, second = (("global", a), ("data", b)) if r == 0 else (("data", b), ("global", a))
        w1 = mlsl.allreduce(first[1], group=first[0], async_op=True)
        w2 = mlsl.allreduce(second[1], group=second[0], async_op=True)
        w1.wait()
        w2.wait()
        return float(a[0]), float(b[0])

    assert run_ranks(2, body, env={"MLSL_NUM_SERVERS": "2"}) == [(2.0, 4.0), (2.0, 4.0)]
    with pytest.raises(Exception, match="watchdog|poisoned"):
        run_ranks(2, body, env={"MLSL_NUM_SERVERS": "0", "MLSL_WATCHDOG_SEC": "3"})


def test_tuning_knobs_roundtrip_and_unknown_key():
    """Environment.set_tuning / get_tuning: by key or by environment-variable name; unknown keys are an error; the value read
    from the environment at init is visible."""
    def body(r, mlsl):
        from mlsl_b200 import MLSLError
        e = mlsl.env()
        before = e.get_tuning("mid_max_kb")
        e.set_tuning("mid_max_kb", 256)
        by_env_name = e.get_tuning("MLSL_MID_MAX_KB")
        e.set_tuning("MLSL_MID_MAX_KB", before)
        try:
            e.set_tuning("no_such_knob", 1)
            unknown = False
        except MLSLError as ex:
            unknown = "unknown tuning key" in str(ex)
        return before, by_env_name, e.get_tuning("mid_max_kb"), unknown, e.get_tuning("pipe_bufs")

    for before, by_env, after, unknown, pipe in run_ranks(2, body, env={"MLSL_PIPE_BUFS": "6"}):
        assert (before, by_env, after, unknown, pipe) == (1024, 256, 1024, True, 6)


def _py(code, env):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r)\n%s" % (ROOT, code)], env=e, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


def test_reference_environment_knobs_are_recognised():
    """Every variable of the reference's table is parsed: the ones about server processes / MPI have nothing to
    configure here and are listed as such at INFO, MLSL_DYNAMIC_SERVER=disable means no progress threads, EPLIB_UUID names
    the job, MLSL_THP_THRESHOLD_MB makes big host allocations 2 MiB aligned."""
    out = _py("import os, torch, mlsl_b200 as mlsl\n"
              "env = mlsl.init()\n"
              "p = env.alloc(6 << 20, 64); q = env.alloc(1 << 20, 64)\n"
              "print('ALIGN', p % (2 << 20), q % 64)\n"
              "x = torch.ones(8); mlsl.allreduce(x); print('SUM', float(x.sum()))\n"
              "env.free(p); env.free(q); mlsl.finalize()\n",
              {"MLSL_BACKEND": "host", "MLSL_LOG_LEVEL": "1", "MLSL_DYNAMIC_SERVER": "disable", "MLSL_COPY_THREADS": "4",
               "MLSL_SERVER_PREFIX": "numactl", "MLSL_MPI_VERSION_CHECK": "0", "MLSL_THP_THRESHOLD_MB": "4", "EPLIB_UUID": "job-77",
               "MLSL_INPROC_RANKS": "0", "MLSL_NUM_SERVERS": "2"})
    assert "MLSL_DYNAMIC_SERVER=disable" in out and "MLSL_THP_THRESHOLD_MB=4" in out, out[-2000:]
    assert "MLSL_NUM_SERVERS=0" in out, out[-2000:]                   # "disable" wins over the server count
    line = [l for l in out.splitlines() if "not applicable" in l]
    assert line and all(k in line[0] for k in ("MLSL_COPY_THREADS", "MLSL_SERVER_PREFIX", "MLSL_MPI_VERSION_CHECK")), out[-2000:]
    assert "SUM 8.0" in out and "ALIGN 0 0" in out, out[-2000:]


def test_python_init_finalize_nest_and_allow_reinit():
    """The reference's Python module counts init() / finalize() calls and, with MLSL_ALLOW_REINIT=1, holds one initialisation
    of its own until close() (reference include/mlsl/mlsl.py:680-703,1216-1229)."""
    code = ("import os\nfrom mlsl_b200 import api\n"
            "e = api.MLSL(); print('AUTO', e.is_initialized())\n"
            "e.init(); e.init(); e.finalize(); print('NESTED', e.is_initialized())\n"
            "e.finalize(); print('OUTER', e.is_initialized())\n"
            "api.close(); print('CLOSED', e.is_initialized())\n")
    out = _py(code, {"MLSL_BACKEND": "host", "MLSL_ALLOW_REINIT": "1"})
    assert "AUTO True" in out and "NESTED True" in out and "OUTER True" in out and "CLOSED False" in out, out
    out = _py(code, {"MLSL_BACKEND": "host", "MLSL_ALLOW_REINIT": "0"})
    assert "AUTO False" in out and "NESTED True" in out and "OUTER False" in out and "CLOSED False" in out, out
