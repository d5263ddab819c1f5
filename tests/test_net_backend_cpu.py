"""The net backend (TCP control plane + TCP data mesh) on one machine playing several nodes: every "node" is its own
mlslrun with --nnodes / --node-rank, exactly how a real multi-node job is started."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, free_port

MLSLRUN = os.path.join(ROOT, "bin", "mlslrun")


def _launch(nnodes, per_node, cmd, extra_env=None, timeout=240):
    port = str(free_port())
    env = dict(os.environ, MLSL_WATCHDOG_SEC="60")
    env.pop("MLSL_BACKEND", None)
    env.update(extra_env or {})
    procs = [subprocess.Popen([MLSLRUN, "-n", str(per_node), "--nnodes", str(nnodes), "--node-rank", str(i), "--master-addr",
                               "127.0.0.1", "--master-port", port, "--timeout", str(timeout - 20)] + cmd,
                              cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for i in range(nnodes)]
    outs = [p.communicate(timeout=timeout)[0] for p in procs]
    return [p.returncode for p in procs], "\n".join(outs)


@pytest.mark.parametrize("nnodes,per_node,args", [(2, 2, ["2", "1"]), (2, 2, ["1"]), (4, 1, ["4", "0", "1"]), (2, 2, ["2", "1", "0", "1"]), (2, 4, ["4", "1"])])
def test_cpp_functional_test_across_nodes(nnodes, per_node, args):
    rcs, out = _launch(nnodes, per_node, [os.path.join(ROOT, "bin", "mlsl_functional_test")] + args)
    assert all(rc == 0 for rc in rcs), out[-3000:]
    assert out.count("0 FAILED") == nnodes * per_node and ": FAILED" not in out


@pytest.mark.parametrize("nnodes,per_node,seed,hier_kb", [(2, 2, 3, "1024"), (3, 2, 4, "1024"), (2, 1, 7, "1024"), (2, 2, 5, "0"), (2, 3, 6, "0")])
def test_random_collectives_and_fused_update_across_nodes(nnodes, per_node, seed, hier_kb):
    """(hier_kb 0: every all-reduce / all-gather / reduce-scatter over a group with several members per node takes the
    two-level route, whatever its size - sub-groups of the randomised programs included)"""
    rcs, out = _launch(nnodes, per_node, [sys.executable, os.path.join(ROOT, "tests", "net_worker.py"), str(seed)],
                       extra_env={"MLSL_NET_HIER_KB": hier_kb})
    assert all(rc == 0 for rc in rcs), out[-3000:]
    assert out.count("NET OK") == nnodes * per_node


@pytest.mark.parametrize("shm,ring_kb,expect", [("1", "4", "1 same-node peers"), ("1", "1024", "1 same-node peers"), ("0", "1024", "0 same-node peers")])
def test_same_node_ranks_talk_through_shared_memory(shm, ring_kb, expect):
    """Two launchers = two nodes with two ranks each: the ranks of a node exchange through a shared-memory ring (4 KiB here:
    every larger message wraps around and stalls the writer many times), the nodes over TCP; MLSL_NET_SHM=0 keeps everything
    on the sockets.  Same collectives, same exact results."""
    rcs, out = _launch(2, 2, [sys.executable, os.path.join(ROOT, "tests", "net_chunk_worker.py"), "describe"],
                       extra_env={"MLSL_NET_SHM": shm, "MLSL_NET_SHM_RING_KB": ring_kb})
    assert all(rc == 0 for rc in rcs), out[-3000:]
    assert out.count("NET CHUNK OK") == 4 and out.count(expect) == 4, out[-3000:]


@pytest.mark.parametrize("nnodes,per_node,hier_kb,chunk_kb", [(2, 2, "0", "512"), (2, 3, "0", "4"), (3, 2, "16", "4"), (2, 2, "0", "4"),
                                                               (2, 2, "-1", "512")])
def test_two_level_allreduce_is_exact(nnodes, per_node, hier_kb, chunk_kb):
    """Groups with the same number of members on every node reduce inside the node first (shared memory), exchange 1/L of the
    message between the nodes and gather inside the node again; MLSL_NET_HIER_KB = smallest message that goes that way
    (-1: never); with 4 KiB pieces the all-reduce runs as a pipeline of pieces through its four steps.  Same exact results for
    every size / dtype / in place or not, also with 3 ranks per node or 3 nodes."""
    rcs, out = _launch(nnodes, per_node, [sys.executable, os.path.join(ROOT, "tests", "net_chunk_worker.py")],
                       extra_env={"MLSL_NET_HIER_KB": hier_kb, "MLSL_NET_CHUNK_KB": chunk_kb})
    assert all(rc == 0 for rc in rcs), out[-3000:]
    assert out.count("NET CHUNK OK") == nnodes * per_node


@pytest.mark.parametrize("nnodes,per_node", [(2, 2), (4, 1), (2, 1)])
def test_reductions_in_pieces_are_exact(nnodes, per_node):
    """Large reductions are cut into pieces that are reduced (and, all-reduce, passed on) while the rest is still on the wire;
    4 KiB pieces here, so that every size class takes that path: in place, send -> recv, reduce-scatter onto slice 0 of its
    own input."""
    rcs, out = _launch(nnodes, per_node, [sys.executable, os.path.join(ROOT, "tests", "net_chunk_worker.py")],
                       extra_env={"MLSL_NET_CHUNK_KB": "4", "MLSL_NET_HIER_KB": "-1"})
    assert all(rc == 0 for rc in rcs), out[-3000:]
    assert out.count("NET CHUNK OK") == nnodes * per_node


@pytest.mark.parametrize("nnodes,per_node", [(4, 1), (5, 1), (3, 2), (7, 1)])
def test_small_broadcasts_and_barriers_take_log_p_steps(nnodes, per_node):
    """From four members on a small broadcast runs down a binomial tree and the barrier is a dissemination barrier: every
    root, member counts that are not a power of two, queued broadcasts, and a barrier that holds everybody until the last
    member arrives."""
    rcs, out = _launch(nnodes, per_node, [sys.executable, os.path.join(ROOT, "tests", "net_tree_worker.py")],
                       extra_env={"MLSL_NET_HIER_KB": "-1"})
    assert all(rc == 0 for rc in rcs), out[-3000:]
    assert out.count("NET TREE OK") == nnodes * per_node


def test_interface_selection_like_the_reference():
    """MLSL_IFACE_NAME (prefix) / MLSL_IFACE_IDX pick the interface of the data connections (reference eplib/server.c:228-330)."""
    psutil = pytest.importorskip("psutil")
    import socket
    nics = [(n, a.address) for n, addrs in psutil.net_if_addrs().items() for a in addrs
            if a.family == socket.AF_INET and not a.address.startswith("127.")]
    if not nics:
        pytest.skip("no IPv4 interface besides loop-back")
    name, ip = nics[0]
    code = "import sys; sys.path.insert(0, %r); import mlsl_b200 as mlsl; e = mlsl.init(); print(e.describe_backend()); mlsl.barrier(); mlsl.finalize()" % ROOT
    rcs, out = _launch(2, 1, [sys.executable, "-c", code], extra_env={"MLSL_IFACE_NAME": name[:3]})
    assert all(rc == 0 for rc in rcs) and out.count("data address " + ip) == 2, out[-2000:]
    rcs, out = _launch(2, 1, [sys.executable, "-c", code], extra_env={"MLSL_IFACE_NAME": "nosuchnic"}, timeout=100)
    assert any(rc != 0 for rc in rcs) and "no IPv4 interface matches" in out, out[-2000:]


def test_job_token_in_the_first_frame_of_every_connection():
    """Control and data connections present a hash of MLSL_JOB_TOKEN: with the same token on every node the job runs; a node
    that was given another one is turned away by the control server (and says so) instead of being taken for a member."""
    code = ("import sys; sys.path.insert(0, %r); import torch, mlsl_b200 as mlsl; mlsl.init(); x = torch.ones(4); mlsl.allreduce(x); "
            "print('TOKEN OK %%d' %% int(x[0]), flush=True); mlsl.finalize()") % ROOT
    rcs, out = _launch(2, 2, [sys.executable, "-c", code], extra_env={"MLSL_JOB_TOKEN": "s3cret"})
    assert all(rc == 0 for rc in rcs) and out.count("TOKEN OK 4") == 4, out[-2000:]
    port = str(free_port())
    env = dict(os.environ, MLSL_WATCHDOG_SEC="10")
    env.pop("MLSL_BACKEND", None)
    cmd = lambda i: [MLSLRUN, "-n", "1", "--nnodes", "2", "--node-rank", str(i), "--master-addr", "127.0.0.1", "--master-port", port,
                     "--timeout", "25", sys.executable, "-c", code]
    procs = [subprocess.Popen(cmd(i), cwd=ROOT, env=dict(env, MLSL_JOB_TOKEN="job-%d" % i), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for i in range(2)]
    outs = [p.communicate(timeout=90)[0] for p in procs]
    assert all(p.returncode != 0 for p in procs) and "TOKEN OK" not in "".join(outs)
    assert "another job token" in outs[0], outs[0][-2000:]


def test_a_dying_rank_fails_the_whole_multi_node_job_fast():
    """One rank exits without finalizing: its control connection drops, rank 0's server poisons everyone, the surviving
    ranks leave their collective with an error instead of waiting for the watchdog."""
    code = ("import sys, os, time; sys.path.insert(0, %r); import torch, mlsl_b200 as mlsl; mlsl.init(); t = torch.ones(4); "
            "mlsl.allreduce(t); r = mlsl.rank();\n"
            "if r == 3: os._exit(7)\n"
            "t0 = time.time()\n"
            "try:\n"
            "    mlsl.allreduce(t); print('survived')\n"
            "except Exception as e:\n"
            "    print('FAILFAST %%.1f %%s' %% (time.time() - t0, 'poisoned' in str(e) or 'closed' in str(e) or 'broke' in str(e)))\n") % ROOT
    rcs, out = _launch(2, 2, [sys.executable, "-c", code], timeout=120)
    import re
    hits = re.findall(r"FAILFAST ([0-9.]+) (True|False)", out)       # ranks share a pipe: lines may run together
    assert hits and "survived" not in out, out[-2000:]
    assert all(float(t) < 30 and ok == "True" for t, ok in hits), hits


def test_torch_distributed_backend_across_nodes(tmp_path):
    """The torch.distributed "mlsl" backend (collectives, member-made sub-groups, DDP) with the ranks spread over two
    "nodes": the same worker as tests/test_torch_backend_cpu.py, traffic on the TCP mesh."""
    rcs, out = _launch(2, 2, [sys.executable, os.path.join(ROOT, "tests", "torch_backend_worker.py"), str(tmp_path / "store")])
    assert all(rc == 0 for rc in rcs) and "torch backend OK" in out, out[-3000:]


def test_hosts_launcher_starts_every_node_through_the_remote_shell():
    """`mlslrun --hosts a,b` is the head of the job (mpiexec.hydra -hosts): it starts one mlslrun per host through the remote
    shell - here a stand-in for ssh that runs the command locally - and forwards -e variables; a failing node stops the job."""
    rsh = "sh " + os.path.join(ROOT, "tests", "fake_rsh.sh")     # a remote shell command with an argument of its own
    env = dict(os.environ, MLSL_WATCHDOG_SEC="60")
    env.pop("MLSL_BACKEND", None)
    p = subprocess.run([MLSLRUN, "-n", "2", "--hosts", "127.0.0.1,127.0.0.1", "--rsh", rsh, "--timeout", "120", "-e", "MARK=from head",
                        sys.executable, "-c",
                        "import os, sys; sys.path.insert(0, %r); import torch, mlsl_b200 as mlsl; mlsl.init(); t = torch.ones(3); "
                        "mlsl.allreduce(t); sys.stdout.write('rank %%d of %%d sum %%d %%s %%s\\n' %% (mlsl.rank(), mlsl.world_size(), int(t[0]), "
                        "os.environ['MARK'], mlsl.env().get_backend_name())); sys.stdout.flush(); mlsl.finalize()" % ROOT],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=200)
    assert p.returncode == 0, p.stdout[-2000:]
    for r in range(4):
        assert "rank %d of 4 sum 4 from head net" % r in p.stdout, p.stdout[-2000:]
    bad = subprocess.run([MLSLRUN, "-n", "1", "--hosts", "127.0.0.1,127.0.0.1", "--rsh", rsh, "--timeout", "60", sys.executable, "-c",
                          "import os, sys, time; sys.exit(5) if os.environ['RANK'] == '1' else time.sleep(30)"],
                         cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=100)
    assert bad.returncode == 5 and "stopping the job" in bad.stdout, bad.stdout[-1500:]
