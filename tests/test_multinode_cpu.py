"""mlsl_b200.parallel.multinode: two-level collectives (this library inside a node, torch.distributed between nodes).
The "nodes" are groups of processes on this machine with torchrun's variables; inside a node the ranks meet in shared
memory under a per-node job id, between nodes over gloo."""
import os
import subprocess
import sys

import pytest

from conftest import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nnodes,per_node", [(2, 2), (3, 1), (1, 3), (2, 3)])
def test_two_level_collectives_and_training(nnodes, per_node):
    world = nnodes * per_node
    port = str(free_port())
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r % per_node),
                   LOCAL_WORLD_SIZE=str(per_node), MASTER_ADDR="127.0.0.1", MASTER_PORT=port, MLSL_BACKEND="host",
                   MLSL_WATCHDOG_SEC="60", OMP_NUM_THREADS="1", GLOO_SOCKET_IFNAME="lo")
        for k in ("MLSL_RANK", "MLSL_WORLD_SIZE", "MLSL_JOB_ID", "MLSL_LOCAL_RANK"):
            env.pop(k, None)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multinode_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=200)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-1500:] for o in outs)
    assert sum("multinode OK" in o for o in outs) == world
