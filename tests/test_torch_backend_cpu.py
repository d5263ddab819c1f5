"""torch.distributed backend "mlsl" (mlsl_b200/torch_backend.py): every collective torch.distributed exposes, overlapping
sub-groups made by new_group, destroy of a single group, and DistributedDataParallel against a single-process reference -
all inside tests/torch_backend_worker.py, one process per rank on the host backend."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(n, tmp_path, extra_env=None, worker="torch_backend_worker.py", expect="torch backend OK"):
    env = dict(os.environ, MLSL_BACKEND="host")
    env.update(extra_env or {})
    store = str(tmp_path / "store")
    res = subprocess.run([os.path.join(ROOT, "bin", "mlslrun"), "-n", str(n), "--timeout", "150", sys.executable,
                          os.path.join(ROOT, "tests", worker), store], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=240)
    out = res.stdout.decode(errors="replace")
    assert res.returncode == 0 and expect in out, out[-3000:]


@pytest.mark.parametrize("n", [1, 2, 4])
def test_torch_distributed_on_mlsl(n, tmp_path):
    _run(n, tmp_path)


def test_torch_distributed_on_mlsl_with_progress_servers(tmp_path):
    """The same traffic when requests go through the progress servers (DDP issues its all-reduces from autograd's
    threads while the main thread is inside backward)."""
    _run(3, tmp_path, {"MLSL_NUM_SERVERS": "2"})


@pytest.mark.parametrize("n", [2, 4])
def test_fsdp2_and_hybrid_sharding_on_mlsl(n, tmp_path):
    """`fully_shard` on a 1-D mesh and (4 ranks) replicate x shard on a 2-D DeviceMesh - the mesh dimensions are sub-groups
    made by their members - must train exactly like one process on the whole batch."""
    _run(n, tmp_path, worker="torch_fsdp_worker.py", expect="torch fsdp OK")


def test_members_only_group_creation_api():
    """Environment.get_group_state / create_distribution_from_ranks without torch: the members exchange the two words over
    an all-gather on the world and build overlapping groups; non-members never take part."""
    code = r'''
import os, struct, sys, torch
sys.path.insert(0, %r)
import mlsl_b200 as mlsl
from mlsl_b200 import comm
env = mlsl.init()
rank, world = comm.rank(), comm.world_size()
def make(ranks):
    rows, mark = env.get_group_state()
    mine = torch.frombuffer(bytearray(struct.pack("<QQ", rows, mark)), dtype=torch.uint8)      # 64-bit words travel as bytes
    words = [struct.unpack("<QQ", bytes(w.tolist())) for w in comm.allgather(mine, group="global").view(world, 16)]
    if rank not in ranks:
        return None
    r = 0
    for p in ranks:
        r |= words[p][0]
    return env.create_distribution_from_ranks(ranks, r, max(words[p][1] for p in ranks))
groups = [(l, make(l)) for l in ([0, 2], [1, 2], [0, 1, 2])]
for l, d in groups:
    if d is None:
        continue
    assert d.get_process_count(mlsl.GroupType.DATA) == len(l) and d.get_process_idx(mlsl.GroupType.DATA) == l.index(rank)
    t = torch.full((100,), float(rank + 1))
    comm.allreduce(t, distribution=d)
    assert torch.equal(t, torch.full((100,), float(sum(p + 1 for p in l)))), (rank, l, t[:3])
for l, d in reversed(groups):
    if d is not None:
        env.delete_distribution(d)
try:
    env.create_distribution_from_ranks([(rank + 1) %% world], 0, 0)
    raise SystemExit("a group without the caller must be rejected")
except mlsl.MLSLError as e:
    assert "not a member" in str(e), e
mlsl.finalize()
print("ok")
''' % ROOT
    res = subprocess.run([os.path.join(ROOT, "bin", "mlslrun"), "-n", "3", "--timeout", "60", sys.executable, "-c", code],
                         env=dict(os.environ, MLSL_BACKEND="host"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=120)
    out = res.stdout.decode(errors="replace")
    assert res.returncode == 0 and out.count("ok") == 3, out[-3000:]
