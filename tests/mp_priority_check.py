"""Message prioritisation on GPUs (one rank per GPU, torchrun): six gradient all-reduces are started in backward order
(last layer first); with MLSL_MSG_PRIORITY=1 the progress thread keeps two in flight and launches the NEWEST queued one
next, so the first layer's gradient - started last - overtakes the bulk.  Checks: same launch order on every rank, first
layer done before the middle layers; prints the device-side completion time of every layer.
    MLSL_MSG_PRIORITY=1 torchrun --nproc-per-node N tests/mp_priority_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MLSL_MSG_PRIORITY", "1")
os.environ.setdefault("MLSL_BACKEND", "cuda")
os.environ.setdefault("MLSL_HEAP_SIZE_GB", "3")
import mlsl_b200 as mlsl  # noqa: E402
from mlsl_b200.api import DataType, OperationType  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
env = mlsl.init()
r, W = mlsl.rank(), mlsl.world_size()
layers, n = 6, (64 << 20) // 4
sess = env.create_session()
sess.set_global_minibatch_size(W)
dist = env.create_distribution(W, 1)
ops = []
for l in range(layers):
    ri = sess.create_operation_reg_info(OperationType.CC)
    ri.add_input(8, 1, DataType.FLOAT)
    ri.add_output(8, 1, DataType.FLOAT)
    ri.add_parameter_set(n, 1, DataType.FLOAT, False)
    ops.append(sess.get_operation(sess.add_operation(ri, dist)))
sess.commit()
grads = []
for op in ops:
    g = mlsl.alloc_tensor(n, torch.float32)
    g.fill_(float(r + 1))
    grads.append((op.get_parameter_set(0), g))
ok = True
for it in range(3):
    for _, g in grads:
        g.fill_(float(r + 1))
    mlsl.barrier()
    torch.cuda.synchronize()
    before = len(env.get_launch_order())
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for ps, g in reversed(grads):          # backward: the last layer's gradient is ready first
        ps.start_gradient_comm(g)
    done = []
    for ps, g in grads:                    # the update needs the first layer first
        ps.wait_gradient_comm()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        done.append(e)
    torch.cuda.synchronize()
    order = env.get_launch_order()[before:]
    ok &= all(bool((g == float(W * (W + 1) // 2)).all().item()) for _, g in grads)
    o = torch.tensor(order[:layers] + [-1] * (layers - len(order[:layers])), dtype=torch.int32, device="cuda")
    allo = mlsl.allgather(o).view(W, layers).cpu()
    torch.cuda.synchronize()
    same = bool((allo == allo[0]).all().item())
    ok &= same
    if r == 0:
        print("iteration %d: launch order (layer uids) %s, identical on all ranks: %s" % (it, order, same))
        print("   first layer usable after %.3f ms, all layers after %.3f ms" % (t0.elapsed_time(done[0]), t0.elapsed_time(done[-1])), flush=True)
first_pos = order.index(min(order)) if order else -1
ok &= 0 <= first_pos < layers - 1          # the first layer did not go out last
env.delete_session(sess)
mlsl.finalize()
if r == 0:
    print("mp_priority_check: %s" % ("ALL PASSED" if ok else "FAILED"), flush=True)
sys.exit(0 if ok else 1)
