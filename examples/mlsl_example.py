#!/usr/bin/env python
"""The documented MLSL workflow in a page (cf. csrc/tests/mlsl_example.cpp and the reference's Developer Guide 2.2):
Environment -> Session -> Distribution -> OperationRegInfo -> Operation -> Commit -> per-iteration Start/Wait calls.

    bin/mlslrun -n 4 python examples/mlsl_example.py          # or: python examples/mlsl_example.py --inproc 4
One fully connected layer (256 -> 256 features), data parallel: every rank computes a gradient, the ParameterSet's
persistent request all-reduces it, every rank applies the same update."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsl_b200 as mlsl  # noqa: E402
from mlsl_b200.api import DataType, OperationType  # noqa: E402


def run():
    env = mlsl.init()                                   # Environment::GetEnv().Init()
    rank, world = env.get_process_idx(), env.get_process_count()
    session = env.create_session()                      # one Session per network / phase
    session.set_global_minibatch_size(8 * world)
    dist = env.create_distribution(world, 1)            # world data-parallel replicas, no model parallelism

    reg = session.create_operation_reg_info(OperationType.CC)
    reg.set_name("fc1")
    reg.add_input(256, 1, DataType.FLOAT)               # 256 input feature maps of 1 element
    reg.add_output(256, 1, DataType.FLOAT)
    reg.add_parameter_set(256 * 256, 1, DataType.FLOAT)
    op = session.get_operation(session.add_operation(reg, dist))
    session.delete_operation_reg_info(reg)
    session.commit()                                    # wires the graph, sizes the communication buffers

    params = op.get_parameter_set(0)
    n = params.get_local_kernel_count() * params.get_kernel_size()
    weights = mlsl.alloc_tensor(n, torch.float32)       # symmetric heap: peers reduce straight out of this memory
    grads = mlsl.alloc_tensor(n, torch.float32)
    weights.fill_(1.0)
    for it in range(3):
        grads.fill_(float(rank + 1))                    # "backward": rank r produces gradient r + 1
        params.start_gradient_comm(grads)               # non-blocking; overlap the next layer's backward here
        params.wait_gradient_comm()                     # grads now holds the sum over the data group
        weights -= 0.01 * grads / world
    expect = 1.0 - 3 * 0.01 * (world + 1) / 2
    ok = abs(float(weights[0]) - expect) < 1e-6
    print("[%d] weight after 3 steps: %.6f (expected %.6f) %s" % (rank, float(weights[0]), expect, "PASSED" if ok else "FAILED"),
          flush=True)
    env.delete_session(session)
    env.delete_distribution(dist)
    del weights, grads
    mlsl.finalize()
    return 0 if ok else 1


def main():
    if "--inproc" in sys.argv:
        n = int(sys.argv[sys.argv.index("--inproc") + 1])

        def body(r):
            mlsl.bind_thread_state()
            if torch.cuda.is_available() and os.environ.get("MLSL_BACKEND") == "cuda":
                torch.cuda.set_device(0)                    # loop-back ranks share the GPU: one stream per rank
                torch.cuda.set_stream(torch.cuda.Stream())
            return run()

        with mlsl.InprocWorld(n) as world:
            return 1 if any(world.run(body)) else 0
    return run()


if __name__ == "__main__":
    sys.exit(main())
