"""One rank of a net-backend job that exercises the log P routes of small operations: binomial-tree broadcasts from every
root (4 and more members, also not a power of two) and the dissemination barrier."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlsl_b200 as mlsl  # noqa: E402


def vec(seed, n, dtype):
    i = torch.arange(n, dtype=torch.int64)
    return ((i * 11 + seed * 17) % 251 - 125).to(dtype)


def main():
    mlsl.init()
    r, P = mlsl.rank(), mlsl.world_size()
    assert mlsl.env().get_backend_name() == "net"
    checks = 0
    for it, n in enumerate((1, 3, 1000, 8191, 32767)):           # all below the scatter + all-gather size
        for dtype in (torch.float32, torch.int64, torch.uint8):
            for root in range(P):
                want = vec(it * 100 + root, n, dtype)
                b = want.clone() if r == root else torch.zeros(n, dtype=dtype)
                mlsl.bcast(b, root=root)
                assert torch.equal(b, want), ("bcast", n, dtype, root)
                checks += 1
    # back to back without anything in between: a later broadcast's message may arrive before the earlier one is asked for
    bufs = [vec(900 + k, 500 + k, torch.float32) if r == k % P else torch.zeros(500 + k) for k in range(3 * P)]
    works = [mlsl.bcast(b, root=k % P, async_op=True) for k, b in enumerate(bufs)]
    for w in works:
        w.wait()
    for k, b in enumerate(bufs):
        assert torch.equal(b, vec(900 + k, 500 + k, torch.float32)), ("queued bcast", k)
    # barrier: nobody leaves before the last member has arrived
    mlsl.barrier()
    for late in range(P):
        t0 = time.monotonic()
        if r == late:
            time.sleep(0.25)
        mlsl.barrier()
        waited = time.monotonic() - t0
        assert waited > 0.05, ("rank %d left a barrier %.3f s after entering it, %.3f s before rank %d arrived" % (r, waited, 0.25 - waited, late))
        mlsl.barrier()
        checks += 1
    mlsl.finalize()
    print("NET TREE OK rank %d of %d (%d checks)" % (r, P, checks), flush=True)


if __name__ == "__main__":
    main()
