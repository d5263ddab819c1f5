"""One rank of a net-backend job whose reductions travel in pieces (MLSL_NET_CHUNK_KB=4 in the environment: every slice of
8 KiB or more is cut up, reduced piece by piece while later pieces are on the wire, all-reduce results leave piece by piece).
Integer-valued data: every result is exact, so it must match the closed form bit for bit."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlsl_b200 as mlsl  # noqa: E402


def vec(r, n, dtype):
    i = torch.arange(n, dtype=torch.int64)
    return ((i * 7 + r * 13) % 101 - 50).to(dtype)


def main():
    mlsl.init()
    r, P = mlsl.rank(), mlsl.world_size()
    assert mlsl.env().get_backend_name() == "net"
    checks = 0
    for n in (P * 2048, 40001, 262144 + 5, 1 << 20):
        for dtype, op in ((torch.float32, "sum"), (torch.float64, "max"), (torch.bfloat16, "max"), (torch.int32, "sum")):
            every = torch.stack([vec(q, n, dtype) for q in range(P)])
            want = every.sum(0) if op == "sum" else every.max(0).values
            x = vec(r, n, dtype)
            got = mlsl.allreduce(x.clone(), op=op)                       # in place
            assert torch.equal(got, want), ("allreduce in place", n, dtype, op)
            out = torch.full((n,), 3, dtype=dtype)
            mlsl.allreduce(x, op=op, out=out)                            # send -> recv
            assert torch.equal(out, want) and torch.equal(x, vec(r, n, dtype)), ("allreduce out of place", n, dtype, op)
            checks += 2
            if n % P == 0:
                m = n // P
                shard = want[r * m:(r + 1) * m]
                out = torch.zeros(m, dtype=dtype)
                mlsl.reduce_scatter(x, out=out, op=op)
                assert torch.equal(out, shard), ("reduce_scatter", n, dtype, op)
                y = x.clone()
                mlsl.reduce_scatter(y, out=y[:m], op=op)                 # in place: the result lands on slice 0 of the input
                assert torch.equal(y[:m], shard), ("reduce_scatter in place", n, dtype, op)
                z = x.clone()
                mlsl.reduce_scatter(z, out=z[r * m:(r + 1) * m], op=op)  # ... or on the rank's own slice
                assert torch.equal(z[r * m:(r + 1) * m], shard), ("reduce_scatter on the own slice", n, dtype, op)
                checks += 3
    # the fp32 mean with a fused scale (exact: the sums are small integers, P a power of two in the tests)
    x = vec(r, 300000, torch.float32)
    mlsl.allreduce(x, scale=1.0 / P)
    want = torch.stack([vec(q, 300000, torch.float32) for q in range(P)]).sum(0) / P
    assert torch.equal(x, want) if P & (P - 1) == 0 else torch.allclose(x, want, rtol=1e-6, atol=1e-6)
    # data movement (these take the same transports): all-gather, all-to-all, broadcast of a buffer larger than any ring
    n = 70001
    mine = vec(r, n, torch.float32)
    got = mlsl.allgather(mine)
    assert torch.equal(got, torch.cat([vec(q, n, torch.float32) for q in range(P)]))
    for m, dtype in ((9, torch.int32), (4099, torch.float64), (300003, torch.bfloat16)):   # in place: my block sits where it belongs
        full = torch.zeros(P * m, dtype=dtype)
        full[r * m:(r + 1) * m] = vec(r, m, dtype)
        mlsl.allgather(full[r * m:(r + 1) * m], out=full)
        assert torch.equal(full, torch.cat([vec(q, m, dtype) for q in range(P)])), ("allgather in place", m, dtype)
    a2a = mlsl.alltoall(torch.cat([vec(r * P + q, n, torch.int32) for q in range(P)]))
    assert torch.equal(a2a, torch.cat([vec(q * P + r, n, torch.int32) for q in range(P)]))
    b = vec(7, 300001, torch.float64) if r == P - 1 else torch.zeros(300001, dtype=torch.float64)
    mlsl.bcast(b, root=P - 1)
    assert torch.equal(b, vec(7, 300001, torch.float64))
    for root in range(P):                                   # every root: a different column / node feeds the others
        m = 100003 + root
        b = vec(root, m, torch.float32) if r == root else torch.zeros(m)
        mlsl.bcast(b, root=root)
        assert torch.equal(b, vec(root, m, torch.float32)), ("bcast", root)
    if len(sys.argv) > 1 and sys.argv[1] == "describe":
        print(mlsl.env().describe_backend(), flush=True)
    mlsl.barrier()
    mlsl.finalize()
    print("NET CHUNK OK rank %d of %d (%d checks)" % (r, P, checks + 1), flush=True)


if __name__ == "__main__":
    main()
