"""[ext] RMA windows on the CUDA backend (kept in its own, last-sorted file: first hardware run happens at round end)."""
import pytest
import torch

from conftest import run_ranks

pytestmark = pytest.mark.gpu


def _gpu(fn, world, **kw):
    kw.setdefault("wait_mode", "host")     # no autograd hooks here: blocking waits keep `.cpu()` off spinning streams
    return run_ranks(world, fn, backend="cuda", env={"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20", **kw.pop("env", {})}, **kw)


def test_rma_window_put_get_fence_device():
    """[ext] one-sided windows on peer-mapped device memory: put into the right neighbour, get from the left one."""
    world, n = 2, 256

    def body(r, mlsl):
        from mlsl_b200.api import GroupType
        d = mlsl.world_distribution()
        mem = mlsl.alloc_tensor(2 * n, torch.float32)
        mem[:n] = float(r)
        win = d.create_window(mem, GroupType.GLOBAL)
        win.fence()
        src = torch.full((n,), 100.0 + r, device="cuda")
        win.put(src, (r + 1) % world, target_disp=n)
        win.fence()
        got_put = mem[n:].clone()
        fetched = torch.zeros(n, device="cuda")
        win.get(fetched, (r - 1) % world, target_disp=0)
        win.fence()
        torch.cuda.current_stream().synchronize()
        out = (got_put.cpu(), fetched.cpu())
        win.free()
        return out

    for r, (got_put, fetched) in enumerate(_gpu(body, world)):
        left = (r - 1) % world
        assert torch.equal(got_put, torch.full((n,), 100.0 + left))
        assert torch.equal(fetched, torch.full((n,), float(left)))
