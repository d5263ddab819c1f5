"""mlsl_b200.models.gpt.ParallelTransformerBlock (tensor + sequence parallel, Megatron layout) against the same block
evaluated by one process on the whole sequence with the gathered weights: output rows, input gradient, every weight
gradient (sharded ones slice for slice, replicated ones after the all-reduce the module asks for)."""
import math

import pytest
import torch

from conftest import run_ranks

D, H, FF, M = 32, 4, 64, 16


def _full(D=D, H=H, FF=FF, M=M):
    g = torch.Generator().manual_seed(17)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(x=r(M, D), t=r(M, D), ln1w=1 + 0.1 * r(D), ln1b=0.1 * r(D), ln2w=1 + 0.1 * r(D), ln2b=0.1 * r(D),
                wq=r(D, D) / math.sqrt(D), wk=r(D, D) / math.sqrt(D), wv=r(D, D) / math.sqrt(D), bq=0.1 * r(D), bk=0.1 * r(D),
                bv=0.1 * r(D), wo=r(D, D) / math.sqrt(D), bo=0.1 * r(D), w1=r(FF, D) / math.sqrt(D), b1=0.1 * r(FF),
                w2=r(D, FF) / math.sqrt(FF), b2=0.1 * r(D))


def _reference(D=D, H=H, FF=FF, M=M, cast=None):
    full = _full(D, H, FF, M)
    if cast is not None:      # what a reduced-precision run starts from
        full = {k: v.to(cast).float() for k, v in full.items()}
    p = {k: v.clone().requires_grad_(True) for k, v in full.items() if k != "t"}
    t = full["t"]
    hd = D // H
    x = p["x"]
    h = torch.nn.functional.layer_norm(x, (D,), p["ln1w"], p["ln1b"])
    q, k, v = (h @ p["w" + n].t() + p["b" + n] for n in "qkv")
    q, k, v = (a.view(M, H, hd).permute(1, 0, 2) for a in (q, k, v))
    ctx = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, scale=1 / math.sqrt(hd))
    x1 = x + ctx.permute(1, 0, 2).reshape(M, D) @ p["wo"].t() + p["bo"]
    h2 = torch.nn.functional.layer_norm(x1, (D,), p["ln2w"], p["ln2b"])
    y = x1 + torch.nn.functional.gelu(h2 @ p["w1"].t() + p["b1"]) @ p["w2"].t() + p["b2"]
    ((y - t) ** 2).sum().backward()
    return y.detach(), {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("world", [1, 2, 4])
def test_parallel_transformer_block(world):
    want_y, want = _reference()

    def body(r, mlsl):
        from mlsl_b200.models.gpt import ParallelTransformerBlock
        f = _full()
        dist = mlsl.env().create_distribution(1, world)
        blk = ParallelTransformerBlock(D, H, d_ff=FF, distribution=dist, group="model")
        dl, fl, rows = D // world, FF // world, M // world
        sl, fs, rs = slice(r * dl, (r + 1) * dl), slice(r * fl, (r + 1) * fl), slice(r * rows, (r + 1) * rows)
        with torch.no_grad():
            blk.ln1.weight.copy_(f["ln1w"]), blk.ln1.bias.copy_(f["ln1b"])
            blk.ln2.weight.copy_(f["ln2w"]), blk.ln2.bias.copy_(f["ln2b"])
            blk.qkv.weight.copy_(torch.cat([f["wq"][sl], f["wk"][sl], f["wv"][sl]]))     # my heads: [q | k | v]
            blk.qkv.bias.copy_(torch.cat([f["bq"][sl], f["bk"][sl], f["bv"][sl]]))
            blk.proj.weight.copy_(f["wo"][:, sl]), blk.proj.bias.copy_(f["bo"])
            blk.fc1.weight.copy_(f["w1"][fs]), blk.fc1.bias.copy_(f["b1"][fs])
            blk.fc2.weight.copy_(f["w2"][:, fs]), blk.fc2.bias.copy_(f["b2"])
        x = f["x"][rs].clone().requires_grad_(True)
        y = blk(x)
        ((y - f["t"][rs]) ** 2).sum().backward()
        repl = {}
        for name, prm in (("ln1w", blk.ln1.weight), ("ln1b", blk.ln1.bias), ("ln2w", blk.ln2.weight), ("ln2b", blk.ln2.bias),
                          ("bo", blk.proj.bias), ("b2", blk.fc2.bias)):
            assert any(prm is q for q in blk.layer_norm_parameters())
            g = prm.grad.clone().contiguous()
            mlsl.allreduce(g, group="model", distribution=dist)
            repl[name] = g
        out = dict(y=y.detach(), x=x.grad.clone(), qkv_w=blk.qkv.weight.grad.clone(), qkv_b=blk.qkv.bias.grad.clone(),
                   wo=blk.proj.weight.grad.clone(), w1=blk.fc1.weight.grad.clone(), b1=blk.fc1.bias.grad.clone(),
                   w2=blk.fc2.weight.grad.clone(), **repl)
        mlsl.env().delete_distribution(dist)
        return out

    res = run_ranks(world, body)

    def close(a, b, what, r):
        assert torch.allclose(a, b, atol=2e-4, rtol=2e-4), (what, r, (a - b).abs().max().item())

    for r, o in enumerate(res):
        dl, fl, rows = D // world, FF // world, M // world
        sl, fs, rs = slice(r * dl, (r + 1) * dl), slice(r * fl, (r + 1) * fl), slice(r * rows, (r + 1) * rows)
        close(o["y"], want_y[rs], "y", r)
        close(o["x"], want["x"][rs], "dx", r)
        close(o["qkv_w"], torch.cat([want["wq"][sl], want["wk"][sl], want["wv"][sl]]), "dWqkv", r)
        close(o["qkv_b"], torch.cat([want["bq"][sl], want["bk"][sl], want["bv"][sl]]), "dbqkv", r)
        close(o["wo"], want["wo"][:, sl], "dWo", r)
        close(o["w1"], want["w1"][fs], "dW1", r)
        close(o["b1"], want["b1"][fs], "db1", r)
        close(o["w2"], want["w2"][:, fs], "dW2", r)
        for name in ("ln1w", "ln1b", "ln2w", "ln2b", "bo", "b2"):
            close(o[name], want[name], name, r)
