"""Pipeline parallelism on collective neighbour exchanges (mlsl_b200/parallel/pipeline_parallel.py): S stages x M
micro-batches must produce the loss and the parameter gradients of the unsplit model; with a (data x stages)
distribution the stage gradients are additionally averaged over the data group."""
import pytest
import torch

from conftest import run_ranks

D_IN, MB = 12, 4


def _blocks(n):
    g = torch.Generator().manual_seed(21)      # the ranks are threads: the global generator is shared
    blocks = []
    for _ in range(n):
        lin = torch.nn.Linear(D_IN, D_IN)
        with torch.no_grad():
            lin.weight.copy_(torch.randn(D_IN, D_IN, generator=g) * 0.3)
            lin.bias.copy_(torch.randn(D_IN, generator=g) * 0.1)
        blocks.append(torch.nn.Sequential(lin, torch.nn.Tanh()))
    return blocks


def _data(m, replica=0):
    g = torch.Generator().manual_seed(5 + 100 * replica)
    return [torch.randn(MB, D_IN, generator=g) for _ in range(m)], [torch.randn(MB, D_IN, generator=g) for _ in range(m)]


def _reference(stages, m, replicas=1):
    blocks = _blocks(stages)
    model = torch.nn.Sequential(*blocks)
    total = 0.0
    for rep in range(replicas):
        xs, ys = _data(m, rep)
        for x, y in zip(xs, ys):
            total = total + torch.nn.functional.mse_loss(model(x), y) / m / replicas
    total.backward()
    return total.item(), [[p.grad.clone() for p in b.parameters()] for b in blocks]


@pytest.mark.parametrize("schedule", ["gpipe", "1f1b"])
@pytest.mark.parametrize("stages,micro", [(1, 3), (2, 1), (3, 4), (4, 2), (3, 7)])
def test_pipeline_matches_unsplit_model(stages, micro, schedule):
    want_loss, want_grads = _reference(stages, micro)

    def body(r, mlsl):
        from mlsl_b200.parallel.pipeline_parallel import PipelineStage, bubble_fraction
        dist = mlsl.env().create_distribution(1, stages)
        block = _blocks(stages)[r]
        st = PipelineStage(block, (MB, D_IN), (MB, D_IN), group="model", distribution=dist)
        assert (st.is_first, st.is_last) == (r == 0, r == stages - 1)
        xs, ys = _data(micro)
        loss = None
        for _ in range(2):      # a second step reuses nothing from the first (buffers are per step)
            block.zero_grad()
            loss = st.step(xs if st.is_first else None, loss_fn=torch.nn.functional.mse_loss if st.is_last else None,
                           targets=ys if st.is_last else None, num_micro=micro, schedule=schedule)
        if schedule == "1f1b":      # the point of the schedule: live activations are bounded by the depth, not by M
            assert st.max_alive <= min(micro, 2 * (stages - 1 - r) + 1), (r, st.max_alive)
        assert abs(bubble_fraction(stages, micro) - (stages - 1) / (micro + stages - 1)) < 1e-12
        mlsl.env().delete_distribution(dist)
        return (loss.item() if loss is not None else None), [p.grad.clone() for p in block.parameters()]

    res = run_ranks(stages, body)
    assert abs(res[stages - 1][0] - want_loss) < 1e-6
    assert all(res[r][0] is None for r in range(stages - 1))
    for r in range(stages):
        for got, want in zip(res[r][1], want_grads[r]):
            assert torch.allclose(got, want, atol=1e-6, rtol=1e-5), (r, (got - want).abs().max())


def test_pipeline_times_data_parallel():
    """2 replicas x 2 stages: model group = the pipeline, data group = the replicas of one stage."""
    stages, micro, replicas = 2, 3, 2
    want_loss, want_grads = _reference(stages, micro, replicas)

    def body(r, mlsl):
        from mlsl_b200.parallel.pipeline_parallel import PipelineStage
        dist = mlsl.env().create_distribution(replicas, stages)
        s, rep = dist.get_process_idx(mlsl.GroupType.MODEL), dist.get_process_idx(mlsl.GroupType.DATA)
        block = _blocks(stages)[s]
        st = PipelineStage(block, (MB, D_IN), (MB, D_IN), group="model", distribution=dist)
        xs, ys = _data(micro, rep)
        st.step(xs if st.is_first else None, loss_fn=torch.nn.functional.mse_loss if st.is_last else None,
                targets=ys if st.is_last else None, num_micro=micro)
        grads = []
        for p in block.parameters():
            g = p.grad.contiguous()
            mlsl.allreduce(g, scale=1.0 / replicas, group="data", distribution=dist)
            grads.append(g.clone())
        mlsl.env().delete_distribution(dist)
        return s, grads

    for s, grads in run_ranks(stages * replicas, body):
        for got, want in zip(grads, want_grads[s]):
            assert torch.allclose(got, want, atol=1e-6, rtol=1e-5), (s, (got - want).abs().max())


def test_pipeline_argument_checks():
    def body(r, mlsl):
        from mlsl_b200.parallel.pipeline_parallel import PipelineStage
        st = PipelineStage(torch.nn.Linear(D_IN, D_IN), (MB, D_IN), (MB, D_IN), group="data")
        with pytest.raises(ValueError):
            st.step(None, loss_fn=torch.nn.functional.mse_loss, targets=[torch.zeros(MB, D_IN)], num_micro=1)
        with pytest.raises(ValueError):
            st.step([torch.zeros(MB, D_IN)], num_micro=1)
        return True

    assert run_ranks(1, body) == [True]
