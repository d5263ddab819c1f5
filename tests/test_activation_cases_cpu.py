"""All five activation-exchange patterns between two layers that use DIFFERENT distributions (reference
src/mlsl_impl.cpp:155-228; its own tests only ever reach "none" and case 1).  Tensors hold a function of the GLOBAL
coordinates (sample, feature map, pixel), so what must arrive where is known in closed form on every rank."""
import os
import sys

import pytest
import torch

from conftest import ROOT, run_ranks

sys.path.insert(0, os.path.join(ROOT, "examples"))
WORLD, GMB, FM, FS = 4, 16, 16, 4


def _f(mb0, nmb, fm0, nfm):
    mb = torch.arange(mb0, mb0 + nmb).view(-1, 1, 1)
    fm = torch.arange(fm0, fm0 + nfm).view(1, -1, 1)
    s = torch.arange(FS).view(1, 1, -1)
    return (mb * 1000 + fm * 10 + s).float().reshape(-1)


# (case, producer op type, producer distribution, consumer distribution); "colors" = CreateDistributionWithColors
CASES = [
    (1, "CC", ("grid", 2, 2), ("grid", 2, 2)),
    (2, "CC", ("grid", 2, 2), ("colors", lambda r: r % 2, lambda r: r)),     # same data groups {0,2},{1,3}, no model split
    (3, "CC", ("grid", 2, 2), ("grid", 4, 1)),
    (4, "CC", ("grid", 4, 1), ("grid", 2, 2)),
    (5, "ACT", ("grid", 2, 2), ("grid", 4, 1)),
]


@pytest.mark.parametrize("case,ptype,pdist,cdist", CASES, ids=["case%d" % c[0] for c in CASES])
def test_activation_exchange_case(case, ptype, pdist, cdist):
    run_case(case, ptype, pdist, cdist, "host")


@pytest.mark.parametrize("case,ptype,pdist,cdist", CASES, ids=["case%d" % c[0] for c in CASES])
def test_activation_exchange_case_fused(case, ptype, pdist, cdist):
    """Activation.start_comm_fused: the unpacked tensors go in and come out, no user-side pack / unpack loops."""
    run_case(case, ptype, pdist, cdist, "host", fused=True)


def run_case(case, ptype, pdist, cdist, backend, fused=False):
    dev = "cuda" if backend == "cuda" else "cpu"

    def body(r, mlsl):
        from mlsl_test import Net                      # the packing loop every integration writes (examples/mlsl_test.py)
        from mlsl_b200.api import DataType, OperationType
        e = mlsl.env()

        def make(spec):
            return e.create_distribution(spec[1], spec[2]) if spec[0] == "grid" else e.create_distribution_with_colors(spec[1](r), spec[2](r))

        dp = make(pdist)
        dc = dp if (case == 1) else make(cdist)
        sess = e.create_session()
        sess.set_global_minibatch_size(GMB)
        r0 = sess.create_operation_reg_info(getattr(OperationType, ptype))
        r0.add_input(FM, FS, DataType.FLOAT)
        r0.add_output(FM, FS, DataType.FLOAT)
        if ptype == "CC":
            r0.add_parameter_set(FM * FM, 1, DataType.FLOAT)
        r1 = sess.create_operation_reg_info(OperationType.CC)
        r1.add_input(FM, FS, DataType.FLOAT)
        r1.add_output(FM, FS, DataType.FLOAT)
        r1.add_parameter_set(FM * FM, 1, DataType.FLOAT)
        op0 = sess.get_operation(sess.add_operation(r0, dp))
        op1 = sess.get_operation(sess.add_operation(r1, dc))
        op1.set_prev(op0, 0, 0)
        sess.commit()
        oa, ia = op0.get_output(0), op1.get_input(0)
        view = lambda addr, nbytes: mlsl.tensor_from_address(addr, (nbytes // 4,), torch.float32)     # noqa: E731
        p_mb, p_off = op0.get_local_minibatch_size(), op0.get_global_minibatch_offset()
        c_mb, c_off = op1.get_local_minibatch_size(), op1.get_global_minibatch_offset()
        p_fm, p_fo = oa.get_local_fm_count(), oa.get_global_fm_offset()
        c_fm, c_fo = ia.get_local_fm_count(), ia.get_global_fm_offset()
        # ---- forward: producer's output -> consumer's input
        if fused:
            out = mlsl.alloc_tensor(p_mb * p_fm * FS, torch.float32)
            out.copy_(_f(p_off, p_mb, p_fo, p_fm))
            inp = mlsl.alloc_tensor(c_mb * c_fm * FS, torch.float32)
            oa.start_comm_fused(out, inp)
            got = ia.wait_comm()
            assert got == inp.data_ptr()
        else:
            out = _f(p_off, p_mb, p_fo, p_fm).to(dev)
            comm = view(oa.get_comm_buf(), oa.get_comm_buf_size())
            Net.move_blocks(oa, comm, out, False)
            oa.start_comm(comm)
            got = ia.wait_comm()
            inp = torch.zeros(c_mb * c_fm * FS, device=dev)
            # the pointer WaitComm returns lies in the PRODUCER's request buffer; the unpack blocks say how much of it is ours
            Net.move_blocks(ia, view(got, max(ia.get_comm_buf_size(), c_mb * c_fm * FS * 4)), inp, True)
        partial = dp.get_process_count(1) if ptype == "CC" else 1          # an OT_CC output is a partial sum per model rank
        fwd_ok = torch.equal(inp.cpu(), partial * _f(c_off, c_mb, c_fo, c_fm))
        # ---- backward: consumer's input gradient -> producer's output gradient
        bwd_ok = True
        if fused and case != 2:
            din = mlsl.alloc_tensor(c_mb * c_fm * FS, torch.float32)
            din.copy_(_f(c_off, c_mb, c_fo, c_fm))
            dout = mlsl.alloc_tensor(p_mb * p_fm * FS, torch.float32)
            ia.start_comm_fused(din, dout)
            back = oa.wait_comm()
            bwd_ok = back == dout.data_ptr() and torch.equal(dout.cpu(), _f(p_off, p_mb, p_fo, p_fm))
            res = (fwd_ok, bwd_ok, oa.get_pack_block_count(), ia.get_unpack_block_count())
            e.delete_session(sess)
            if dc is not dp:
                e.delete_distribution(dc)
            e.delete_distribution(dp)
            return res
        din = _f(c_off, c_mb, c_fo, c_fm).to(dev)
        addr = ia.get_comm_buf()
        commi = view(addr, ia.get_comm_buf_size()) if addr else din
        if addr:
            Net.move_blocks(ia, commi, din, False)
        ia.start_comm(commi)
        back = oa.wait_comm()
        if case == 2:
            bwd_ok = back is None                                           # nothing travels backward in this pattern
        else:
            dout = torch.zeros(p_mb * p_fm * FS, device=dev)
            Net.move_blocks(oa, view(back, max(oa.get_comm_buf_size(), p_mb * p_fm * FS * 4)), dout, True)
            bwd_ok = torch.equal(dout.cpu(), _f(p_off, p_mb, p_fo, p_fm))
        res = (fwd_ok, bwd_ok, oa.get_pack_block_count(), ia.get_unpack_block_count())
        e.delete_session(sess)
        if dc is not dp:
            e.delete_distribution(dc)
        e.delete_distribution(dp)
        return res

    env = {"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20"} if backend == "cuda" else None
    outs = run_ranks(WORLD, body, backend=backend, env=env)
    for r, (fwd_ok, bwd_ok, npack, nunpack) in enumerate(outs):
        assert fwd_ok, (case, r, "forward")
        assert bwd_ok, (case, r, "backward")
        assert npack >= 1 and nunpack >= 1


def test_operation_with_two_outputs_feeding_two_consumers():
    """A small DAG: op0 has two outputs, wired with SetPrev (to op1) and SetNext (to op2); both edges are model-parallel
    exchanges (case 1) that run concurrently; inputs / outputs that are not wired stay communication-free."""
    def body(r, mlsl):
        from mlsl_test import Net
        from mlsl_b200.api import DataType, OperationType
        e = mlsl.env()
        dist = e.create_distribution(2, 2)
        sess = e.create_session()
        sess.set_global_minibatch_size(GMB)

        def reg(n_out):
            ri = sess.create_operation_reg_info(OperationType.CC)
            ri.add_input(FM, FS, DataType.FLOAT)
            for _ in range(n_out):
                ri.add_output(FM, FS, DataType.FLOAT)
            ri.add_parameter_set(FM * FM, 1, DataType.FLOAT)
            return ri

        op0 = sess.get_operation(sess.add_operation(reg(2), dist))
        op1 = sess.get_operation(sess.add_operation(reg(1), dist))
        op2 = sess.get_operation(sess.add_operation(reg(1), dist))
        op1.set_prev(op0, 0, 0)          # op0.out0 -> op1.in0
        op0.set_next(op2, 1, 0)          # op0.out1 -> op2.in0
        sess.commit()
        view = lambda addr, nbytes: mlsl.tensor_from_address(addr, (nbytes // 4,), torch.float32)     # noqa: E731
        mb, off = op0.get_local_minibatch_size(), op0.get_global_minibatch_offset()
        oks = []
        outs = [op0.get_output(0), op0.get_output(1)]
        for k, oa in enumerate(outs):    # start both transfers before waiting for either
            out = (k + 1) * _f(off, mb, 0, FM)
            comm = view(oa.get_comm_buf(), oa.get_comm_buf_size())
            Net.move_blocks(oa, comm, out, False)
            oa.start_comm(comm)
        for k, cons in enumerate((op1, op2)):
            ia = cons.get_input(0)
            got = ia.wait_comm()
            inp = torch.zeros(mb * ia.get_local_fm_count() * FS)
            Net.move_blocks(ia, view(got, mb * ia.get_local_fm_count() * FS * 4), inp, True)
            want = 2 * (k + 1) * _f(off, mb, ia.get_global_fm_offset(), ia.get_local_fm_count())   # 2 partial sums
            oks.append(torch.equal(inp, want))
        unwired = (op0.get_input(0).get_comm_buf_size(), op1.get_output(0).get_comm_buf_size(), op0.get_input(0).wait_comm())
        e.delete_session(sess)
        e.delete_distribution(dist)
        return oks, unwired

    for oks, unwired in run_ranks(WORLD, body):
        assert oks == [True, True]
        assert unwired == (0, 0, None)
