#!/usr/bin/env python
"""Python twin of the functional test (csrc/tests/mlsl_functional_test.cpp) on the object model of `mlsl_b200.api`
(the reference ships the same scenario three times: C++, C and Python - tests/examples/mlsl_test/).

Two OT_CC layers (ifm 128 -> ofm 256 -> 256, 12x12 feature maps, 3x3 kernels, global minibatch 16), 2 epochs of 3
minibatches, index-valued tensors so that every exchanged element has a closed-form expected value:
    forward  : layer 0 emits out[i] = i; layer 1 must receive M * (global index)      (M = model group size)
    backward : layer 1 emits dIn[i] = global index; layer 0 must receive dOut[i] = i
    gradients: dW[i] = i must come back as D * (ownedOffset + i)                       (D = data group size)
    increment: after the all-gather every rank holds W[i] = i again

    bin/mlslrun -n 4 python examples/mlsl_test.py <num_groups> [dist_update] [user_buf] [use_test]
    python examples/mlsl_test.py 2 1 --inproc 4           # four virtual ranks inside one process
num_groups = model parts: 1 = data parallel, world = model parallel, in between = hybrid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsl_b200 as mlsl  # noqa: E402
from mlsl_b200.api import DataType, GroupType, OperationType  # noqa: E402

SHAPES = [(128, 256, 12, 3), (256, 256, 12, 3)]       # ifm, ofm, map width, kernel width
GLOBAL_MB, EPOCHS, MB_PER_EPOCH = 16, 2, 3


class Layer:
    pass


class Net:
    def __init__(self, model_parts, dist_update, user_buf, use_test):
        self.model_parts, self.dist_update, self.user_buf, self.use_test = model_parts, dist_update, user_buf, use_test
        self.passed = self.failed = 0
        self.loopback = False

    def check(self, ok, what, layer):
        if ok:
            self.passed += 1
        else:
            self.failed += 1
        print("[%d] %s_%d: %s" % (self.rank, what, layer, "PASSED" if ok else "FAILED"), flush=True)

    def alloc(self, n):
        n = max(int(n), 1)
        if self.user_buf:        # memory the library has never seen: it is staged through the symmetric heap
            return torch.zeros(n, dtype=torch.float32, device="cuda" if mlsl.is_device() else "cpu")
        return mlsl.alloc_tensor(n, torch.float32)

    def view(self, addr, n):
        return mlsl.tensor_from_address(addr, (int(n),), torch.float32)

    @staticmethod
    def move_blocks(act, comm, local, unpack):
        """Copy between the packed communication buffer and the [minibatch, feature map, pixel] tensor following the
        CommBlockInfo list (the loop nest every framework integration of MLSL has to write)."""
        lfm = act.get_local_fm_count()
        count = act.get_unpack_block_count() if unpack else act.get_pack_block_count()
        for b in range(count):
            bi = act.get_unpack_block(b) if unpack else act.get_pack_block(b)
            fs, fc, fo = bi.get_fm_size(), bi.get_fm_count(), bi.get_fm_offset()
            mo, mc, bo = bi.get_mb_offset(), bi.get_mb_count(), bi.get_buf_offset()
            c = comm[bo:bo + mc * fc * fs].view(mc, fc, fs)
            loc = local.view(-1, lfm, fs)[mo:mo + mc, fo:fo + fc, :]
            if unpack:
                loc.copy_(c)
            else:
                c.copy_(loc)

    def global_index(self, L):
        """value layer 1 expects / emits at every position of its input tensor"""
        ia = L.op.get_input(0)
        lmb, lfm, fs, off = L.op.get_local_minibatch_size(), ia.get_local_fm_count(), ia.get_fm_size(), ia.get_global_fm_offset()
        M = L.op.get_distribution().get_process_count(GroupType.MODEL)
        mb = torch.arange(lmb).view(-1, 1, 1)
        fm = torch.arange(lfm).view(1, -1, 1)
        s = torch.arange(fs).view(1, 1, -1)
        return (mb * lfm * fs * M + (off + fm) * fs + s).reshape(-1).float(), M

    def forward(self, L):
        ia, oa, ps = L.op.get_input(0), L.op.get_output(0), L.op.get_parameter_set(0)
        got = ia.wait_comm()
        if got:
            self.move_blocks(ia, self.view(got, ia.get_comm_buf_size() // 4), L.inp, True)
        ps.wait_increment_comm()
        if L.idx == 0:
            L.out.copy_(torch.arange(L.out.numel(), dtype=torch.float32))
        else:
            want, M = self.global_index(L)
            self.check(torch.allclose(L.inp.cpu(), M * want, atol=1e-4), "forward_input", L.idx)
        self.check(torch.allclose(L.w.cpu(), torch.arange(L.w.numel(), dtype=torch.float32), atol=1e-4), "forward_param", L.idx)
        addr = oa.get_comm_buf()
        if addr:
            comm = self.view(addr, oa.get_comm_buf_size() // 4)
            self.move_blocks(oa, comm, L.out, False)
            oa.start_comm(comm)
        else:
            oa.start_comm(L.out)
        L.got_out_grad = False

    def fetch_out_grad(self, L):
        if L.got_out_grad:
            return
        oa = L.op.get_output(0)
        got = oa.wait_comm()
        if got:
            self.move_blocks(oa, self.view(got, oa.get_comm_buf_size() // 4), L.out_grad, True)
        L.got_out_grad = True

    def backward_data(self, L):
        self.fetch_out_grad(L)
        ia, oa = L.op.get_input(0), L.op.get_output(0)
        if L.idx == 0:
            if oa.get_unpack_block_count() > 0:
                want = torch.arange(L.out_grad.numel(), dtype=torch.float32)
                self.check(torch.allclose(L.out_grad.cpu(), want, atol=1e-4), "backward_outgrad", L.idx)
        else:
            L.in_grad.copy_(self.global_index(L)[0])
        addr = ia.get_comm_buf()
        if addr:
            comm = self.view(addr, ia.get_comm_buf_size() // 4)
            self.move_blocks(ia, comm, L.in_grad, False)
            ia.start_comm(comm)
        else:
            ia.start_comm(L.in_grad)

    def backward_weights(self, L):
        self.fetch_out_grad(L)
        L.dw.copy_(torch.arange(L.dw.numel(), dtype=torch.float32))
        L.op.get_parameter_set(0).start_gradient_comm(L.dw)

    def update(self, L):
        ps = L.op.get_parameter_set(0)
        if self.use_test:
            done, addr = False, None
            while not done:
                addr, done = ps.test_gradient_comm()
        else:
            addr = ps.wait_gradient_comm()
        D = L.op.get_distribution().get_process_count(GroupType.DATA)
        off = ps.get_owned_kernel_offset() * ps.get_kernel_size()
        own = ps.get_owned_kernel_count() * ps.get_kernel_size()
        g = self.view(addr, own) if addr else L.dw[:own]
        want = D * (off + torch.arange(own, dtype=torch.float32))
        self.check(torch.allclose(g.cpu(), want, atol=1e-4), "update_grad", L.idx)
        L.w[off:off + own].copy_(off + torch.arange(own, dtype=torch.float32))
        ps.start_increment_comm(L.w)

    def run(self):
        # loop-back ranks (several ranks of ONE process on one GPU): Wait blocks the host until the collective is done,
        # so the `.cpu()` checks below never sit in a pageable copy behind a kernel that still spins for a peer
        env = mlsl.init(wait_mode="host" if self.loopback else None)
        self.rank, world = env.get_process_idx(), env.get_process_count()
        M = min(max(self.model_parts, 1), world)
        if world % M:
            print("world size %d not divisible by num_groups %d" % (world, M))
            mlsl.finalize()
            return 2
        sess = env.create_session()
        sess.set_global_minibatch_size(GLOBAL_MB)
        dist = env.create_distribution(world // M, M)
        if self.rank == 0:
            print("world %d: data parts %d, model parts %d, dist_update %d user_buf %d use_test %d (%s)" % (
                world, world // M, M, self.dist_update, self.user_buf, self.use_test, env.describe_backend()), flush=True)
        layers = []
        for l, (ifm, ofm, w, k) in enumerate(SHAPES):
            ri = sess.create_operation_reg_info(OperationType.CC)
            ri.set_name("layer_%d" % l)
            ri.add_input(ifm, w * w, DataType.FLOAT)
            ri.add_output(ofm, w * w, DataType.FLOAT)
            ri.add_parameter_set(ifm * ofm, k * k, DataType.FLOAT, bool(self.dist_update))
            L = Layer()
            L.idx, L.op = l, sess.get_operation(sess.add_operation(ri, dist))
            sess.delete_operation_reg_info(ri)
            if l:
                L.op.set_prev(layers[-1].op, 0, 0)
            layers.append(L)
        sess.commit()
        for L in layers:
            ia, oa, ps = L.op.get_input(0), L.op.get_output(0), L.op.get_parameter_set(0)
            lmb = L.op.get_local_minibatch_size()
            ni, no = ia.get_local_fm_count() * lmb * ia.get_fm_size(), oa.get_local_fm_count() * lmb * oa.get_fm_size()
            np_ = ps.get_local_kernel_count() * ps.get_kernel_size()
            L.inp, L.in_grad, L.out, L.out_grad = self.alloc(ni), self.alloc(ni), self.alloc(no), self.alloc(no)
            L.w, L.dw = self.alloc(np_), self.alloc(np_)
            L.w.copy_(torch.arange(np_, dtype=torch.float32))
        # no exchange between two layers (WaitComm returns None): the consumer reads the producer's tensor directly
        for a, b in zip(layers[:-1], layers[1:]):
            if b.op.get_input(0).get_comm_buf_size() == 0 and a.op.get_output(0).get_comm_buf_size() == 0:
                b.inp, a.out_grad = a.out, b.in_grad
        stats = sess.get_stats()
        stats.start()
        for _ in range(EPOCHS * MB_PER_EPOCH):
            for L in layers:
                self.forward(L)
            for L in reversed(layers):
                self.backward_data(L)       # dX first: its transfer overlaps the dW computation
                self.backward_weights(L)
            for L in layers:
                self.update(L)
        for L in layers:
            L.op.get_parameter_set(0).wait_increment_comm()
            L.op.get_input(0).wait_comm()
        stats.stop()
        if stats.is_enabled():
            stats.print()
        env.delete_session(sess)
        env.delete_distribution(dist)
        del layers
        mlsl.finalize()
        print("[%d] summary: %d PASSED, %d FAILED" % (self.rank, self.passed, self.failed), flush=True)
        return 1 if self.failed else 0


def main():
    args, inproc = [], 0
    it = iter(sys.argv[1:])
    for a in it:
        if a == "--inproc":
            inproc = int(next(it))
        else:
            args.append(int(a))
    if not args:
        print(__doc__)
        return 2
    cfg = (args + [0, 0, 0])[:4]
    if inproc <= 0:
        if torch.cuda.is_available() and os.environ.get("MLSL_BACKEND") == "cuda":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))      # torchrun: one rank per GPU
            torch.cuda.set_stream(torch.cuda.Stream())
        return Net(*cfg).run()

    def body(r):
        mlsl.bind_thread_state()
        if torch.cuda.is_available() and os.environ.get("MLSL_BACKEND") == "cuda":
            # loop-back ranks share the GPU: each needs its own stream (kernels of different ranks wait for each other)
            torch.cuda.set_device(0)
            torch.cuda.set_stream(torch.cuda.Stream())
            net = Net(*cfg)
            net.loopback = True
            return net.run()
        return Net(*cfg).run()

    with mlsl.InprocWorld(inproc) as world:
        rcs = world.run(body)
    bad = any(rcs)
    print("Run FAILED." if bad else "Run PASSED.")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
