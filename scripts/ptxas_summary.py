#!/usr/bin/env python
"""Resource usage of every kernel as ptxas reports it (-Xptxas -v, written to build/*.ptxas.log by the Makefile):
registers per thread, spills, static shared memory, stack -> profiles/sass/ptxas_resources.txt
    python scripts/ptxas_summary.py        (after `make`; CPU only)"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rows = []
    for log in sorted(glob.glob(os.path.join(ROOT, "build", "*.ptxas.log"))):
        text = open(log).read()
        # ptxas info    : Compiling entry function '<mangled>' for 'sm_100a'
        # ptxas info    : Function properties for <mangled>
        #     N bytes stack frame, N bytes spill stores, N bytes spill loads
        # ptxas info    : Used N registers, used N barriers, N bytes smem, ...
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'(.*?)(?=Compiling entry function|\Z)", text, flags=re.S):
            name, body = m.group(1), m.group(2)
            st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", body)
            us = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", body)
            if not us:
                continue
            rows.append((os.path.basename(log).split(".")[0], name, int(us.group(1)), int(us.group(3) or 0), int(st.group(1)) if st else 0,
                         int(st.group(2)) if st else 0, int(st.group(3)) if st else 0))
    names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    out = ["ptxas -v resource usage of every kernel (sm_100a, nvcc 12.9; scripts/ptxas_summary.py)",
           "dynamic shared memory (TMA stages of the GEMM kernels, bulk-copy rings) is set at launch and not listed here",
           "", "%-14s %5s %6s %6s %7s %7s  %s" % ("file", "regs", "smem", "stack", "spill_st", "spill_ld", "kernel")]
    for (f, _, regs, smem, stack, ss, sl), nm in sorted(zip(rows, names), key=lambda t: (t[0][0], t[1])):
        out.append("%-14s %5d %6d %6d %7d %7d  %s" % (f, regs, smem, stack, ss, sl, nm[:150]))
    spilled = [nm for (r, nm) in zip(rows, names) if r[5] or r[6]]
    out += ["", "%d kernels, %d with register spills%s" % (len(rows), len(spilled), (": " + "; ".join(s[:80] for s in spilled[:8])) if spilled else "")]
    path = os.path.join(ROOT, "profiles", "sass", "ptxas_resources.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-3:]))
    print("wrote", path)


if __name__ == "__main__":
    main()
