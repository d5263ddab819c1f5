#!/usr/bin/env python
"""Full SASS listing of one representative instance of every kernel family -> profiles/sass/<kernel>.sass
    python scripts/sass_listings.py        (needs the built library; CPU only)"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mlsl_b200", "lib", "libmlsl_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
# output name -> regex on the demangled kernel name (first match wins)
WANT = [
    ("k_allreduce_f32_sum_u4_nvls", r"k_allreduce<float, mlslb::OpSum, 4, true>"),
    ("k_allreduce_f32_sum_u4_p2p", r"k_allreduce<float, mlslb::OpSum, 4, false>"),
    ("k_allreduce_bf16_sum_u4_nvls", r"k_allreduce<__nv_bfloat16, mlslb::OpSum, 4, true>"),
    ("k_allreduce_ll_f32_sum", r"k_allreduce_ll<float, mlslb::OpSum>"),
    ("k_allreduce_mid_f32_sum_oneshot", r"k_allreduce_mid<float, mlslb::OpSum, false>"),
    ("k_allreduce_mid_f32_sum_twoshot", r"k_allreduce_mid<float, mlslb::OpSum, true>"),
    ("k_reduce_pull_f32_sum_u4_nvls", r"k_reduce_pull<float, mlslb::OpSum, 4, true>"),
    ("k_reduce_pull_f32_sum_u4_p2p", r"k_reduce_pull<float, mlslb::OpSum, 4, false>"),
    ("k_pull_copy", r"k_pull_copy\("),
    ("k_pull_copy_bulk", r"k_pull_copy_bulk\("),
    ("k_barrier", r"k_barrier\("),
    ("k_pack_blocks", r"k_pack_blocks\("),
    ("k_scale_copy_f32", r"k_scale_copy<float>"),
    ("k_allreduce_quant", r"k_allreduce_quant<false>"),
    ("k_allreduce_quant_mx", r"k_allreduce_quant<true>"),
    ("k_fused_update_f32_bf16", r"k_fused_update<float, __nv_bfloat16>"),
    ("k_fused_update_f32_f32", r"k_fused_update<float, float>"),
    ("k_gemm_rs", r"k_gemm_rs[<(]"),
    ("k_gemm_rs2", r"k_gemm_rs2[<(]"),
    ("k_ag_gemm", r"k_ag_gemm[<(]"),
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    mangled = re.findall(r"Function : (\S+)", sass)
    names = subprocess.run(["c++filt"], input="\n".join(mangled), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    bodies = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    os.makedirs(OUT, exist_ok=True)
    for out, pat in WANT:
        for name, body in zip(names, bodies):
            if re.search(pat, name):
                with open(os.path.join(OUT, out + ".sass"), "w") as f:
                    text = "\n".join(re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln) for ln in body.split("\n\t\t.....")[0].splitlines() if ln.strip())
                    f.write("// %s\n// sm_100a, cuobjdump -sass mlsl_b200/lib/libmlsl_b200.so (encodings stripped)\n%s\n" % (name, text))
                ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
                key = sorted({o.split(".")[0] for o in ops if o.startswith(("UTC", "LDTM", "UTMA", "UBLKCP", "LDGMC", "STGMC", "SYNCS", "UCGABAR"))})
                print("%-36s %6d instructions  %s" % (out, len(ops), " ".join(key)))
                break
        else:
            print("%-36s NOT FOUND (%s)" % (out, pat))


if __name__ == "__main__":
    main()
