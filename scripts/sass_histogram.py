#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels of libbsx.so (cuobjdump -sass): which pipes a kernel's instructions go to
(IMAD.WIDE / narrow IMAD = fmaheavy, IADD3 / LOP3 / SHF = ALU, DFMA / DADD = FP64), whether bulk copies (UBLKCP) or tensor
instructions are present, and the code size.  usage: python scripts/sass_histogram.py > profiles/rNN_sass_histograms.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = ["ed25519_batch_kernel", "ed25519_keyed_kernel", "ed25519_key_bases_kernel", "ed25519_key_table_kernel", "ed25519_key_assign_kernel", "ed25519_quad_kernel", "subchain_proofs_kernel", "subchain_commit_kernel", "reduce_subchains_kernel",
       "verify_kernel", "gl_gate_eval_kernel", "gl_gate_quotient_kernel", "gl_poseidon_batch_kernel", "ntt_dif_strided_kernel",
       "ntt_dif_contig_kernel", "gl_merkle_leaves_kernel", "gl_merkle_layer_kernel", "sha256_trace_kernel", "pack_bytes_kernel",
       "encode_headers_kernel", "range_inputs_kernel", "prove_subchain_kernel", "data_commitment_kernel"]
PIPE = [("fp64", r"^(DFMA|DADD|DMUL|DSETP)"), ("fmaheavy (wide)", r"^IMAD\.WIDE"), ("fma (narrow IMAD)", r"^IMAD(\.|$)(?!WIDE)"),
        ("alu", r"^(IADD3|LOP3|SHF|VIADD|PRMT|LEA|ISETP|SEL|MOV|IABS|FLO|POPC|BREV|VIMNMX|IMNMX)"),
        ("lsu", r"^(LD|ST|ATOM|RED)[GSLE.]"), ("bulk copy (TMA 1-D)", r"^UBLKCP"), ("tensor", r"^(UTC|HMMA|IMMA|DMMA)"),
        ("conversion", r"^(I2F|F2I|F2F|I2I)"), ("control", r"^(BRA|CALL|RET|EXIT|BAR|BSSY|BSYNC|WARPSYNC|NANOSLEEP|YIELD|NOP|DEPBAR)")]


def main():
    so = os.path.join(ROOT, "blobstreamx_b200", "libbsx.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    for f in re.split(r"\n\s+Function : ", txt)[1:]:
        name = f.split("\n")[0].strip()
        if not any(h in name for h in HOT):
            continue
        ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
        c = collections.Counter(ops)
        total = sum(c.values())
        pipes = collections.Counter()
        for op, k in c.items():
            for pname, rx in PIPE:
                if re.match(rx, op):
                    pipes[pname] += k
                    break
            else:
                pipes["other"] += k
        print(f"== {demangle(name)[:150]}")
        print(f"   {total} instructions ({16 * total / 1024:.1f} KB); by pipe: " + ", ".join(f"{p} {k}" for p, k in pipes.most_common()))
        print("   top: " + ", ".join(f"{op} {k}" for op, k in c.most_common(14)))


if __name__ == "__main__":
    main()
