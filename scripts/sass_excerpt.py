"""SASS evidence for profiles/: opcode counts of the kernels that matter and an excerpt of the table kernel's inner loop.

    python scripts/sass_excerpt.py > profiles/r02_sass_excerpt.txt        (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gnn_decode_b200", "lib", "libgnn_decode_b200.so")
KERNELS = [("lean_tables_kernel", "double-precision table builder"),
           ("lean_prep_kernel", "input packing + prior discovery + hash check + max|mlp2|"),
           ("lean_decode_kernelILb0E", "check-owner table kernel (csrc/gd_lean.cu): the decoder_v2_4 headline path"),
           ("lean_decode_kernelILb1E", "the same with the training stash"),
           ("lean_bwd_kernel", "training backward on the tables: adjoint bins on 32-bit shared atomics with carry (ATOMS.ADD, no CAS loop)"),
           ("lean_contract_kernel", "contraction of the adjoint bins in double"),
           ("decode_streamed_tma_kernelILi1E", "streamed global-memory decoder: bulk async copies (UBLKCP) + mbarrier (SYNCS)"),
           ("decode_kernelILi2ELi1024ELi2ELi2E", "edge-owner resident kernel, decoder_v2_4 (direct evaluation: MUFU.EX2 / MUFU.LG2, packed FFMA2)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    print("# SASS evidence (cuobjdump -sass gnn_decode_b200/lib/libgnn_decode_b200.so, sm_100a), round 2, final kernels\n")
    for key, what in KERNELS:
        for f in funcs:
            name = f.split("\n", 1)[0].strip()
            if key not in name:
                continue
            ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", f)
            cnt = collections.Counter(ops)
            print("## %s\n   %s\n   %d SASS instructions; opcode counts of interest:" % (name, what, len(ops)))
            print("   " + ", ".join("%s x%d" % kv for kv in cnt.most_common(22)))
            extra = [k for k in cnt if k.startswith(("MUFU", "UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED", "BAR", "FFMA2", "F2I", "MATCH", "REDUX"))]
            print("   pipes / special: " + ", ".join("%s x%d" % (k, cnt[k]) for k in sorted(extra)) + "\n")
            if key == "lean_decode_kernelILb0E":
                lines = f.split("\n")
                bars = [i for i, l in enumerate(lines) if "BAR.SYNC" in l]
                if len(bars) > 2:
                    i0 = bars[1]
                    print("   excerpt (after the named barrier that ends an iteration: metadata broadcast loads, sibling message loads, the variable-phase\n"
                          "   table look-up = FFMA, FADD (magic number), IMAD (piece * stride), LDS.128, 3 FFMA; the check-phase look-up likewise; no MUFU, no F2I):")
                    for l in lines[i0:i0 + 90]:
                        if "/*" in l and not re.match(r"\s*/\* 0x", l):
                            print("   " + l.rstrip()[:130])
                    print()
            break


if __name__ == "__main__":
    main()
