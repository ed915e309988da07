"""Instruction histogram of the hot kernels from the built library's SASS (cuobjdump -sass).

    python profiles/sass_histogram.py [--kernel SUBSTR ...] [--dump DIR] > profiles/rN_sass_histogram.txt

Classes: IMAD.WIDE (the 32x32->64 multiplier op), IMAD (32-bit multiply-add incl. IMAD.MOV/IADD/SHL that ptxas
places on the multiplier unit, listed separately), ALU (IADD3/LOP3/SHF/SEL/ISETP/...), LSU global / shared,
TMA (UTMALDG/UTMASTG/UBLKCP), barriers.  Static counts: a kernel that is fully unrolled (the NTT tile kernels) executes
every instruction once per tile, so static counts / butterflies per thread = instructions per butterfly.
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "fhe_rs_b200", "libfhe_b200.so")

DEFAULT = ["ntt_tma_cols_kernelILi9", "ntt_tma_cols_kernelILi8", "ntt_tma_rows_kernel", "scale_tma_kernel",
           "ksmac_tma_kernel", "tensor_kernel", "ntt_fast_kernelILi9ELb1ELb0ELi11", "ntt_fast_kernelILi6ELb0ELb0ELi10",
           "scale_kernel", "12ksmac_kernel"]


def classify(op: str) -> str:
    base = op.split(".")[0]
    if op.startswith("IMAD.WIDE"):
        return "IMAD.WIDE"
    if op.startswith(("IMAD.MOV", "IMAD.IADD", "IMAD.SHL", "IMAD.X")):
        return "IMAD(move/add on mul unit)"
    if base in ("IMAD", "IMUL"):
        return "IMAD"
    if base in ("UTMALDG", "UTMASTG", "UBLKCP", "UTMACCTL", "UTMACMDFLUSH", "UTMAPF", "UBLKPF", "UBLKRED", "UTMAREDG"):
        return "TMA"
    if base in ("LDG", "STG", "LD", "ST", "LDGSTS", "ATOMG", "REDG"):
        return "LSU global"
    if base in ("LDS", "STS", "LDSM", "STSM", "ATOMS"):
        return "LSU shared"
    if base in ("LDC", "ULDC", "LDCU"):
        return "const load"
    if base in ("BAR", "SYNCS", "ARRIVES", "MEMBAR", "FENCE", "DEPBAR", "WARPSYNC", "ERRBAR", "CCTL", "UCGABAR_ARV",
                "UCGABAR_WAIT", "ACQBULK", "ELECT"):
        return "sync"
    if base in ("BRA", "EXIT", "RET", "CALL", "BSSY", "BSYNC", "NOP", "BRX", "JMP", "BREAK", "YIELD", "NANOSLEEP"):
        return "control"
    if base.startswith("U") and base not in ("UTMALDG",):
        return "uniform ALU"
    return "ALU"


def kernels():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            body.append((m.group(1), line.strip()))
    if name:
        yield name, body


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", action="append")
    ap.add_argument("--dump", help="directory for the full SASS text of each selected kernel")
    a = ap.parse_args()
    want = a.kernel or DEFAULT
    demangle = {}
    for name, body in kernels():
        if not any(w in name for w in want):
            continue
        try:
            pretty = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        except FileNotFoundError:
            pretty = name
        demangle[name] = pretty
        h = collections.Counter(classify(op) for op, _ in body)
        ops = collections.Counter(op.split(".")[0] if not op.startswith("IMAD") else ".".join(op.split(".")[:2])
                                  for op, _ in body)
        total = sum(h.values())
        print("== %s" % pretty)
        print("   total %d instructions" % total)
        for k, v in sorted(h.items(), key=lambda kv: -kv[1]):
            print("   %-28s %6d  %5.1f%%" % (k, v, 100.0 * v / total))
        print("   top opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common(14)))
        if a.dump:
            os.makedirs(a.dump, exist_ok=True)
            short = re.sub(r"[^A-Za-z0-9]+", "_", pretty)[:80]
            with open(os.path.join(a.dump, short + ".sass"), "w") as f:
                f.write("\n".join(l for _, l in body) + "\n")
        print()


if __name__ == "__main__":
    main()
