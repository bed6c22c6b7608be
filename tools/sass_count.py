#!/usr/bin/env python
"""Static SASS instruction mix per kernel (no GPU needed): python tools/sass_count.py 12 [13 ...]"""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
KIND = {0: "C2C_FWD", 1: "C2C_BWD", 2: "R2C", 3: "C2R"}


def main():
    for logm in sys.argv[1:]:
        obj = os.path.join(ROOT, "chowdsp_fft_b200", "csrc", "build", f"fft_inst_{logm}.o")
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, counts = None, {}
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and cur:
                counts[cur][m.group(1)] += 1
        for fn, c in sorted(counts.items()):
            m = re.search(r"fft_kernelILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)E", fn)
            if not m:
                continue
            name = f"M=2^{m.group(1)} {KIND[int(m.group(3))]:8s} logW={m.group(4)}"
            total = sum(c.values())
            fp = c["FADD2"] + c["FMUL2"] + c["FFMA2"] + c["FADD"] + c["FMUL"] + c["FFMA"]
            mem = {k: c[k] for k in ("LDG", "STG", "LDS", "STS", "BAR")}
            ints = total - fp - sum(mem.values())
            print(f"{name}: total={total:5d} fp={fp:4d} int/other={ints:4d} " + " ".join(f"{k}={v}" for k, v in mem.items()))


if __name__ == "__main__":
    main()
