#!/usr/bin/env python
"""Static instruction mix of the kernels of a cubin / .so (cuobjdump -sass), grouped FP64 / other: on B200 an FP64
instruction holds a scheduler's issue port for two cycles and every other instruction for about one
(tools/micro/fp64_issue.cu), so an FP64-bound kernel's time goes as 2 x FP64 + other.

    python tools/sass_mix.py nls_b200/libnls_b200.so rk4_1d_resident [more name fragments ...]
"""
import collections
import re
import subprocess
import sys


def main(path, frags):
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True).stdout
    name, mix = None, None
    kernels = []
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            mix = collections.Counter()
            kernels.append((name, mix))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and mix is not None:
            mix[m.group(2)] += 1
    for name, mix in kernels:
        if frags and not all(f in name for f in frags):
            continue
        dp = sum(c for op, c in mix.items() if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        total = sum(mix.values())
        print("%s\n   total %d  fp64 %d  other %d  | 2*fp64+other = %d" % (name[:150], total, dp, total - dp, 2 * dp + total - dp))
        print("   " + ", ".join("%s %d" % (op, c) for op, c in mix.most_common(16)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
