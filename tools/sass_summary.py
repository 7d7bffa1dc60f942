"""Opcode counts per kernel of libhf6d.so (cuobjdump -sass): which kernels use the Blackwell tensor / TMA paths, where
the atomics and warp collectives are.  Runs without a GPU.

  python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "object_detector_6d_b200", "libhf6d.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "RED", "ATOMG", "ATOM", "ATOMS", "MATCH", "REDUX",
         "VOTE", "SHFL", "LDS", "STS", "LDG", "STG", "DADD", "DMUL", "DFMA", "F2F", "I2F", "F2I", "MUFU", "FFMA", "FMUL", "FADD", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1).split(".")[0]] += 1
            kernels[cur]["*"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (static instruction counts, cuobjdump -sass, sm_100a)")
    print("# columns: total | " + " ".join(WATCH))
    for (name, cnt), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\((?!anonymous).*", "", dm).replace("(anonymous namespace)::", "").replace("hf6d::", "").replace("void ", "")
        cols = " ".join(f"{op}={cnt[op]}" for op in WATCH if cnt[op])
        print(f"{short[:70]:70s} {cnt['*']:6d} | {cols}")


if __name__ == "__main__":
    sys.exit(main())
