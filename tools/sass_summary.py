"""SASS opcode summary of every kernel in libbsr_b200.so (run here, no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
Counts the mnemonics that prove what a kernel is built from (B200_PROFILING.md): UBLKCP = cp.async.bulk
(TMA engine), SYNCS = mbarrier, UTCHMMA[.2CTA] = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, ATOMG/REDG
= global atomics, plus the FP32 and shared-memory instruction mix of the hot loops."""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(HERE, "better_search_rag_rust_b200", "libbsr_b200.so")
KEYS = ["UBLKCP", "SYNCS", "UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOM", "ATOMG", "REDG", "RED.",
        "LDS.128", "LDS", "STS", "LDG", "STG", "ST.E", "LD.E", "FMUL", "FADD", "FFMA", "FMNMX", "SHFL", "BAR.SYNC",
        "MEMBAR", "LDC", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            kernels[name]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[name][k] += 1
                    break
    print(f"# SASS opcode summary of {os.path.relpath(LIB, HERE)} (cuobjdump -sass, sm_100a)")
    for name, c in kernels.items():
        parts = [f"{k}={c[k]}" for k in KEYS if c[k]]
        print(f"{name}\n    instructions={c['_total']}  " + "  ".join(parts))


if __name__ == "__main__":
    main()
