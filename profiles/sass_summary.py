"""Per-kernel SASS opcode evidence of libcrog_b200.so (run anywhere: cuobjdump needs no GPU).

    python profiles/sass_summary.py > profiles/sass_summary.txt

For every kernel in the shipped library: counts of the Blackwell-native mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR) plus the legacy tensor path (HMMA) that
must NOT appear, and the packed-math / special-function mnemonics the kernel notes in DESIGN.md refer to.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "crog_b200", "lib", "libcrog_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "FFMA2", "FADD2",
        "MUFU.EX2", "FMNMX3", "SYNCS", "LDS", "STS", "LDG", "STG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?PT?\d*\s+)?([A-Z][A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for k in KEYS:
            if k == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    kernels[cur][k] += 1
            elif k in ("LDS", "STS", "LDG", "STG"):
                if op.split(".")[0] == k:
                    kernels[cur][k] += 1
            elif op.startswith(k):
                kernels[cur][k] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    print(f"# {os.path.relpath(SO, ROOT)}: {len(kernels)} kernels; columns: instructions | " + " ".join(KEYS))
    for (name, c), dn in zip(kernels.items(), demangle):
        short = dn.replace("(anonymous namespace)::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", short)
        print(f"{short[:110]:110s} {c['_total']:6d} | " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
        tot.update(c)
    print("# TOTAL " + " ".join(f"{k}={tot[k]}" for k in KEYS))
    assert tot["HMMA"] == 0, "legacy mma.sync path present"


if __name__ == "__main__":
    main()
