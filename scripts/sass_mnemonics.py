"""Regenerate profiles/<tag>_sass_mnemonics.txt from the built library (no GPU needed): for every kernel of libmerv_fusion.so, the count of
the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA ->
UTMALDG/UTMASTG/UBLKCP; multimem.st.relaxed.sys.v4 assembles to STG.E.128.STRONG.SYS — the multicast is a property of the ADDRESS — and
griddepcontrol.launch_dependents / .wait to PREEXIT / ACQBULK; legacy mma.sync would show as HMMA).

    python scripts/sass_mnemonics.py [tag]        # default tag r2
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "merv_b200", "libmerv_fusion.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.[A-Z0-9_]+)*|LDTM(?:\.[A-Za-z0-9_]+)*|STTM(?:\.[A-Za-z0-9_]+)*|UTMALDG(?:\.[A-Z0-9_]+)*|UTMASTG(?:\.[A-Z0-9_]+)*|"
                 r"UTMAPF(?:\.[A-Z0-9_]+)*|UBLKCP(?:\.[A-Z0-9_.]+)*|HMMA(?:\.[A-Z0-9_]+)*|SYNCS(?:\.[A-Z0-9_]+)*|UTCBAR(?:\.[A-Z0-9_]+)*|"
                 r"FFMA2|FMUL2|FADD2|MUFU\.[A-Z0-9]+|ACQBULK|STG\.E\.128\.STRONG\.SYS|UCGABAR[A-Z0-9_.]*|ACQ?GRID[A-Z0-9_.]*|PREEXIT|DEPBAR[A-Z0-9_.]*)")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
    out = [f"# SASS mnemonics per kernel of merv_b200/libmerv_fusion.so (cuobjdump -sass; regenerate with scripts/sass_mnemonics.py {tag})"]
    kernel, counts = None, None
    kernels = []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kernel, counts = demangle(m.group(1)), collections.Counter()
            kernels.append((kernel, counts))
            continue
        if counts is not None:
            for hit in PAT.findall(line):
                counts[hit] += 1
    for name, c in sorted(kernels):
        depth, cut = 0, len(name)
        for i, ch in enumerate(name):  # the parameter list starts at the first '(' outside the template brackets
            depth += (ch == "<") - (ch == ">")
            if ch == "(" and depth == 0:
                cut = i
                break
        short = name[:cut].replace("void merv::", "").replace("merv::", "").replace("(anonymous namespace)::", "")
        key = ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if not k.startswith(("SYNCS", "MUFU", "FFMA2", "FMUL2", "FADD2", "DEPBAR")))
        extra = ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if k.startswith(("FFMA2", "FMUL2", "FADD2")))
        out.append(f"{short}\n    {key or '(no tensor-core / TMA instructions: SIMT kernel)'}" + (f"\n    packed fp32: {extra}" if extra else ""))
    dst = os.path.join(REPO, "profiles", f"{tag}_sass_mnemonics.txt")
    open(dst, "w").write("\n".join(out) + "\n")
    print(f"{dst}: {len(kernels)} kernels")


if __name__ == "__main__":
    main()
