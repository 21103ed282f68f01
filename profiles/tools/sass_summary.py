"""Per-kernel SASS opcode summary of lib/libb200taco.so: the Blackwell-specific instructions (tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP, TMA tensor copies -> UTMALDG/UTMASTG, cluster / DSMEM traffic,
mbarrier waits) next to instruction count, registers are in the build log (`build.py --verbose`).

    python profiles/tools/sass_summary.py > profiles/r2_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "pytorch-tacotron2-implementation_b200", "lib", "libb200taco.so")
OPS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"), ("UTMALDG/STG", r"\bUTMA(LDG|STG)"),
       ("SYNCS (mbarrier)", r"\bSYNCS"), ("UCGABAR (cluster barrier)", r"\bUCGABAR"), ("ST/ATOM .SHARED cluster (DSMEM)", r"\b(ST|ATOM|RED)S?\.[A-Z0-9.]*CLUSTER|MAPA"),
       ("LDG.STRONG.GPU (flag polls)", r"\bLDG\.E[.0-9A-Z]*STRONG\.GPU"), ("STG.STRONG.GPU", r"\bSTG\.E[.0-9A-Z]*STRONG\.GPU"),
       ("HMMA (legacy tensor path)", r"\bHMMA"), ("FFMA", r"\bFFMA"), ("MUFU", r"\bMUFU"), ("BAR", r"\bBAR\.")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cut = cur.rfind(">(")
            cur = cur[:cut + 1] if cut >= 0 else re.sub(r"\(.*", "", cur)
            cur = cur.replace("void ", "").replace("b200taco::<unnamed>::", "").replace("b200taco::", "").replace("(int)", "").replace("(bool)", "")
            kernels[cur] = collections.Counter()
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            kernels[cur]["instr"] += 1
            for name, pat in OPS:
                if re.search(pat, line):
                    kernels[cur][name] += 1
    print("# SASS opcode summary of lib/libb200taco.so (sm_100a), per kernel\n")
    print("| kernel | instr | " + " | ".join(n for n, _ in OPS) + " |")
    print("|---|---:|" + "---:|" * len(OPS))
    for k, c in sorted(kernels.items(), key=lambda kv: -kv[1]["instr"]):
        if c["instr"] < 40:
            continue
        print(f"| `{k}` | {c['instr']} | " + " | ".join(str(c[n]) if c[n] else "" for n, _ in OPS) + " |")


if __name__ == "__main__":
    sys.exit(main())
