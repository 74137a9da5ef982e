"""SASS opcode histogram of the built library (runs anywhere cuobjdump is installed; no GPU needed):
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "asy-vrnet_b200", "csrc", "libvrcoc.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    per, cur, i = collections.OrderedDict(), None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[i].replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")[:70]
            i += 1
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    per[cur][o] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("# SASS opcode histogram of asy-vrnet_b200/csrc/libvrcoc.so (sm_100a), cuobjdump -sass: tcgen05 = UTCHMMA (MMA), LDTM/STTM (TMEM load/store), "
          "UTCBAR (commit);")
    print("# TMA = UTMALDG / UTMASTG; mbarrier = SYNCS; HMMA = mma.sync (the patch-embed kernel only); packed fp32 = FFMA2.  Whole library: "
          + "  ".join(f"{o}={tot[o]}" for o in OPS))
    print("# kernels that contain tensor-core / TMA instructions:")
    print(f"{'kernel':70s} " + " ".join(f"{o:>8s}" for o in OPS))
    rows = [(k, c) for k, c in per.items() if c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["HMMA"]]
    for k, c in sorted(rows, key=lambda kc: -(kc[1]["UTCHMMA"] * 1000 + kc[1]["HMMA"])):
        print(f"{k:70s} " + " ".join(f"{c[o]:8d}" for o in OPS))


if __name__ == "__main__":
    main()
