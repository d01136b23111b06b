"""Writes profiles/<tag>_sass_tcgen05_tma.md: per-kernel counts of the tcgen05 / TMEM / TMA mnemonics in the built
libswr_b200.so (cuobjdump -sass) plus one excerpt per mnemonic, so the evidence is in the repository.
usage: python tools/sass_evidence.py <tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "scenario-wise-rec_b200", "scenario_wise_rec_b200", "libswr_b200.so")
PAT = re.compile(r"\b(UTMALDG|UTMASTG|UBLKCP|UTMACCTL|UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|SYNCS|USETMAXREG|ELECT|RED)\b[.\w]*")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    fn, counts, first = None, collections.defaultdict(collections.Counter), {}
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = PAT.search(line)
        if m and fn:
            counts[fn][m.group(1)] += 1
            first.setdefault((fn, m.group(1)), re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line.strip()))
    cols = ["UTMALDG", "UTMACCTL", "UTCHMMA", "UTCBAR", "UTCATOMSWS", "STTM", "LDTM", "SYNCS", "USETMAXREG", "ELECT", "RED"]
    out = [f"# SASS evidence `{tag}`: tcgen05 / TMEM / TMA instructions in `libswr_b200.so`", "",
           "`cuobjdump -sass scenario-wise-rec_b200/scenario_wise_rec_b200/libswr_b200.so`, counted per kernel by "
           "`tools/sass_evidence.py` (UTMALDG = `cp.async.bulk.tensor` loads, UTMACCTL.PF = `prefetch.tensormap`, UTCHMMA = "
           "`tcgen05.mma.kind::tf32`, UTCBAR = `tcgen05.commit`, UTCATOMSWS = `tcgen05.alloc/dealloc`, STTM / LDTM = "
           "`tcgen05.st / ld`, SYNCS = mbarrier ops, USETMAXREG = `setmaxnreg`, RED = `red.global.add`).", "",
           "| kernel | " + " | ".join(cols) + " |", "|---|" + "---|" * len(cols)]
    for f in sorted(counts):
        c = counts[f]
        if not any(c[k] for k in ("UTCHMMA", "UTMALDG", "LDTM", "STTM")):
            continue
        out.append(f"| `{f}` | " + " | ".join(str(c[k]) for k in cols) + " |")
    out += ["", "## First occurrence of each mnemonic per tcgen05 kernel", "", "```"]
    for (f, k), line in sorted(first.items()):
        if any(counts[f][x] for x in ("UTCHMMA", "UTMALDG")) and k in cols[:9]:
            out.append(f"{f[:40]:40s} {line}")
    out.append("```")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_tcgen05_tma.md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
