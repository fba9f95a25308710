#!/usr/bin/env python
"""cuobjdump -sass of ecog2txt_b200/libe2t.so -> profiles/sass/: per kernel, the count of every Blackwell-specific mnemonic
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier, UBLKCP = bulk
copy) plus the first instruction lines that carry them (with the source line when -lineinfo is present).  Evidence that the
hot kernels are tcgen05 / TMEM / TMA code, not recompiled mma.sync (SURVEY.md section 8d asks for a committed listing).
usage: python tools/sass_excerpts.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ecog2txt_b200", "libe2t.so")
OUT = os.path.join(ROOT, "profiles", "sass")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOM",
             "SYNCS", "UBLKCP", "UTCCP", "HMMA", "MUFU", "LDG", "STG", "ATOMG", "RED", "MEMBAR", "ERRBAR", "CCTL", "BAR.SYNC", "ELECT")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    os.makedirs(OUT, exist_ok=True)
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        if cur is not None and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
            kernels[cur].append(line.rstrip())
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    summary = []
    for (mangled, lines), name in zip(kernels.items(), demangle):
        counts = collections.Counter()
        first = {}
        for ln in lines:
            body = ln.split("*/", 1)[-1]
            for mn in MNEMONICS:
                if re.search(r"\b" + re.escape(mn), body):
                    counts[mn] += 1
                    first.setdefault(mn, []).append(ln.strip())
        short = re.sub(r"\(.*", "", name)
        if not any(counts[m] for m in ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "UTMALDG", "UBLKCP")):
            continue
        summary.append((short, len(lines), counts))
        fn = os.path.join(OUT, f"{tag}_{re.sub(r'[^A-Za-z0-9_]+', '_', short)[:80]}.txt")
        with open(fn, "w") as f:
            f.write(f"# {name}\n# mangled: {mangled}\n# {len(lines)} SASS instructions (sm_100a, cuobjdump -sass ecog2txt_b200/libe2t.so)\n")
            f.write("# mnemonic counts: " + ", ".join(f"{k}={v}" for k, v in sorted(counts.items())) + "\n\n")
            for mn in ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "UBLKCP"):
                for ln in first.get(mn, [])[:6]:
                    f.write(ln + "\n")
    with open(os.path.join(OUT, f"{tag}_SUMMARY.txt"), "w") as f:
        f.write("# kernels of libe2t.so that contain tcgen05 / TMEM / TMA instructions: SASS mnemonic counts per kernel\n")
        f.write("# (UTCHMMA = tcgen05.mma kind::tf32/f16, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit)\n")
        for short, n, c in summary:
            f.write(f"{short[:100]:100s} n_instr={n:6d}  " + " ".join(f"{k}={c[k]}" for k in MNEMONICS if c[k]) + "\n")
    print(f"{len(summary)} tcgen05/TMA kernels -> {OUT}")


if __name__ == "__main__":
    main()
