#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libekl_b200.so (run here, no GPU needed):

    python tools/sass_counts.py > profiles/r02_sass_mnemonics.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "text2img_ekl_b200", "libekl_b200.so")
PAT = collections.OrderedDict([("UTCHMMA", "UTCHMMA"), ("UTMALDG", "UTMALDG"), ("UTMASTG", "UTMASTG"), ("UBLKCP", "UBLKCP"), ("LDTM", "LDTM"),
                               ("UTCBAR", "UTCBAR"), ("SYNCS", "SYNCS"), ("RED", " RED."), ("REDG", " REDG."), ("ATOM", " ATOM")])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            cnt[cur]["instr"] += 1
            for k, p in PAT.items():
                if p in line:
                    cnt[cur][k] += 1
    names = list(cnt)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel of `libekl_b200.so`\n")
    print("`cuobjdump -sass text2img_ekl_b200/libekl_b200.so`, counted per `Function :` block by `tools/sass_counts.py`. "
          "UTCHMMA = `tcgen05.mma` (bf16), UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = `cp.async.bulk` (1-D bulk copy), LDTM = `tcgen05.ld`, UTCBAR = "
          "`tcgen05.commit`, SYNCS = mbarrier operations, RED/REDG = `red.global.add`.\n")
    keys = ["instr"] + list(PAT)
    print("| kernel | " + " | ".join(keys) + " |\n|---|" + "---:|" * len(keys))
    tot = collections.Counter()
    for n, d in zip(names, dem):
        c = cnt[n]
        tot.update(c)
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"^void ", "", d)
        d = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", d)
        print("| `%s` | " % d[:80] + " | ".join(str(c[k]) for k in keys) + " |")
    print("| **total (%d kernels)** | " % len(names) + " | ".join(str(tot[k]) for k in keys) + " |")


if __name__ == "__main__":
    main()
