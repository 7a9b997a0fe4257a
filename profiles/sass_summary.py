"""Counts the Blackwell-specific SASS mnemonics per kernel of the shipped library (evidence that the hot path is
hand-written tcgen05 / TMEM / TMA code, not a library call):

    python profiles/sass_summary.py [out.md]        # needs cuobjdump (CUDA toolkit); no GPU
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stabletts_b200", "libstabletts_b200.so")
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UTCCP", "SYNCS",
             "MUFU.EX2", "HMMA", "FFMA", "STG", "LDG", "STS", "LDS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = name.replace("(anonymous namespace)::", "").replace("st::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    cur[mn] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
    cols = [m for m in MNEMONICS if any(c[m] for c in per.values())]
    lines = ["| kernel | SASS instr | " + " | ".join(cols) + " |", "|---|---:|" + "---:|" * len(cols)]
    for name, c in per.items():
        if not any(c[m] for m in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG")) and c["_total"] < 400:
            continue
        lines.append(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[m]) if c[m] else "" for m in cols) + " |")
    text = ("# SASS summary of libstabletts_b200.so (sm_100a)\n\n`cuobjdump -sass` of the in-tree library, mnemonic counts per kernel "
            "(UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; LDTM / STTM = tcgen05.ld / .st; UTMALDG / UTMASTG = TMA load / store; "
            "UTCBAR = tcgen05.commit).  Small glue kernels without any of these are omitted.\n\n" + "\n".join(lines) + "\n")
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
