"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), LDGSTS (cp.async), plus the legacy
HMMA for contrast.  Runs on the CPU-only box: cuobjdump -sass on the built library.
usage: python tools/sass_opcodes.py [lib.so] > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "panoswintransformerobjectdetection_b200", "libpanoswin_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDGSTS", "HMMA", "MUFU"]
rows = []
for blk, name in zip(sass.split("Function : ")[1:], names):
    cnt = collections.Counter()
    total = 0
    for line in blk.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            total += 1
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    cnt[w] += 1
    short = re.sub(r"\(.*", "", name).replace("void psw::", "psw::")
    targs = re.search(r"<(.*)>", name)
    rows.append((short if not targs else short, total, cnt))
print(f"# SASS opcode summary of {os.path.basename(lib)} (cuobjdump -sass, sm_100a); columns = static instruction counts")
print("| kernel | SASS instr | " + " | ".join(WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
for name, total, cnt in sorted(rows, key=lambda r: r[0]):
    print(f"| `{name[:110]}` | {total} | " + " | ".join(str(cnt.get(w, 0)) if cnt.get(w, 0) else "" for w in WATCH) + " |")
