"""Blackwell-native evidence per kernel of libot_b200.so: counts of the SASS mnemonics behind tcgen05.mma (UTC*MMA), tcgen05.ld/st
(LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), mbarrier (SYNCS), st.async (STAS), packed fp32 (FFMA2) and the legacy paths (HMMA,
IDP.4A = dp4a), plus the first line of each tensor-core / TMA mnemonic.  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "onnx-transformer_b200", "libot_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "STAS", "UCGABAR", "FFMA2", "HMMA", "IDP.4A", "BRA.U.ANY"]
cur, counts, first, sizes = None, collections.defaultdict(collections.Counter), collections.defaultdict(dict), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if not m or cur is None:
        continue
    sizes[cur] += 1
    ins = m.group(1)
    for k in keys:
        if re.search(r"(^|\s|@!?U?P\d\s+)" + re.escape(k), ins):
            counts[cur][k] += 1
            first[cur].setdefault(k, ins.strip())
print("# %s: SASS mnemonic counts per kernel (cuobjdump -sass); kernels without any of them are omitted" % os.path.basename(lib))
for fn in sorted(counts, key=lambda f: -sizes[f]):
    c = counts[fn]
    print("\n%s   [%d SASS instructions]" % (fn, sizes[fn]))
    print("   " + "  ".join("%s=%d" % (k, c[k]) for k in keys if c[k]))
    for k in ("UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "STAS", "FFMA2"):
        if k in first[fn]:
            print("      e.g. " + first[fn][k][:150])
