#!/bin/bash
# Blackwell evidence from the built library (runs without a GPU): per-kernel counts of the SASS mnemonics that prove
# tcgen05 (UTC*MMA), TMEM loads (LDTM), TMA loads / stores (UTMALDG / UTMASTG), and the absence of the legacy tensor paths.
# usage: bash scripts/sass_evidence.sh > profiles/sass_r02.txt
LIB=${1:-mask-yolo_b200/lib/libmyolo_sm100.so}
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | grep release | sed 's/.*release //'))"
cuobjdump -sass "$LIB" | python3 -c '
import re, sys, collections
pat = {"UTCHMMA": r"UTCHMMA", "UTCHMMA.2CTA": r"UTCHMMA\.2CTA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"UTMALDG",
       "UTMASTG": r"UTMASTG", "UTCBAR": r"UTCBAR", "HMMA (legacy mma.sync)": r"\bHMMA", "HGMMA (Hopper wgmma)": r"HGMMA|QGMMA|IGMMA"}
cur, counts, arch = None, collections.OrderedDict(), set()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m: arch.add(m.group(1))
    if cur:
        for k, p in pat.items():
            if re.search(p, line): counts[cur][k] += 1
tot = collections.Counter()
print("# architectures in the fatbin:", ", ".join(sorted(arch)))
print("# %d kernels; only kernels with at least one tensor / TMA instruction are listed" % len(counts))
import subprocess
for fn, c in counts.items():
    if not c: continue
    try: name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception: name = fn
    print("%-110s %s" % (name[:110], "  ".join("%s=%d" % kv for kv in sorted(c.items()))))
    tot.update(c)
print("# library totals:", "  ".join("%s=%d" % kv for kv in sorted(tot.items())))
for k in pat:
    if k not in tot: print("# %s: 0" % k)
'
