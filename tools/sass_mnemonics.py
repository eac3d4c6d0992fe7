#!/usr/bin/env python
"""profiles/r02_sass_mnemonics.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (tcgen05 /
TMEM / TMA / bulk copies) in the shipped libgca_b200.so.  CPU only: `cuobjdump -sass` on the built library."""
import collections, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "video-graph-ssl_b200", "gca_b200", "libgca_b200.so")
COLS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU.EX2", "HMMA", "FFMA2", "FADD2",
        "RED.E", "LDG.E", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "MEMBAR", "UCGABAR"]
exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
txt = subprocess.run([exe, "-sass", SO], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for c in COLS:
            if op == c or op.startswith(c + ".") or (c.endswith(".SYS") and c in op):
                counts[cur][c] += 1
out = ["# r02 (final) -- SASS mnemonic counts per kernel of the shipped libgca_b200.so (`cuobjdump -sass`, sm_100a; tools/sass_mnemonics.py); static instruction counts.",
       "# tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, TMA tensor loads -> UTMALDG, bulk copies -> UBLKCP, tcgen05.commit -> UTCBAR,",
       "# mbarrier -> SYNCS; HMMA (legacy mma.sync) must be 0 everywhere.  *.SYS columns: system-scope stores/loads of the peer-memory steps.",
       "# kernel | " + " | ".join(COLS)]
for k, c in counts.items():
    if any(c.values()):
        out.append(k + " | " + " | ".join(str(c[x]) for x in COLS))
open(os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt"), "w").write("\n".join(out) + "\n")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print({k: tot[k] for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "HMMA")}, len(counts), "kernels")
