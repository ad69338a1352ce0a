"""Opcode evidence for the tcgen05 / TMEM / TMA path: per-kernel histogram of the Blackwell-native SASS mnemonics in
vcvits_b200/libvcd.so (B200_PROFILING.md "What proves a Blackwell-native kernel").  Usage:
    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "vcvits_b200", "libvcd.so")
WANT = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UBLKRED", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "HMMA", "HGMMA",
        "LDGSTS", "RED", "ELECT", "FFMA")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        cur["_total"] += 1
        if op in WANT:
            cur[op] += 1
print(f"# SASS opcode histogram of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, cp.async.bulk.tensor -> UTMALDG,")
print("# mbarrier -> SYNCS; HMMA / HGMMA (legacy mma.sync / Hopper wgmma) must be absent")
tot = collections.Counter()
for name, c in per.items():
    keys = [k for k in WANT if c[k]]
    if not any(k in ("UTCHMMA", "LDTM", "UBLKCP", "UTMALDG", "UTCBAR") for k in keys) and "tc::" not in name:
        continue
    print(f"{name}: {c['_total']} instructions; " + ", ".join(f"{k} {c[k]}" for k in keys))
    tot.update({k: c[k] for k in keys})
print("TOTAL (kernels above): " + ", ".join(f"{k} {tot[k]}" for k in WANT if tot[k]))
print("legacy tensor opcodes in the whole library: HMMA", sum(c["HMMA"] for c in per.values()), " HGMMA", sum(c["HGMMA"] for c in per.values()))
