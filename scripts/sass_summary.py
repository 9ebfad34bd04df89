"""SASS evidence per kernel: counts of the Blackwell-native mnemonics (B200_PROFILING.md) in the built library.
    python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cogstream_b200", "lib", "libcogstream_b200.so")
sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "STAS", "UCGABAR_ARV", "REDUX", "HMMA", "HADD2.BF16", "LDG.E.128", "STG.E.128",
        "DADD", "DFMA", "MUFU.RSQ64H", "ATOM", "RED", "BAR.SYNC", "ELECT", "ACQBULK", "CCTL", "NANOSLEEP"]
cur = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "UTCHMMA" and op.startswith("UTCHMMA")):
                counts[cur][k] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
print(f"# SASS summary of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); counts per kernel of the mnemonics that matter")
tot = collections.Counter()
for fn, c in counts.items():
    tot.update(c)
    items = ", ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{fn:70s} instr={c['_total']:6d}  {items}")
print()
print("TOTAL " + ", ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
print("tcgen05 (UTCHMMA) present:", tot["UTCHMMA"] > 0, "| TMA tensor loads (UTMALDG):", tot["UTMALDG"] > 0, "| bulk copies (UBLKCP):", tot["UBLKCP"] > 0,
      "| TMEM loads (LDTM):", tot["LDTM"] > 0, "| legacy mma.sync (HMMA):", tot["HMMA"])
