"""Blackwell tensor-core / TMEM / bulk-copy mnemonics per kernel of the shipped library (cuobjdump -sass):
usage: python scripts/sass_counts.py [relearn_b200/librelearn_b200.so] > profiles/<round>_sass_tensor_ops.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "relearn_b200/librelearn_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
MNEMONICS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA2", "DFMA")
counts, cur, order = collections.OrderedDict(), None, iter(names)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(order)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        for k in MNEMONICS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (sm_100a); only kernels with tensor-core / TMEM / bulk-copy instructions, then totals")
print("# UTCHMMA = tcgen05.mma (bf16/f16 kind), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit -> mbarrier, UBLKCP = cp.async.bulk, "
      "SYNCS = mbarrier ops")
tot = collections.Counter()
for name, c in counts.items():
    tot.update(c)
    if any(c[k] for k in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR")):
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        short = re.sub(r"\(.*", "", short)
        print(f"{short[:90]:90s} " + " ".join(f"{k}={c[k]}" for k in MNEMONICS if c[k]))
print("TOTAL " + " ".join(f"{k}={tot[k]}" for k in MNEMONICS))
print(f"kernels: {len(counts)}")
