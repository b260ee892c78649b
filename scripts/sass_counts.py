"""SASS mnemonic counts of the shipped library (cuobjdump -sass; runs without a GPU):

    python scripts/sass_counts.py > profiles/<round>_sass_counts.txt

The file is the evidence that the convolution kernels are tcgen05 / TMEM / bulk-copy code (UTCHMMA = tcgen05.mma, LDTM =
tcgen05.ld, UBLKCP = cp.async.bulk, UTMALDG / UTMASTG = cp.async.bulk.tensor, UCGABAR = cluster barrier) and of what
they are not (no tensor-map TMA: operand tiles are gathered and hi/lo-split by producer warps).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "corenet_b200", "lib", "libcorenet_b200.so")
KEYS = ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "ELECT", "HMMA", "FFMA", "DFMA",
        "ATOMG", "REDG")


def main():
  sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
  head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
  total = collections.Counter()
  per = collections.defaultdict(collections.Counter)
  fn = None
  for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
      fn = m.group(1)
      continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or fn is None:
      continue
    op = m.group(1)
    for k in KEYS:
      if op == k or op.startswith(k + ".") or op.startswith(k + "_"):
        total[k] += 1
        per[fn][k] += 1
  print(f"# SASS mnemonic counts of corenet_b200/lib/libcorenet_b200.so (cuobjdump -sass, sm_100a) at {head}")
  print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM load), UBLKCP = cp.async.bulk (TMA bulk copy), UTMALDG/UTMASTG =")
  print("# cp.async.bulk.tensor (none: operand tiles are gathered + hi/lo-split by producer warps, weights stream through")
  print("# cp.async.bulk), SYNCS = mbarrier ops, UCGABAR = cluster barrier, ELECT = elect.sync, HMMA = legacy mma.sync, REDG = global reductions")
  print("# (the float ones are the split-K / accumulator flushes whose order makes train-mode runs differ in the last bits)")
  for k in KEYS:
    print(f"{k:10s} {total[k]:7d}")
  print("\n# per kernel (functions with tcgen05 / bulk-copy / cluster instructions): UTCHMMA LDTM UBLKCP UCGABAR ELECT  name")
  rows = [(c["UTCHMMA"], c["LDTM"], c["UBLKCP"], c["UCGABAR"], c["ELECT"], f) for f, c in per.items()
          if c["UTCHMMA"] or c["UBLKCP"] or c["UCGABAR"] or c["LDTM"]]
  for r in sorted(rows, reverse=True):
    name = subprocess.run(["c++filt", r[5]], capture_output=True, text=True).stdout.strip()
    print(f"{r[0]:6d} {r[1]:5d} {r[2]:5d} {r[3]:5d} {r[4]:5d}  {name[:150]}")


if __name__ == "__main__":
  sys.exit(main())
