#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use, from ``cuobjdump -sass`` of the built library.

    python tools/sass_summary.py [lsdm_b200/liblsdm_b200.so] > profiles/r2_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::tf32 / f16), UTCQMMA = block-scaled, LDTM / STTM = tcgen05.ld / st (tensor memory),
UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA bulk tensor loads / stores, UBLKCP = cp.async.bulk,
SYNCS = mbarrier ops, LDGSTS = cp.async, REDUX = redux.sync, HMMA/IMMA = legacy mma.sync (should be 0).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "REDUX", "HMMA", "IMMA", "FFMA2", "MUFU")


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lsdm_b200", "liblsdm_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MNEMS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
    try:
        import ctypes
        sys.path.insert(0, ROOT)
        from lsdm_b200 import _lib
        ver = _lib.load().lsdm_version().decode()
    except Exception as e:  # pragma: no cover
        ver = f"(lsdm_version unavailable: {e})"
    print(f"# {os.path.relpath(so, ROOT)}  --  {ver}")
    print("# kernel\tinstructions\t" + "\t".join(MNEMS))
    tot = collections.Counter()
    for fn in order:
        c = counts[fn]
        if c["_total"] == 0:
            continue
        name = demangle(fn)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\((int|bool|unsigned int)\)", "", name)   # template-argument casts
        name = re.sub(r"\(.*\)$", "", name)                       # the parameter list
        print(name[:110] + "\t" + str(c["_total"]) + "\t" + "\t".join(str(c[k]) for k in MNEMS))
        tot.update(c)
    print("TOTAL\t" + str(tot["_total"]) + "\t" + "\t".join(str(tot[k]) for k in MNEMS))


if __name__ == "__main__":
    main()
