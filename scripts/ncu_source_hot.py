"""Top stalled SASS lines per kernel from an `ncu --page source --csv` export (optionally gz)."""
import collections
import csv
import gzip
import re
import sys


def main(path, which=0, top=30):
    op = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(op))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    k = kernels[which]
    h = k["hdr"]
    iS, iI, iSamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iI]) for r in k["rows"])
    tots = sum(int(r[iSamp]) for r in k["rows"])
    print(k["name"], "instr", tot, "samples", tots, "kernels in file", len(kernels))
    st = collections.Counter()
    for r in k["rows"]:
        for i in stall_cols:
            st[h[i]] += int(r[i])
    print({a: b for a, b in st.most_common(8)})
    for idx, r in sorted(enumerate(k["rows"]), key=lambda t: -int(t[1][iSamp]))[:top]:
        reasons = sorted(((int(r[i]), h[i]) for i in stall_cols if int(r[i])), reverse=True)[:2]
        print(f"{idx:5d} samp {int(r[iSamp]):6d} inst {int(r[iI]):8d}  {r[iS][:70]:70s} {reasons}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 30)
