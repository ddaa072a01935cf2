"""Instruction mix / hottest SASS lines of every kernel in an `ncu --page source --csv` dump.

    ncu -i x.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_mix.py src.csv [top_n]
"""
import collections
import csv
import sys


def main(path, topn):
    kernels = []
    cur = None
    with open(path) as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "Kernel Name":
                cur = {"name": row[1], "hdr": None, "rows": []}
                kernels.append(cur)
            elif cur is not None and cur["hdr"] is None:
                cur["hdr"] = row
            elif cur is not None:
                cur["rows"].append(row)
    seen = set()
    for k in kernels:
        if k["name"] in seen:
            continue
        seen.add(k["name"])
        h = k["hdr"]
        ia, isrc, iex, ism = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        rows = []
        for r in k["rows"]:
            try:
                rows.append((r[ia], r[isrc], int(r[iex]), int(r[ism])))
            except (ValueError, IndexError):
                pass
        tot = sum(x[2] for x in rows) or 1
        tots = sum(x[3] for x in rows) or 1
        print("==", k["name"][:110])
        print("   warp instructions %d, samples %d, SASS lines %d" % (tot, tots, len(rows)))
        op, ops = collections.Counter(), collections.Counter()
        for a, s, ex, sm in rows:
            m = s.split()
            key = (m[1] if m[0].startswith("@") else m[0]).split(".")[0]
            op[key] += ex
            ops[key] += sm
        for key, v in op.most_common(topn):
            print("   %-10s exec %5.1f%%  samples %5.1f%%" % (key, 100.0 * v / tot, 100.0 * ops[key] / tots))
        print("   -- hottest lines by samples")
        for a, s, ex, sm in sorted(rows, key=lambda x: -x[3])[:topn]:
            print("   %5.1f%%  x%-9d %s" % (100.0 * sm / tots, ex, s[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
