"""Top SASS instructions of an `ncu --page source --csv` dump by warp-stall samples, with the dominant stall reasons.

  ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N]
"""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    col = {name: i for i, name in enumerate(h)}
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "(Not Issued)" not in n]
    data = []
    for r in rows[hdr + 1:]:
        if len(r) < len(h):
            continue
        try:
            smp = int(r[col["# Samples"]])
        except ValueError:
            continue
        data.append((smp, r))
    total = sum(s for s, _ in data) or 1
    print(f"total samples {total}")
    for idx, (smp, r) in enumerate(data):
        r.append(idx)
    for smp, r in sorted(data, key=lambda t: -t[0])[:top]:
        st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:3]
        print(f"{100.0 * smp / total:5.1f}%  #{r[-1]:5d}  {r[col['Source']].strip()[:70]:70s}  " +
              " ".join(f"{n}:{v}" for v, n in st if v))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
