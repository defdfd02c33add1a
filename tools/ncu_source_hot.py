"""Top stall sites of one kernel from `ncu --page source --csv` (needs -lineinfo; SASS view with source correlation).

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:<k> > X_src.csv
    python tools/ncu_source_hot.py X_src.csv [top]
"""
import csv
import sys


def main(path, top=25):
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    hdr = None
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            hdr = i
            break
    cols = {n: j for j, n in enumerate(rows[hdr])}
    body = [r for r in rows[hdr + 1:] if len(r) == len(rows[hdr]) and r[0] != "Address"]
    s_all = cols["Warp Stall Sampling (All Samples)"]
    total = sum(int(r[s_all] or 0) for r in body)
    stall_cols = [n for n in cols if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[cols[n]] or 0) for r in body) for n in stall_cols}
    print("total samples", total)
    print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    body.sort(key=lambda r: -int(r[s_all] or 0))
    for r in body[:top]:
        n = int(r[s_all] or 0)
        reasons = sorted(((int(r[cols[c]] or 0), c[6:]) for c in stall_cols if int(r[cols[c]] or 0)), reverse=True)[:3]
        print(f"{100.0*n/total:5.1f}%  {r[cols['Source']][:90]:90s} {reasons}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
