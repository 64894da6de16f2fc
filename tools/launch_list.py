"""Per-kernel times from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_list.py gpurun_out/x.csv [--last-fraction 0.5]"""
import csv
import sys


def main():
    path = sys.argv[1]
    frac = float(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[2] == "--last-fraction" else 1.0
    hdr, rows = None, []
    for r in csv.reader(open(path)):
        if r and r[0] == "ID":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            rows.append(r)
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    rows = rows[int(len(rows) * (1 - frac)):]
    tot = 0.0
    for r in rows:
        v = float(r[vi].replace(",", ""))
        us = v / 1000 if r[ui].startswith("n") else v
        tot += us
        print(f"{r[ki][:48]:48s} {us:9.1f} us  grid {r[gi]}")
    print(f"total {tot:.1f} us over {len(rows)} launches")


if __name__ == "__main__":
    main()
