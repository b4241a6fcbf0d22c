"""Transpose `ncu --page raw --csv` exports (one very wide row per kernel) into metric,value,unit tables."""
import csv
import sys


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows[2:]:
            w.writerow(["Kernel Name", r[hdr.index("Kernel Name")], ""])
            for h, u, v in zip(hdr, units, r):
                if v != "" and h not in ("ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Device", "CC"):
                    w.writerow([h, v, u])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
