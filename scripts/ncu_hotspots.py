#!/usr/bin/env python
"""Hottest SASS instructions of one captured kernel, from the source page of an `ncu --set full --import-source on` report:

    python scripts/ncu_hotspots.py <capture.ncu-rep> [<top n>]

prints each instruction's share of the warp-stall samples, its execution count and its dominant stall reasons."""
import csv
import io
import subprocess
import sys


def main(rep, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[0][1][:120] if start else "")
    hdr = rows[start]
    i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data, tot = [], 0
    for n, r in enumerate(rows[start + 1:]):
        if len(r) <= i_s or not r[i_s].isdigit():
            if r and r[0] == "Kernel Name":
                break
            continue
        v = int(r[i_s])
        tot += v
        why = sorted(((int(r[i]) if r[i].isdigit() else 0, h[6:]) for i, h in stalls), reverse=True)[:2]
        data.append((v, n, r[i_src], r[i_ex], why))
    print("total samples", tot)
    for v, n, s, e, why in sorted(data, reverse=True)[:top]:
        print(f"{100 * v / max(tot, 1):5.1f}%  #{n:<5d} exec {e:>9}  {s[:84]:84s} {' '.join(f'{h}={c}' for c, h in why if c)}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
