#!/usr/bin/env python
"""Records the DRAM traffic of one kernel from an `ncu --set full` capture in profiles/ncu_traffic.json.

  python scripts/ncu_traffic.py <capture.ncu-rep> <workload> <kernel-regex> [<key>]

dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches whose name matches, stored under
"<workload>:<key or kernel>" together with the commit the capture was taken at and the capture's file name.  bench.py prints
these as roofline.traffic / traffic_source / traffic_commit (null when no capture of the running build's kernel exists)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(rep, workload, pattern, key=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    tot, dur, cnt = 0.0, 0.0, 0
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if not re.search(pattern, name):
            continue
        tot += float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
        dur += float(r[it].replace(",", ""))
        cnt += 1
    if not cnt:
        raise SystemExit(f"no launch matching {pattern!r} in {rep}")
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[f"{workload}:{key or pattern}"] = {
        "dram_bytes_per_launch": tot / cnt, "launches": cnt, "duration_under_ncu": f"{dur / cnt:.1f} {units[it]}",
        "source": os.path.basename(rep), "commit": commit,
    }
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(f"{workload}:{key or pattern}", data[f"{workload}:{key or pattern}"])


if __name__ == "__main__":
    main(*sys.argv[1:5])
