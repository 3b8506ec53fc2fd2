#!/usr/bin/env python
"""Where a step of the persistent loop spends its time: %globaltimer stamps of one step (MD_LOOP_TRACE build of the library,
loaded through MOLDYN_B200_LIBRARY).  python scripts/loop_trace.py c1|c2|c3 [steps_before]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import moldyn_b200 as md  # noqa: E402
from moldyn_b200 import _ffi  # noqa: E402

w = bench.WORKLOADS[sys.argv[1]]
before = int(sys.argv[2]) if len(sys.argv) > 2 else 500
pos, vel, box = bench.make_state(w)
s = md.Solver()
s.upload_arrays(pos, vel, bench.ARGON_MASS, box)
s.update_force()
th = (md.Thermostat.Berendsen(w["thermostat"][0]), w["thermostat"][1])
ba = (md.Barostat.Berendsen(w["barostat"][0], w["barostat"][1]), w["barostat"][2]) if w["barostat"] else None
s.step(before, bench.DT, thermostat=th, barostat=ba)
L = _ffi.lib()
L.md_debug_trace.argtypes = [C.c_void_p, C.c_void_p]
names = {0: "step start", 1: "drift done", 2: "barrier passed", 3: "forces done", 4: "block sums", 5: "ticket taken",
         6: "released (next step)", 8: "last block: epilogue", 9: "last block: folded", 10: "last block: finalized",
         11: "last block: seq released"}
acc = {}
for rep in range(20):
    s.step(7, bench.DT, thermostat=th, barostat=ba)
    t = np.zeros(16, dtype=np.uint64)
    L.md_debug_trace(s._ctx, t.ctypes.data_as(C.c_void_p))
    t0 = int(t[0])
    for k in names:
        acc.setdefault(k, []).append((int(t[k]) - t0) * 1e-3)
print(sys.argv[1], "stats", {k: s.stats()[k] for k in ("nbr_mean", "persistent_loop", "rebuilds")})
for k in sorted(names):
    v = np.array(acc[k])
    print(f"  [{k:2d}] {names[k]:28s} median {np.median(v):8.2f} us   min {v.min():8.2f}")
