"""Phase timing of k_force from %globaltimer probes (build with MD_NVCC_EXTRA=-DMD_TIMING_PROBES)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import moldyn_b200 as md
from moldyn_b200 import _ffi
from bench import make_state, WORKLOADS, ARGON_MASS, DT
L = _ffi.lib()
L.md_probe_read.argtypes = [C.c_void_p, C.c_void_p]; L.md_probe_reset.argtypes = [C.c_void_p]
for wl, warm in (("c2", 200), ("c3", 200), ("c3", 12000)):
    w = WORKLOADS[wl]
    pos, vel, box = make_state(w)
    s = md.Solver(host_loop=True)
    s.upload_arrays(pos, vel, ARGON_MASS, box); s.update_force()
    th = (md.Thermostat.Berendsen(10.0), 300.0)
    s.step(warm, DT, thermostat=th)
    acc = np.zeros(5); n = 0
    for _ in range(50):
        L.md_probe_reset(s._ctx)
        s.step(1, DT, thermostat=th)
        p = (C.c_ulonglong * 8)(); L.md_probe_read(s._ctx, p)
        t = np.array(p[:5], dtype=np.float64)
        acc += t - t[0]; n += 1
    print(wl, "warm", warm, "ns from first block start: loop_end %.0f  last_block_start %.0f  reduced %.0f  finalized %.0f" % tuple(acc[1:] / n))
    s.close()
