import sys, numpy as np
sys.path.insert(0, '.')
import moldyn_b200 as md
from oracle import oracle as orc
DT = 0.002

def run(o, split, exact, ksteps, host_loop=False, th=None):
    st = md.State(o.pos, o.vel, o.mass, o.box)
    with md.Solver(exact=exact, split_step=split, host_loop=host_loop) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        for k in ksteps:
            s.step(k, DT, thermostat=th)
        s.download(st)
        return st.position.copy(), st.velocity.copy(), st.force.copy(), s.stats()

for side, temp in ((10, 900.0), (40, 900.0), (56, 900.0)):
    o = orc.argon_lattice(side, orc.GAS_CELL, temp, 7)
    for exact in (False, True):
        for ks in ((1,), (2,), (3,), (1, 1, 1), (10,), (100,)):
            for hl in (False, True):
                a = run(o, True, exact, ks, hl)
                b = run(o, False, exact, ks, hl)
                d = [float(np.abs(x - y).max()) for x, y in zip(a[:3], b[:3])]
                print(side**3, 'exact' if exact else 'fast', ks, 'host' if hl else 'graph', d, b[3]['fused_steps'], b[3]['rebuilds'], a[3]['rebuilds'], flush=True)
