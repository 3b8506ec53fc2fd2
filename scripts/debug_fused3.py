import sys, json, numpy as np
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
        return st.position.copy(), st.velocity.copy(), st.force.copy(), s.stats(), s.macro()

def cmp(o, tag, kss, hls=(False, True)):
    for ks in kss:
        for hl in hls:
            a = run(o, True, False, ks, hl)
            b = run(o, False, False, ks, hl)
            d = [float(np.abs(x - y).max()) for x, y in zip(a[:3], b[:3])]
            print(tag, ks, 'host' if hl else 'graph', d, b[3]['fused_steps'], b[3]['rebuilds'], a[3]['rebuilds'], b[4]['temperature'], a[4]['temperature'], flush=True)

o = orc.argon_lattice(56, orc.GAS_CELL, 900.0, 7)
cmp(o, 'n175616', ((150,), (400,)))
o = orc.argon_lattice(100, orc.GAS_CELL, 900.0, 7)
cmp(o, 'n1e6', ((120,), (300,)))
