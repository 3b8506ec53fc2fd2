import sys, json, numpy as np
sys.path.insert(0, '.')
import moldyn_b200 as md
from oracle import oracle as orc
DT = 0.002

def run(o, split, ksteps, host_loop=False, th=None):
    st = md.State(o.pos, o.vel, o.mass, o.box)
    with md.Solver(split_step=split, host_loop=host_loop) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        for k in ksteps:
            s.step(k, DT, thermostat=(md.Thermostat.Berendsen(10.0), 300.0) if th else None)
            print('   after', k, s.macro()['temperature'], s.stats()['rebuilds'], s.stats()['fused_steps'], flush=True)
        s.download(st)
        return st.position.copy(), st.velocity.copy(), st.force.copy(), s.stats(), s.macro()

def cmp(o, tag, kss, hls=(False,), th=True):
    for ks in kss:
        for hl in hls:
            a = run(o, True, ks, hl, th)
            b = run(o, False, ks, hl, th)
            d = [float(np.abs(x - y).max()) for x, y in zip(a[:3], b[:3])]
            print(tag, ks, 'host' if hl else 'graph', d, b[3]['fused_steps'], b[3]['rebuilds'], a[3]['rebuilds'], b[4]['temperature'], a[4]['temperature'], flush=True)

o = orc.argon_lattice(100, orc.GAS_CELL, 273.15, 42)
cmp(o, 'n1e6 th', ((100, 200, 300, 400, 1000, 1000),), (False, True))
cmp(o, 'n1e6 noth', ((100, 200, 300, 400, 1000, 1000),), (False,), th=False)
