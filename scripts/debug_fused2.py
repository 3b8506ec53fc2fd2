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
        return st.position.copy(), st.velocity.copy(), st.force.copy(), s.stats()

def cmp(o, tag, kss=((1,), (2,), (3,), (10,), (100,))):
    for exact in (False, True):
        for ks in kss:
            a = run(o, True, exact, ks)
            b = run(o, False, exact, ks)
            d = [float(np.abs(x - y).max()) for x, y in zip(a[:3], b[:3])]
            print(tag, 'exact' if exact else 'fast', ks, d, b[3]['fused_steps'], b[3]['rebuilds'], a[3]['rebuilds'], flush=True)

k = json.load(open('tests/golden/reference_kats.json'))["two_body"]
o = orc.State(np.array(k["pos"]), np.array(k["vel"]), k["mass"], np.array(k["box"]))
cmp(o, 'two_body', ((1,), (2,), (3,), (10,), (100,), (999,)))
o = orc.argon_lattice(10, orc.GAS_CELL, 900.0, 7)
o2 = orc.State(o.pos[:901].copy(), o.vel[:901].copy(), o.mass, o.box)
cmp(o2, 'n901')
o = orc.argon_lattice(100, orc.GAS_CELL, 900.0, 7)
cmp(o, 'n1e6', ((1,), (2,), (10,)))
