import sys, json, numpy as np
sys.path.insert(0, '.')
import moldyn_b200 as md
from oracle import oracle as orc
DT = 0.002
o = orc.argon_lattice(100, orc.GAS_CELL, 273.15, 42)
res = []
for split in (True, False):
    st = md.State(o.pos, o.vel, o.mass, o.box)
    with md.Solver(split_step=split) as s:
        s.upload(st, with_forces=False); s.update_force()
        th = (md.Thermostat.Berendsen(10.0), 300.0)
        for k in (1900, 1, 1, 1, 97, 3000, 5000):
            s.step(k, DT, thermostat=th)
        s.download(st)
        res.append((st.position.copy(), st.velocity.copy(), s.macro()['temperature'], s.stats()))
print('maxdiff', np.abs(res[0][0]-res[1][0]).max(), np.abs(res[0][1]-res[1][1]).max(), res[0][2], res[1][2], res[1][3]['fused_steps'], res[1][3]['rebuilds'], res[0][3]['rebuilds'])
