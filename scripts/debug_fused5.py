import sys, json, numpy as np
sys.path.insert(0, '.')
import moldyn_b200 as md
from oracle import oracle as orc
DT = 0.002
o = orc.argon_lattice(100, orc.GAS_CELL, 273.15, 42)
sa = md.Solver(split_step=True, host_loop=True); sb = md.Solver(split_step=False, host_loop=True)
sts = []
for s in (sa, sb):
    st = md.State(o.pos, o.vel, o.mass, o.box); sts.append(st)
    s.upload(st, with_forces=False); s.update_force()
done = 0
chunk = 100
while done < 2600:
    for s, st in zip((sa, sb), sts):
        s.step(chunk, DT); s.download(st)
    done += chunk
    dp = np.abs(sts[0].position - sts[1].position).max(axis=1)
    dv = np.abs(sts[0].velocity - sts[1].velocity).max(axis=1)
    bad = np.nonzero((dp > 0) | (dv > 0))[0]
    print(done, len(bad), sa.stats()['rebuilds'], sb.stats()['rebuilds'], sb.stats()['fused_steps'], flush=True)
    if len(bad):
        cnts = np.zeros(sb.n, dtype=np.int64)
        off, par = sb.neighbour_lists()
        cnt = off[1:] - off[:-1]
        print('bad atoms', bad[:20], 'counts', cnt[bad[:20]], 'max cnt', cnt.max(), 'hist', np.bincount(cnt))
        for i in bad[:6]:
            print(i, sts[0].position[i], sts[1].position[i], sts[0].velocity[i], sts[1].velocity[i], 'partners', par[off[i]:off[i+1]])
        if chunk == 1 or len(bad) > 1000: break
        break
