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
for s in (sa, sb):
    s.step(1900, DT)
done = 1900
prev_rb = sb.stats()['rebuilds']
while done < 2100:
    for s, st in zip((sa, sb), sts):
        s.step(1, DT); s.download(st)
    done += 1
    d = [np.abs(sts[0].position[:, k] - sts[1].position[:, k]) > 0 for k in range(3)] + [np.abs(sts[0].velocity[:, k] - sts[1].velocity[:, k]) > 0 for k in range(3)]
    anyb = np.zeros(len(d[0]), bool)
    for x in d: anyb |= x
    bad = np.nonzero(anyb)[0]
    rb = sb.stats()['rebuilds']
    if len(bad) or rb != prev_rb:
        print(done, 'bad', len(bad), 'per-plane', [int(x.sum()) for x in d], 'rebuilds', sa.stats()['rebuilds'], rb, 'fused', sb.stats()['fused_steps'], flush=True)
    prev_rb = rb
    if len(bad):
        cell, dims = sb.cells()
        order = np.lexsort((np.arange(sb.n), cell))   # sorted slot -> atom id
        slot = np.empty(sb.n, dtype=np.int64); slot[order] = np.arange(sb.n)
        sl = np.sort(slot[bad])
        print('slots', sl[:10], '...', sl[-10:], 'tiles', np.unique(sl // 256), 'tile mod 592', np.unique(sl // 256) % 592)
        # where did the wrong x come from?
        xb = sts[1].position[bad[0], 0]
        src = np.nonzero(sts[0].position[:, 0] == xb)[0]
        print('bad atom', bad[0], 'slot', slot[bad[0]], 'fused x', xb, 'equals split x of atoms', src, 'slots', slot[src] if len(src) else None)
        break
