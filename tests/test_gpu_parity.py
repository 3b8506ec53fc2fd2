"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star / SURVEY §8d):
  * cell assignments and neighbour sets: bit-exact;
  * MD_FORCE_EXACT: per-atom force / potential / virial and NVE trajectories BIT-IDENTICAL to the oracle;
  * MD_FORCE_FAST: |Δ| <= 1e-10 * max(Σ_j|f_ij|, global RMS) per atom; 100-step trajectories within 1e-8;
  * thermostat / barostat runs: 1e-8 over 100 steps (the reductions are tree sums, not sequential sums);
  * the reference's golden values to 8 decimals through the GPU path.
"""
import numpy as np
import pytest

import moldyn_b200 as md
from oracle import oracle as orc

from helpers import LONG_CUT, dense_gas, fmt3, f8, force_scale, gas, lj_pair, liquid, to_gpu_state

pytestmark = pytest.mark.gpu

DT = 0.002


@pytest.fixture(scope="module", params=["exact", "fast"])
def mode(request):
    return request.param


def make_solver(mode, **kw):
    return md.Solver(exact=(mode == "exact"), **kw)


# ---------------------------------------------------------------------------------------------------
# the reference's own golden values, through the GPU path
def two_body(kats):
    k = kats["two_body"]
    return md.State(k["pos"], k["vel"], k["mass"], k["box"]), k


def check_golden_step(st, g):
    for key, arr in (("pos1", st.position[0]), ("pos2", st.position[1]), ("vel1", st.velocity[0]),
                     ("vel2", st.velocity[1]), ("force1", st.force[0]), ("force2", st.force[1])):
        if key in g:
            assert fmt3(arr) == g[key], (g.get("step"), key)


@pytest.mark.parametrize("host_loop", [False, True])
def test_golden_verlet_with_lennard_jones(kats, mode, host_loop):  # solver/src/lib.rs:109-262
    st, k = two_body(kats)
    with make_solver(mode, host_loop=host_loop) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        s.download(st)
        check_golden_step(st, k["steps"][0])
        for g in k["steps"][1:]:
            s.step(1, k["dt"])
            s.download(st)
            check_golden_step(st, g)


def test_golden_per_call_api(kats):  # cli/src/tests.rs:35-98 semantics: State in, State out, every call
    st, k = two_body(kats)
    db = md.PotentialsDatabase()
    md.update_force(db, st)
    check_golden_step(st, k["steps"][0])
    for g in k["steps"][1:]:
        md.Integrator.VerletMethod.calculate(db, st, k["dt"], None, None)
        check_golden_step(st, g)


def test_per_call_api_npt_matches_oracle():
    """Integrator::calculate on a HOST State, every step, with thermostat + barostat: the incoming virial feeds
    calculate_myu (barostat.rs:23-29), the incoming potential is dead — 25 calls must track the oracle."""
    o = liquid(8)
    olj = orc.LennardJones()
    orc.update_force(olj, o, mode="cells")
    st = to_gpu_state(md, o)
    db = md.PotentialsDatabase()
    oth, gth = orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, 120.0), (md.Thermostat.Berendsen(10.0), 120.0)
    oba, gba = orc.Barostat(1.0, 5.0, 1.01325), (md.Barostat.Berendsen(1.0, 5.0), 1.01325)
    for _ in range(25):
        st.potential[:] = 123.0  # dead on entry
        md.Integrator.VerletMethod.calculate(db, st, DT, gba, gth)
    orc.step(olj, o, DT, thermostat=oth, barostat=oba, mode="cells", n_steps=25)
    assert np.abs(st.boundary_box / o.box - 1.0).max() <= 1e-12
    assert abs(gba[0].myu / oba.myu - 1.0) <= 1e-12 and abs(gth[0].lambda_ / oth.lambda_ - 1.0) <= 1e-12
    dx = np.abs(st.position - o.pos)
    assert np.minimum(dx, np.abs(dx - o.box)).max() <= 1e-8
    assert np.abs(st.velocity - o.vel).max() <= 1e-8
    assert np.abs(st.potential - o.pot).max() <= 1e-9 * max(1.0, np.abs(o.pot).max())
    assert np.abs(st.temp - o.vir).max() <= 1e-9 * max(1.0, np.abs(o.vir).max())


def test_golden_1000_iterations_and_macro(kats, mode):  # solver/src/lib.rs:264-332
    st, k = two_body(kats)
    g = kats["verlet_lj_1000_iterations"]
    with make_solver(mode) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        s.step(g["n_steps"], k["dt"])
        s.download(st)
        m = s.macro()
    check_golden_step(st, g)
    assert f8(m["kinetic"]) == g["kinetic"]
    assert f8(m["thermal"]) == g["thermal"]
    assert f8(m["potential"]) == g["potential"]
    assert f8(m["thermal"] + m["potential"]) == g["internal"]
    assert f8(m["kinetic"] + m["potential"]) == g["full"]
    assert f8(m["temperature"] / 100.0) == g["temperature_over_100"]
    assert f8(m["pressure"]) == g["pressure"]


def test_golden_energies_temperature_pressure(kats, mode):  # solver/src/lib.rs:334-427
    st, _ = two_body(kats)
    with make_solver(mode) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        m = s.macro()
    g = kats["energies"]
    assert list(m["vcom"]) == g["vcom"]
    assert f8(m["kinetic"]) == g["kinetic"]
    assert f8(m["thermal"]) == g["thermal"]
    assert f8(m["potential"]) == g["potential"]
    assert f8(m["temperature"]) == kats["temperature"]["value"]
    assert f8(m["pressure"]) == kats["pressure"]["value"]


def test_golden_update_force(kats, mode):  # solver/src/lib.rs:91-107
    k = kats["update_force_lennard_jones"]
    st = md.State(k["pos"], np.zeros((2, 3)), k["mass"], k["box"])
    with make_solver(mode) as s:
        s.update_force_host(st)
    assert fmt3(st.force[0]) == k["force_p1"]


def test_momentum_invariant(kats, mode):  # solver/src/lib.rs:49-75
    g = kats["momentum"]
    o = orc.argon_lattice(tuple(g["grid"]), g["cell"], g["temperature"])
    st = to_gpu_state(md, o)
    with make_solver(mode) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        for _ in range(20):
            s.step(500, g["dt"])
            s.download(st)
            assert np.all(np.abs(st.velocity.sum(axis=0)) < g["tolerance"])


# ---------------------------------------------------------------------------------------------------
SYSTEMS = {
    "gas1000": lambda: (gas(10), None),
    "dense_gas3000": lambda: (dense_gas(3000), None),
    "liquid1000": lambda: (liquid(10), None),
    "liquid4096_3.5sigma": lambda: (liquid(16), LONG_CUT),
    "liquid_small_box": lambda: (liquid(5), None),   # box 1.81 nm: fewer than 3 cells per axis
    "liquid5832": lambda: (liquid(18), None),        # 13 cells per axis: the dense FAST path runs the tile kernels
    "liquid729": lambda: (liquid(9), None),          # odd atom count: the last thread owns one atom
}


@pytest.mark.parametrize("name", list(SYSTEMS))
def test_forces_match_oracle(name, mode):
    o, cut = SYSTEMS[name]()
    olj, plj = lj_pair(md, *(cut or (None, None)))
    orc.update_force(olj, o, mode="n2")
    st = to_gpu_state(md, o)
    with make_solver(mode) as s:
        s.set_potential(plj)
        s.upload(st, with_forces=False)
        s.update_force()
        s.download(st)
    if mode == "exact":
        assert np.array_equal(st.force, o.force)
        assert np.array_equal(st.potential, o.pot)
        assert np.array_equal(st.temp, o.vir)
    else:
        scale = force_scale(olj, o)
        rms = np.sqrt((o.force ** 2).sum(axis=1).mean())
        tol = 1e-10 * np.maximum(scale, rms)[:, None]
        assert np.all(np.abs(st.force - o.force) <= tol + 1e-300)
        assert np.all(np.abs(st.potential - o.pot) <= 1e-10 * (np.abs(o.pot) + 4 * olj.eps))
        assert np.all(np.abs(st.temp - o.vir) <= 1e-10 * (scale * olj.r_cut + 1e-300) + 1e-300)


@pytest.mark.parametrize("name", ["dense_gas3000", "liquid1000", "gas1000", "liquid5832"])
def test_cells_and_neighbour_sets_bit_exact(name, mode):
    o, cut = SYSTEMS[name]()
    olj, plj = lj_pair(md, *(cut or (None, None)))
    st = to_gpu_state(md, o)
    with make_solver(mode) as s:
        s.set_potential(plj)
        s.upload(st, with_forces=False)
        s.update_force()
        cell, dims = s.cells()
        off, partners = s.neighbour_lists()
        skin = s.stats()["skin"]
    # cell assignment: c_d = min(nc_d-1, (int)(frac(x_d / L_d) * nc_d)), linear index (cx*ny + cy)*nz + cz
    c = []
    for d in range(3):
        f = o.pos[:, d] / o.box[d]
        f = f - np.floor(f)
        c.append(np.minimum((f * float(dims[d])).astype(np.int64), dims[d] - 1))
    want_cell = (c[0] * dims[1] + c[1]) * dims[2] + c[2]
    assert np.array_equal(cell.astype(np.int64), want_cell)
    # neighbour sets: the reference's pair predicate widened by the skin
    woff, wpartners = orc.neighbour_sets(o.pos, o.box, olj.r_cut + skin)
    assert np.array_equal(off, woff)
    assert np.array_equal(partners, wpartners)


@pytest.mark.parametrize("name", ["liquid4096_3.5sigma", "liquid5832", "liquid13824"])
def test_tile_kernels_dense(name):
    """Dense systems in MD_FORCE_FAST on one GPU run the tile kernels (md_tile.cuh: brick order, TMA-staged shells, 16-bit
    brick-local lists, warp-cooperative pair loop): pair sets bit-exact, per-atom force / potential / virial within the FAST
    bar, 100-step NVT and NPT trajectories within 1e-8 of the oracle (forces-only, + virial and + potential instances of the
    loop all run), run-to-run bit-reproducible."""
    if name == "liquid13824":
        o, cut = liquid(24), None      # 6 x 6 x 6 bricks, partial ones at the upper faces
    else:
        o, cut = SYSTEMS[name]()
    olj, plj = lj_pair(md, *(cut or (None, None)))
    ref = o.copy()
    orc.update_force(olj, ref, mode="cells")
    scale = force_scale(olj, ref)
    rms = np.sqrt((ref.force ** 2).sum(axis=1).mean())
    runs = []
    for ensemble in ("nvt", "npt", "npt"):
        st = to_gpu_state(md, o)
        gth, oth = (md.Thermostat.Berendsen(10.0), 120.0), orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, 120.0)
        gba = oba = None
        if ensemble == "npt":
            gba, oba = (md.Barostat.Berendsen(1.0, 5.0), 1.01325), orc.Barostat(1.0, 5.0, 1.01325)
        with md.Solver() as s:
            s.set_potential(plj)
            s.upload(st, with_forces=False)
            s.update_force()
            stats = s.stats()
            assert stats["tile_lists"] == 1 and stats["nbr_mean"] >= 8.0, stats
            if ensemble == "nvt":
                off, partners = s.neighbour_lists()
                woff, wpartners = orc.neighbour_sets(o.pos, o.box, olj.r_cut + stats["skin"])
                assert np.array_equal(off, woff) and np.array_equal(partners, wpartners)
                cell, dims = s.cells()
                frac = o.pos / o.box
                frac -= np.floor(frac)
                c3 = np.minimum((frac * dims).astype(np.int64), dims - 1)
                assert np.array_equal(cell, (c3[:, 0] * dims[1] + c3[:, 1]) * dims[2] + c3[:, 2])
            s.download(st)
            assert np.all(np.abs(st.force - ref.force) <= 1e-10 * np.maximum(scale, rms)[:, None])
            assert np.all(np.abs(st.potential - ref.pot) <= 1e-10 * (np.abs(ref.pot) + 4 * olj.eps))
            assert np.all(np.abs(st.temp - ref.vir) <= 1e-10 * (scale * olj.r_cut + 1e-300) + 1e-300)
            for k in (1, 36, 63):
                s.step(k, DT, thermostat=gth, barostat=gba)
            s.download(st)
            m = s.macro()
            assert s.stats()["tile_lists"] == 1
        r = o.copy()
        run_oracle(olj, r, 100, oth, oba)
        dx = np.abs(st.position - r.pos)
        assert np.minimum(dx, np.abs(dx - r.box)).max() <= 1e-8
        assert np.abs(st.velocity - r.vel).max() <= 1e-8
        assert np.abs(st.boundary_box - r.box).max() <= 1e-9
        runs.append((st.position.copy(), st.velocity.copy(), st.force.copy(), st.temp.copy(), m["pressure"]))
    for a_, b_ in zip(runs[1][:4], runs[2][:4]):
        assert np.array_equal(a_, b_)
    assert runs[1][4] == runs[2][4]


def run_oracle(olj, o, n_steps, th=None, ba=None):
    orc.update_force(olj, o, mode="cells")
    orc.step(olj, o, DT, thermostat=th, barostat=ba, mode="cells", n_steps=n_steps)


@pytest.mark.parametrize("name", ["liquid1000", "dense_gas3000", "gas1000", "liquid5832"])
def test_nve_trajectory_100_steps(name, mode):
    o, cut = SYSTEMS[name]()
    olj, plj = lj_pair(md, *(cut or (None, None)))
    st = to_gpu_state(md, o)
    with make_solver(mode) as s:
        s.set_potential(plj)
        s.upload(st, with_forces=False)
        s.update_force()
        s.step(100, DT)
        s.download(st)
        stats = s.stats()
    run_oracle(olj, o, 100)
    if mode == "exact":
        assert np.array_equal(st.position, o.pos)
        assert np.array_equal(st.velocity, o.vel)
        assert np.array_equal(st.force, o.force)
    else:
        dx = np.abs(st.position - o.pos)
        dx = np.minimum(dx, np.abs(dx - o.box))  # an atom may sit on either side of the wrap
        assert dx.max() <= 1e-8
        assert np.abs(st.velocity - o.vel).max() <= 1e-8
    assert stats["steps"] == 100


@pytest.mark.parametrize("name", ["liquid1000", "gas1000"])
@pytest.mark.parametrize("ensemble", ["nvt", "npt"])
def test_thermostat_barostat_trajectory(name, ensemble, mode):
    o, cut = SYSTEMS[name]()
    olj, plj = lj_pair(md, *(cut or (None, None)))
    st = to_gpu_state(md, o)
    t0 = 300.0 if name == "gas1000" else 120.0
    oth, gth = orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, t0), (md.Thermostat.Berendsen(10.0), t0)
    oba = gba = None
    if ensemble == "npt":
        oba, gba = orc.Barostat(1.0, 5.0, 1.01325), (md.Barostat.Berendsen(1.0, 5.0), 1.01325)
    with make_solver(mode) as s:
        s.set_potential(plj)
        s.upload(st, with_forces=False)
        s.update_force()
        s.step(60, DT, thermostat=gth, barostat=gba)
        s.step(40, DT, thermostat=gth, barostat=gba)  # batches must compose like single calls
        s.download(st)
        m = s.macro()
    run_oracle(olj, o, 100, oth, oba)
    om = orc.macro(o)
    assert np.abs(st.boundary_box / o.box - 1.0).max() <= 1e-12
    dx = np.abs(st.position - o.pos)
    dx = np.minimum(dx, np.abs(dx - o.box))
    assert dx.max() <= 1e-8
    assert np.abs(st.velocity - o.vel).max() <= 1e-8
    assert abs(gth[0].lambda_ / oth.lambda_ - 1.0) <= 1e-12
    if oba is not None:
        assert abs(gba[0].myu / oba.myu - 1.0) <= 1e-12
    for key in ("kinetic", "thermal", "temperature", "pressure"):
        assert abs(m[key] - om[key]) <= 1e-9 * max(1.0, abs(om[key])), key
    assert abs(m["potential"] - om["potential"]) <= 1e-9 * max(1.0, abs(om["potential"]))


@pytest.mark.parametrize("name", ["liquid1000", "gas1000"])
def test_nose_hoover_trajectory(name, mode):  # thermostat.rs:35-39, 59-65
    o, cut = SYSTEMS[name]()
    olj, plj = lj_pair(md, *(cut or (None, None)))
    st = to_gpu_state(md, o)
    t0 = 300.0 if name == "gas1000" else 120.0
    oth, gth = orc.Thermostat(orc.Thermostat.NOSE_HOOVER, 0.5, t0), (md.Thermostat.NoseHoover(0.5), t0)
    with make_solver(mode) as s:
        s.set_potential(plj)
        s.upload(st, with_forces=False)
        s.update_force()
        s.step(1, DT, thermostat=gth)     # psi is carried by the caller between calls, like the enum field
        s.step(59, DT, thermostat=gth)
        s.step(40, DT, thermostat=gth)
        s.download(st)
    run_oracle(olj, o, 100, oth)
    assert abs(gth[0].psi - oth.psi) <= 1e-10 * max(1.0, abs(oth.psi))
    assert abs(gth[0].lambda_ / oth.lambda_ - 1.0) <= 1e-12
    assert np.abs(st.velocity - o.vel).max() <= 1e-8
    dx = np.abs(st.position - o.pos)
    assert np.minimum(dx, np.abs(dx - o.box)).max() <= 1e-8


@pytest.mark.parametrize("ensemble", ["nvt", "npt"])
def test_long_run_statistics_match_oracle(ensemble):
    """over long thermostat/barostat runs (>= 10^4 steps) trajectories diverge chaotically, so the gate is
    statistical — <T>, <P> and their spreads over the second half of a 10 000-step run must agree with the oracle's run from
    the same frame 0 within the sampling error; total momentum must not drift.  (512 atoms keep the CPU oracle's share short.)"""
    o = liquid(8)
    olj = orc.LennardJones()
    st = to_gpu_state(md, o)
    oth, gth = orc.Thermostat(orc.Thermostat.BERENDSEN, 1.0, 120.0), (md.Thermostat.Berendsen(1.0), 120.0)
    oba = gba = None
    if ensemble == "npt":
        oba, gba = orc.Barostat(1.0, 5.0, 1.01325), (md.Barostat.Berendsen(1.0, 5.0), 1.01325)
    n_steps, every = 10000, 20
    gt, gp, ot, op_ = [], [], [], []
    with md.Solver() as s:
        s.upload(st, with_forces=False)
        s.update_force()
        for k in range(n_steps // every):
            s.step(every, DT, thermostat=gth, barostat=gba)
            if k >= n_steps // every // 2:
                m = s.macro()
                gt.append(m["temperature"]); gp.append(m["pressure"])
        mom = np.abs(s.macro()["momentum"]).max()
    orc.update_force(olj, o, mode="cells")
    for k in range(n_steps // every):
        orc.step(olj, o, DT, thermostat=oth, barostat=oba, mode="cells", n_steps=every)
        if k >= n_steps // every // 2:
            m = orc.macro(o)
            ot.append(m["temperature"]); op_.append(m["pressure"])
    gt, gp, ot, op_ = map(np.array, (gt, gp, ot, op_))
    # 250 samples 20 steps apart: allow the mean to differ by 0.3 standard deviations of the samples (several standard
    # errors of the mean even for strongly correlated samples), the spreads by a factor 1.5
    assert abs(gt.mean() - ot.mean()) <= 0.3 * ot.std() + 1e-9, (gt.mean(), ot.mean(), ot.std())
    assert abs(gp.mean() - op_.mean()) <= 0.3 * op_.std() + 1e-9, (gp.mean(), op_.mean(), op_.std())
    assert 1 / 1.5 <= gt.std() / ot.std() <= 1.5 and 1 / 1.5 <= gp.std() / op_.std() <= 1.5
    assert abs(gt.mean() - 120.0) < 5.0  # Berendsen tau = 1 ps: the thermostat has long since reached the target
    assert mom < 1e-9


@pytest.mark.parametrize("cell", ["u", "fcc"])
def test_device_side_initializer(cell):
    """`initialize` on the device.  Positions and box are bit-identical to the reference's formulas
    (position.rs:24-104, oracle's restatement); velocities can only match in distribution (the reference's RNG is
    unseeded): second half = negated first half exactly, zero net momentum, variance sigma^2 = K_B*T/100/m."""
    side, lc, t_init, mass = (24, orc.GAS_CELL, 273.15, orc.ARGON_MASS) if cell == "u" else (12, 0.5256, 120.0, orc.ARGON_MASS)
    n = side ** 3 * (1 if cell == "u" else 4)
    with md.Solver() as s:
        s.initialize_lattice((side, side, side), lc, mass, t_init, cell=cell, seed=5)
        st = md.State(np.zeros((n, 3)), np.zeros((n, 3)), mass, np.ones(3))
        s.download(st)
        m = s.macro()
        s.update_force()
        s.step(10, DT, thermostat=(md.Thermostat.Berendsen(10.0), t_init))   # the State is usable right away
        assert s.stats()["steps"] == 10
    want = orc.argon_lattice(side, lc, t_init, seed=1, kind=cell).pos   # the oracle's restatement of position.rs:24-104
    assert np.array_equal(st.position, want)
    assert np.array_equal(st.boundary_box, np.array([lc * side] * 3))
    half = n // 2
    assert np.array_equal(st.velocity[half:2 * half], -st.velocity[:half])
    sigma2 = orc.K_B * (t_init * 0.01) / mass
    v = st.velocity[:half].ravel()
    assert abs(v.mean()) < 4.0 * np.sqrt(sigma2 / v.size)
    assert abs(v.var() / sigma2 - 1.0) < 5.0 * np.sqrt(2.0 / v.size)
    assert abs(np.mean(v ** 4) / (3.0 * sigma2 ** 2) - 1.0) < 0.1          # Gaussian kurtosis
    assert np.abs(m["momentum"]).max() < 1e-9
    assert abs(m["temperature"] / t_init - 1.0) < 0.05
    assert not st.force.any() and not st.potential.any()


def _npt_run(o, mode, t0, **kw):
    st = to_gpu_state(md, o)
    th, ba = (md.Thermostat.Berendsen(10.0), t0), (md.Barostat.Berendsen(1.0, 5.0), 1.01325)
    with make_solver(mode, **kw) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        for k in (7, 1, 30, 63, 200):
            s.step(k, DT, thermostat=th, barostat=ba)
        s.download(st)
        m = s.macro()
        return (st.position.copy(), st.velocity.copy(), st.force.copy(), st.potential.copy(), st.temp.copy(),
                st.boundary_box.copy(), np.array([m["temperature"], m["pressure"], th[0].lambda_, ba[0].myu]), s.stats())


def test_loop_drivers_agree(mode):
    """Dense systems: pre-enqueued graph chunks of the two-kernel step (default) and one host round trip per step run the
    same kernels on the same data — identical bits.  Dilute systems: the persistent step loop (default) and the same
    kernel launched once per step (host loop) — identical bits; the two-kernel chunk loop (chunk_loop=True) sums the K5
    terms in another order, so lambda/myu differ in the last bits and the trajectories agree to 1e-9 only."""
    dense = [_npt_run(liquid(10), mode, 120.0, **kw) for kw in ({}, {"host_loop": True})]
    for a, b in zip(dense[0][:7], dense[1][:7]):
        assert np.array_equal(a, b)
    assert dense[0][7]["rebuilds"] == dense[1][7]["rebuilds"]
    assert dense[0][7]["graph_launches"] > 0 and dense[1][7]["graph_launches"] == 0
    assert dense[0][7]["persistent_loop"] == 0

    # 1728 atoms on a jittered 1.7 nm lattice, hot: ~1.3 listed partners each (below 2 the persistent loop is the default
    # driver on one GPU), collisions and list rebuilds
    hot = orc.argon_lattice(12, 1.7, 900.0, 7)
    hot.pos += np.random.default_rng(5).uniform(-0.6, 0.6, hot.pos.shape)
    orc.apply_boundary_conditions(hot)
    dil = [_npt_run(hot, mode, 300.0, skin=0.3, **kw) for kw in ({}, {"host_loop": True}, {"chunk_loop": True})]
    for a, b in zip(dil[0][:7], dil[1][:7]):
        assert np.array_equal(a, b)
    assert dil[0][7]["persistent_loop"] == 1 and dil[0][7]["loop_steps"] > 250 and dil[0][7]["loop_launches"] < 60
    assert dil[1][7]["loop_launches"] == dil[1][7]["loop_steps"] > 250          # host loop: one launch per step
    assert dil[2][7]["loop_launches"] == 0 and dil[2][7]["graph_launches"] > 0  # two-kernel chunk loop
    assert dil[0][7]["rebuilds"] == dil[1][7]["rebuilds"] == dil[2][7]["rebuilds"] > 1
    assert np.abs(dil[0][2]).max() > 0.0   # pair terms were exercised
    dx = np.abs(dil[0][0] - dil[2][0])
    assert np.minimum(dx, np.abs(dx - dil[0][5])).max() <= 1e-9   # (an atom may sit on either side of the wrap)
    for a, b in zip(dil[0][1:6], dil[2][1:6]):
        assert np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(a).max())


def test_run_to_run_determinism():
    o = liquid(12)
    res = []
    for _ in range(2):
        st = to_gpu_state(md, o)
        with md.Solver() as s:
            s.upload(st, with_forces=False)
            s.update_force()
            s.step(200, DT, thermostat=(md.Thermostat.Berendsen(10.0), 120.0))
            s.download(st)
            res.append((st.position.copy(), st.velocity.copy(), st.force.copy(), s.macro()["pressure"]))
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][2], res[1][2])
    assert res[0][3] == res[1][3]


def test_rebuild_stress_small_skin(mode):
    """Skin of 0.02 nm on the liquid forces a list rebuild every few steps; results must not change."""
    o = liquid(10, temperature=300.0)
    olj, _ = lj_pair(md)
    st = to_gpu_state(md, o)
    with make_solver(mode, skin=0.02) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        s.step(100, DT)
        s.download(st)
        stats = s.stats()
    run_oracle(olj, o, 100)
    assert stats["rebuilds"] >= 10
    if mode == "exact":
        assert np.array_equal(st.position, o.pos) and np.array_equal(st.velocity, o.vel)
    else:
        assert np.abs(st.velocity - o.vel).max() <= 1e-8


def test_cell_subdivision_gives_same_pairs(mode):
    o = liquid(12)
    lists = []
    for sub in (1, 2):
        st = to_gpu_state(md, o)
        with make_solver(mode, cell_subdiv=sub, skin=0.1) as s:
            s.upload(st, with_forces=False)
            s.update_force()
            lists.append(s.neighbour_lists())
            s.download(st)
    assert np.array_equal(lists[0][0], lists[1][0]) and np.array_equal(lists[0][1], lists[1][1])


def test_per_call_api_matches_session(mode):
    o = liquid(8)
    a, b = to_gpu_state(md, o), to_gpu_state(md, o)
    th = lambda: (md.Thermostat.Berendsen(10.0), 120.0)  # noqa: E731
    with make_solver(mode) as s:
        s.upload(a, with_forces=False)
        s.update_force()
        s.step(5, DT, thermostat=th())
        s.download(a)
    with make_solver(mode) as s:
        s.update_force_host(b)
        t = th()
        for _ in range(5):
            s.calculate_host(b, DT, thermostat=t)
    if mode == "exact":
        # same pair sets and ascending-index sums → independent of when lists were built
        assert np.array_equal(a.position, b.position) and np.array_equal(a.velocity, b.velocity)
    else:
        assert np.abs(a.velocity - b.velocity).max() <= 1e-12


def test_macro_parameters_match_oracle(mode):
    o = dense_gas(3000)
    olj, plj = lj_pair(md)
    orc.update_force(olj, o)
    om = orc.macro(o)
    st = to_gpu_state(md, o)
    with make_solver(mode) as s:
        s.upload(st, with_forces=False)
        s.update_force()
        m = s.macro()
    for key in ("kinetic", "thermal", "potential", "temperature", "pressure"):
        assert abs(m[key] - om[key]) <= 1e-12 * max(1.0, abs(om[key])), key
    assert np.abs(m["vcom"] - om["vcom"]).max() <= 1e-15
    # shifted one-pass thermal sum stays accurate with a large centre-of-mass drift
    o.vel += np.array([50.0, -20.0, 10.0])
    om = orc.macro(o)
    st = to_gpu_state(md, o)
    with make_solver(mode) as s:
        s.upload(st)
        m = s.macro()
    assert abs(m["thermal"] / om["thermal"] - 1.0) <= 1e-12
    assert abs(m["kinetic"] / om["kinetic"] - 1.0) <= 1e-12


def test_errors():
    st = md.State(np.zeros((2, 3)), np.zeros((2, 3)), 1.0, [2.0, 2.0, 2.0])
    with md.Solver() as s:
        with pytest.raises(md.MdError) as e:
            s.step(1, DT)
        assert e.value.code == 6  # MD_ERR_NO_STATE
        s.upload(st)
        bad = md.Thermostat.Berendsen(1.0)
        bad.kind = 99  # Thermostat::Custom is todo!() in the reference
        with pytest.raises(md.MdError) as e:
            s.step(1, DT, thermostat=(bad, 300.0))
        assert e.value.code == 4  # MD_ERR_UNSUPPORTED
        with pytest.raises(md.MdError) as e:  # T = 0 → Berendsen lambda is not finite (no guard in the reference)
            s.step(1, DT, thermostat=(md.Thermostat.Berendsen(1.0), 300.0))
        assert e.value.code == 7
    with pytest.raises(md.MdError):
        md.Integrator.Custom("x").calculate(md.PotentialsDatabase(), st, DT)
    with pytest.raises(md.MdError):
        md.State(np.zeros((0, 3)), np.zeros((0, 3)), 1.0, [1, 1, 1])


# ---------------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs): size-independent checks + sampled rows against the oracle, in BOTH force
# modes — MD_FORCE_FAST is the configuration bench.py times
def _perturbed_gas(n_side, seed=1):
    """The gas lattice moved off its zero-force symmetry point (uniform displacements of up to 1.45 nm)."""
    o = gas(n_side)
    o.pos += np.random.default_rng(seed).uniform(-1.45, 1.45, o.pos.shape)
    orc.apply_boundary_conditions(o)
    return o


def _row_scale(olj, o, rows):
    """Σ_j |f_ij| of the given atoms (periodic KD-tree): the scale the 1e-10 force bar is relative to."""
    from scipy.spatial import cKDTree
    pos = np.mod(o.pos, o.box)
    tree = cKDTree(pos, boxsize=o.box)
    out = np.zeros(len(rows))
    for k, i in enumerate(rows):
        js = np.array([j for j in tree.query_ball_point(pos[i], olj.r_cut * (1 + 1e-12)) if j != i], dtype=np.int64)
        if len(js) == 0:
            continue
        d = pos[js] - pos[i]
        d -= o.box * np.round(d / o.box)
        r = np.sqrt((d * d).sum(axis=1))
        r = r[r <= olj.r_cut]
        s6 = (olj.sigma / r) ** 6
        out[k] = np.abs(24.0 * olj.eps / r * (s6 - 2.0 * s6 * s6)).sum()
    return out


@pytest.mark.parametrize("n_side,cut", [(100, None), (64, LONG_CUT)])
def test_full_size_sampled_rows_and_invariants(n_side, cut):
    o = _perturbed_gas(n_side) if cut is None else liquid(n_side, jitter=0.03)
    olj, plj = lj_pair(md, *(cut or (None, None)))
    starts = (0, o.n // 2 - 128, o.n - 256)
    rows = np.concatenate([np.arange(i0, i0 + 256) for i0 in starts])
    for i0 in starts:  # the reference's Θ(N) row scan for the sampled atoms
        orc.update_force(olj, o, rows=(i0, i0 + 256))
    scale = _row_scale(olj, o, rows)
    for exact in (True, False):
        st = to_gpu_state(md, o)
        with md.Solver(exact=exact) as s:
            s.set_potential(plj)
            s.upload(st, with_forces=False)
            s.update_force()
            s.download(st)
            m0 = s.macro()
            if exact:  # bit-exact
                assert np.array_equal(st.force[rows], o.force[rows])
                assert np.array_equal(st.potential[rows], o.pot[rows])
                assert np.array_equal(st.temp[rows], o.vir[rows])
            else:      # |Δ| <= 1e-10 * max(Σ_j|f_ij|, global RMS)
                rms = np.sqrt((st.force ** 2).sum(axis=1).mean())
                bar = 1e-10 * np.maximum(scale, rms)
                assert np.all(np.abs(st.force[rows] - o.force[rows]) <= bar[:, None])
                assert np.all(np.abs(st.potential[rows] - o.pot[rows]) <= 1e-10 * (np.abs(o.pot[rows]) + 4 * olj.eps))
                assert np.all(np.abs(st.temp[rows] - o.vir[rows]) <= 1e-10 * (scale * olj.r_cut) + 1e-300)
            # Newton's third law holds pairwise → net force is rounding noise
            assert np.abs(st.force.sum(axis=0)).max() <= 1e-9 * np.abs(st.force).sum()
            s.step(50, DT)
            m1 = s.macro()
            stats = s.stats()
        assert stats["steps"] == 50
        e0, e1 = m0["kinetic"] + m0["potential"], m1["kinetic"] + m1["potential"]
        assert abs(e1 - e0) <= 1e-3 * abs(m0["kinetic"])       # NVE energy conservation at dt = 2 fs (oracle: 1.1e-4)
        assert np.abs(m1["momentum"]).max() <= 1e-9 * o.n       # Σ m v stays ~0


@pytest.mark.parametrize("config", ["c2", "c5"])
def test_config_size_trajectory_100_steps(config):
    """BASELINE configs C2 (argon gas 32^3, NVT Berendsen) and C5 (liquid 64^3 = 262 144 atoms, 3.5 sigma cutoff, NVT) in the
    benchmarked setup — MD_FORCE_FAST, default loop driver — against the oracle's Θ(N) cell-list step from the same frame 0,
    100 steps: |Δx|, |Δv| <= 1e-8, lambda within 1e-12, macro parameters within 1e-9."""
    if config == "c2":
        o, cut, t0 = _perturbed_gas(32, seed=2), None, 300.0
    else:
        o, cut, t0 = liquid(64, jitter=0.03), LONG_CUT, 120.0
    olj, plj = lj_pair(md, *(cut or (None, None)))
    st = to_gpu_state(md, o)
    gth, oth = (md.Thermostat.Berendsen(10.0), t0), orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, t0)
    with md.Solver() as s:
        s.set_potential(plj)
        s.upload(st, with_forces=False)
        s.update_force()
        s.step(100, DT, thermostat=gth)
        s.download(st)
        m = s.macro()
        stats = s.stats()
    run_oracle(olj, o, 100, oth)
    om = orc.macro(o)
    dx = np.abs(st.position - o.pos)
    assert np.minimum(dx, np.abs(dx - o.box)).max() <= 1e-8
    assert np.abs(st.velocity - o.vel).max() <= 1e-8
    assert abs(gth[0].lambda_ / oth.lambda_ - 1.0) <= 1e-12
    for key in ("kinetic", "thermal", "potential", "temperature", "pressure"):
        assert abs(m[key] - om[key]) <= 1e-9 * max(1.0, abs(om[key])), key
    assert stats["steps"] == 100
    assert stats["persistent_loop"] == (1 if config == "c2" else 0)
    if config == "c5":
        assert stats["rebuilds"] >= 2 and stats["nbr_mean"] > 150   # the liquid rebuilds its lists within 100 steps


def test_c4_size_npt_rows_and_box():
    """BASELINE config C4 (argon 216^3 = 10 077 696 atoms, NPT Berendsen thermostat + barostat) on one GPU, MD_FORCE_FAST:
    sampled force rows within the 1e-10 bar, then 50 NPT steps against the oracle — the box (scaled by myu every step,
    barostat.rs:39-49) within 1e-12 relative, positions and velocities within 1e-8, lambda and myu within 1e-12."""
    o = _perturbed_gas(216, seed=3)
    olj, plj = lj_pair(md)
    starts = (0, o.n // 2 - 128, o.n - 256)
    rows = np.concatenate([np.arange(i0, i0 + 256) for i0 in starts])
    for i0 in starts:
        orc.update_force(olj, o, rows=(i0, i0 + 256))
    scale = _row_scale(olj, o, rows)
    ref_f, ref_u, ref_w = o.force[rows].copy(), o.pot[rows].copy(), o.vir[rows].copy()
    st = to_gpu_state(md, o)
    gth, oth = (md.Thermostat.Berendsen(10.0), 300.0), orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, 300.0)
    gba, oba = (md.Barostat.Berendsen(1.0, 5.0), 1.01325), orc.Barostat(1.0, 5.0, 1.01325)
    with md.Solver() as s:
        s.upload(st, with_forces=False)
        s.update_force()
        s.download(st)
        rms = np.sqrt((st.force ** 2).sum(axis=1).mean())
        assert np.all(np.abs(st.force[rows] - ref_f) <= 1e-10 * np.maximum(scale, rms)[:, None])
        assert np.all(np.abs(st.potential[rows] - ref_u) <= 1e-10 * (np.abs(ref_u) + 4 * olj.eps))
        assert np.all(np.abs(st.temp[rows] - ref_w) <= 1e-10 * (scale * olj.r_cut) + 1e-300)
        s.step(20, DT, thermostat=gth, barostat=gba)
        s.step(30, DT, thermostat=gth, barostat=gba)
        s.download(st)
        m = s.macro()
    run_oracle(olj, o, 50, oth, oba)
    assert np.abs(st.boundary_box / o.box - 1.0).max() <= 1e-12
    assert abs(gba[0].myu / oba.myu - 1.0) <= 1e-12 and abs(gth[0].lambda_ / oth.lambda_ - 1.0) <= 1e-12
    dx = np.abs(st.position - o.pos)
    assert np.minimum(dx, np.abs(dx - o.box)).max() <= 1e-8
    assert np.abs(st.velocity - o.vel).max() <= 1e-8
    om = orc.macro(o)
    for key in ("temperature", "pressure"):
        assert abs(m[key] - om[key]) <= 1e-9 * max(1.0, abs(om[key])), key


# ---------------------------------------------------------------------------------------------------
# the reference's two #[ignore]d long-run tests, through the GPU path for the full 100 000 steps
# The reference's assertions (|T - 273.15| < 1e-5, |P - 0.101325| < 1e-5 after 100 000 steps of 8 gas atoms) hold exactly when
# no collision falls into the last ~3000 steps (Berendsen relaxes geometrically, 1/250 per step) — with the reference's
# unseeded velocities that is a coin flip (oracle, seeds 1..12: the thermostat criterion holds for 6, the barostat's for 5),
# which is why both tests are #[ignore]d there.  Here the 8 atoms get a collision-free draw: the two atoms of every x-row
# share one velocity along x and the rows (3.34 nm apart, r_cut 0.85) never approach — forces stay exactly zero, nothing is
# chaotic, and GPU and oracle can also be compared with each other after all 100 000 steps.
def _eight_atoms(kat):
    o = orc.argon_lattice(kat["grid"][0], kat["cell"], 273.15, 42)
    a, b = 0.31, 0.23
    o.vel[:] = 0.0
    o.vel[:, 0] = np.array([a, -a, b, -b])[np.arange(o.n) % 4]   # index = x*4 + y*2 + z: atoms i and i+4 form a row
    return o


def test_reference_berendsen_thermostat_100000_steps(kats):  # solver/src/lib.rs:429-454 (#[ignore] there)
    k = kats["berendsen_thermostat"]
    o = _eight_atoms(k)
    olj = orc.LennardJones()
    st = to_gpu_state(md, o)
    gth, oth = (md.Thermostat.Berendsen(k["tau"]), k["target"]), orc.Thermostat(orc.Thermostat.BERENDSEN, k["tau"], k["target"])
    t_start = orc.macro(o)["temperature"]
    assert abs(t_start - k["target"]) > 10.0                               # the thermostat has work to do
    with md.Solver() as s:
        s.upload(st, with_forces=False)
        s.update_force()
        s.step(k["n_steps_reference"], DT, thermostat=gth)
        s.download(st)
        m = s.macro()
        stats = s.stats()
    assert stats["steps"] == k["n_steps_reference"]
    assert abs(m["temperature"] - k["target"]) < k["tolerance"]          # the reference's own assertion
    run_oracle(olj, o, k["n_steps_reference"], oth)
    assert abs(orc.macro(o)["temperature"] - k["target"]) < k["tolerance"]
    assert abs(m["temperature"] - orc.macro(o)["temperature"]) < 1e-9
    assert np.abs(st.velocity - o.vel).max() < 1e-9
    assert not st.force.any()


def test_reference_berendsen_barostat_100000_steps(kats):  # solver/src/lib.rs:456-482 (#[ignore] there)
    k = kats["berendsen_barostat"]
    o = _eight_atoms(k)
    olj = orc.LennardJones()
    st = to_gpu_state(md, o)
    gba, oba = (md.Barostat.Berendsen(k["beta"], k["tau"]), k["target"]), orc.Barostat(k["beta"], k["tau"], k["target"])
    with md.Solver() as s:
        s.upload(st, with_forces=False)
        s.update_force()
        s.step(k["n_steps_reference"], DT, barostat=gba)
        s.update_force()                                                   # lib.rs:476: forces at the final box
        m = s.macro()
    assert abs(m["pressure"] - k["target"]) < k["tolerance"]              # the reference's own assertion
    run_oracle(olj, o, k["n_steps_reference"], None, oba)
    orc.update_force(olj, o, mode="cells")
    assert abs(orc.macro(o)["pressure"] - k["target"]) < k["tolerance"]
    assert abs(m["pressure"] - orc.macro(o)["pressure"]) < 1e-9
    assert np.abs(m["box"] / o.box - 1.0).max() < 1e-9
    assert abs(m["box"][0] / (k["grid"][0] * k["cell"]) - 1.0) > 1e-3     # the box did move
