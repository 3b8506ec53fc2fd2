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
    return md.Solver(exact=(mode == "exact"), union_lists=(mode == "fast_union"), **kw)


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
    "liquid5832": lambda: (liquid(18), None),        # 13 cells per axis: the dense FAST path uses union lists
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
        assert s.stats()["union_lists"] == 0
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


def test_union_lists_opt_in():
    """MD_FORCE_FAST_UNION (k_build_union + the union force loop): pair sets — reconstructed from the membership bits —
    bit-exact, forces within the FAST bar, 100-step NVT trajectory within 1e-8 of the oracle."""
    o, _ = SYSTEMS["liquid5832"]()
    olj, plj = lj_pair(md)
    ref = o.copy()
    orc.update_force(olj, ref, mode="cells")
    st = to_gpu_state(md, o)
    gth, oth = (md.Thermostat.Berendsen(10.0), 120.0), orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, 120.0)
    with make_solver("fast_union") as s:
        s.upload(st, with_forces=False)
        s.update_force()
        assert s.stats()["union_lists"] == 1
        off, partners = s.neighbour_lists()
        skin = s.stats()["skin"]
        s.download(st)
        f0 = st.force.copy()
        s.step(100, DT, thermostat=gth)
        s.download(st)
    woff, wpartners = orc.neighbour_sets(o.pos, o.box, olj.r_cut + skin)
    assert np.array_equal(off, woff) and np.array_equal(partners, wpartners)
    scale = force_scale(olj, ref)
    rms = np.sqrt((ref.force ** 2).sum(axis=1).mean())
    assert np.all(np.abs(f0 - ref.force) <= 1e-10 * np.maximum(scale, rms)[:, None])
    run_oracle(olj, o, 100, oth)
    dx = np.abs(st.position - o.pos)
    assert np.minimum(dx, np.abs(dx - o.box)).max() <= 1e-8
    assert np.abs(st.velocity - o.vel).max() <= 1e-8


@pytest.mark.parametrize("name", ["liquid4096_3.5sigma", "liquid5832", "dense_gas3000", "liquid1000", "liquid729"])
def test_warp_cooperative_dense_kernel(name):
    """MD_FORCE_FAST_COOP (k_transpose_list + k_force_coop: one warp per atom over an atom-major list): per-atom force,
    potential and virial within the FAST bar, 100-step NVT and NPT trajectories within 1e-8 of the oracle (forces-only,
    + virial and + potential instances of the loop all run), run-to-run bit-reproducible."""
    o, cut = SYSTEMS[name]()
    olj, plj = lj_pair(md, *(cut or (None, None)))
    ref = o.copy()
    orc.update_force(olj, ref, mode="cells")
    scale = force_scale(olj, ref)
    rms = np.sqrt((ref.force ** 2).sum(axis=1).mean())
    T0 = 120.0 if name.startswith("liquid") else 273.15
    runs = []
    for ensemble in ("nvt", "npt", "npt"):
        st = to_gpu_state(md, o)
        gth, oth = (md.Thermostat.Berendsen(10.0), T0), orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, T0)
        gba = oba = None
        if ensemble == "npt":
            gba, oba = (md.Barostat.Berendsen(1.0, 5.0), 1.01325), orc.Barostat(1.0, 5.0, 1.01325)
        with md.Solver(coop=True) as s:
            s.set_potential(plj)
            s.upload(st, with_forces=False)
            s.update_force()
            assert s.stats()["coop_lists"] == 1 and s.stats()["nbr_mean"] >= 8.0
            s.download(st)
            assert np.all(np.abs(st.force - ref.force) <= 1e-10 * np.maximum(scale, rms)[:, None])
            assert np.all(np.abs(st.potential - ref.pot) <= 1e-10 * (np.abs(ref.pot) + 4 * olj.eps))
            assert np.all(np.abs(st.temp - ref.vir) <= 1e-10 * (scale * olj.r_cut + 1e-300) + 1e-300)
            for k in (1, 36, 63):
                s.step(k, DT, thermostat=gth, barostat=gba)
            s.download(st)
            m = s.macro()
        r = o.copy()
        run_oracle(olj, r, 100, oth, oba)
        dx = np.abs(st.position - r.pos)
        assert np.minimum(dx, np.abs(dx - r.box)).max() <= 1e-8
        assert np.abs(st.velocity - r.vel).max() <= 1e-8
        assert np.abs(st.boundary_box - r.box).max() <= 1e-9
        runs.append((st.position.copy(), st.velocity.copy(), st.force.copy(), st.temp.copy(), m["pressure"]))
    for a, b in zip(runs[1][:4], runs[2][:4]):
        assert np.array_equal(a, b)
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
    """over long thermostat/barostat runs trajectories diverge chaotically, so the gate is statistical —
    <T>, <P> and their spreads over the second half of a 3000-step run must agree with the oracle's run from the same
    frame 0 within the sampling error; total momentum must not drift.  (512 atoms keep the CPU oracle's share short.)"""
    o = liquid(8)
    olj = orc.LennardJones()
    st = to_gpu_state(md, o)
    oth, gth = orc.Thermostat(orc.Thermostat.BERENDSEN, 1.0, 120.0), (md.Thermostat.Berendsen(1.0), 120.0)
    oba = gba = None
    if ensemble == "npt":
        oba, gba = orc.Barostat(1.0, 5.0, 1.01325), (md.Barostat.Berendsen(1.0, 5.0), 1.01325)
    n_steps, every = 3000, 20
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
    # 75 samples 20 steps apart: allow the mean to differ by half a standard deviation of the samples (several standard
    # errors of the mean even for strongly correlated samples), the spreads by a factor 2
    assert abs(gt.mean() - ot.mean()) <= 0.5 * ot.std() + 1e-9, (gt.mean(), ot.mean(), ot.std())
    assert abs(gp.mean() - op_.mean()) <= 0.5 * op_.std() + 1e-9, (gp.mean(), op_.mean(), op_.std())
    assert 0.5 <= gt.std() / ot.std() <= 2.0 and 0.5 <= gp.std() / op_.std() <= 2.0
    assert abs(gt.mean() - 120.0) < 15.0  # the melting lattice is still being cooled towards the target
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


def test_graph_loop_equals_host_loop(mode):
    """The three loop drivers — chunked graphs of guarded steps (default), the conditional WHILE graph, and one host round
    trip per step — run the same kernels on the same data: identical bits."""
    o = liquid(10)
    out = []
    for kw in ({}, {"while_loop": True}, {"host_loop": True}):
        st = to_gpu_state(md, o)
        with make_solver(mode, **kw) as s:
            s.upload(st, with_forces=False)
            s.update_force()
            s.step(150, DT, thermostat=(md.Thermostat.Berendsen(10.0), 120.0),
                   barostat=(md.Barostat.Berendsen(1.0, 5.0), 1.01325))
            s.download(st)
            out.append((st.position.copy(), st.velocity.copy(), st.force.copy(), st.boundary_box.copy(), s.stats()))
    for other in out[1:]:
        for a, b in zip(out[0][:4], other[:4]):
            assert np.array_equal(a, b)
        assert out[0][4]["rebuilds"] == other[4]["rebuilds"]
    assert out[0][4]["graph_launches"] > 0 and out[1][4]["graph_launches"] > 0 and out[2][4]["graph_launches"] == 0


@pytest.mark.parametrize("host_loop", [False, True])
def test_fused_step_equals_split_step(mode, host_loop):
    """Dilute systems step with ONE fused kernel (k_step_dilute: partners are drifted on the fly, x/v ping-pong between
    two plane sets); it must reproduce the k_kick_drift + k_force path bit for bit, for any batch length/parity."""
    o = orc.argon_lattice(12, 1.0, 900.0, 7)   # 1728 atoms, ~6 listed partners each, hot: collisions and list rebuilds
    out = []
    for step_mode in ("split", "fused"):
        st = to_gpu_state(md, o)
        th = (md.Thermostat.Berendsen(10.0), 300.0)
        ba = (md.Barostat.Berendsen(1.0, 5.0), 1.01325)
        with make_solver(mode, host_loop=host_loop, step_mode=step_mode, skin=0.3) as s:
            s.upload(st, with_forces=False)
            s.update_force()
            for k in (7, 1, 30, 63, 200):
                s.step(k, DT, thermostat=th, barostat=ba)
            s.download(st)
            m = s.macro()
            out.append((st.position.copy(), st.velocity.copy(), st.force.copy(), st.potential.copy(), st.temp.copy(),
                        st.boundary_box.copy(), np.array([m["temperature"], m["pressure"], th[0].lambda_, ba[0].myu]),
                        s.stats()))
    for a, b in zip(out[0][:7], out[1][:7]):
        assert np.array_equal(a, b)
    assert out[0][7]["fused_steps"] == 0 and out[1][7]["fused_steps"] > 250
    assert out[0][7]["rebuilds"] == out[1][7]["rebuilds"] > 1
    assert np.abs(out[1][2]).max() > 0.0   # pair terms were exercised


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
# full-size properties (BASELINE configs): size-independent checks + sampled rows against the oracle
@pytest.mark.parametrize("n_side,cut", [(100, None), (64, LONG_CUT)])
def test_full_size_sampled_rows_and_invariants(n_side, cut):
    if cut is None:
        o = gas(n_side)
        # move the lattice off its zero-force symmetry point
        rng = np.random.default_rng(1)
        o.pos += rng.uniform(-1.45, 1.45, o.pos.shape)
        orc.apply_boundary_conditions(o)
    else:
        o = liquid(n_side, jitter=0.03)
    olj, plj = lj_pair(md, *(cut or (None, None)))
    st = to_gpu_state(md, o)
    with md.Solver(exact=True) as s:
        s.set_potential(plj)
        s.upload(st, with_forces=False)
        s.update_force()
        s.download(st)
        m0 = s.macro()
        # sampled rows, bit-exact against the reference's Θ(N) row scan
        for i0 in (0, o.n // 2 - 128, o.n - 256):
            orc.update_force(olj, o, rows=(i0, i0 + 256))
            assert np.array_equal(st.force[i0:i0 + 256], o.force[i0:i0 + 256])
            assert np.array_equal(st.potential[i0:i0 + 256], o.pot[i0:i0 + 256])
            assert np.array_equal(st.temp[i0:i0 + 256], o.vir[i0:i0 + 256])
        # Newton's third law holds pairwise → net force is rounding noise
        assert np.abs(st.force.sum(axis=0)).max() <= 1e-9 * np.abs(st.force).sum()
        s.step(50, DT)
        m1 = s.macro()
    e0, e1 = m0["kinetic"] + m0["potential"], m1["kinetic"] + m1["potential"]
    assert abs(e1 - e0) <= 1e-3 * abs(m0["kinetic"])       # NVE energy conservation at dt = 2 fs (oracle: 1.1e-4)
    assert np.abs(m1["momentum"]).max() <= 1e-9 * o.n       # Σ m v stays ~0


# ---------------------------------------------------------------------------------------------------
# programmatic dependent launch of the step kernels (MOLDYN_B200_PDL, read once per process → a child process each)
_PDL_CHILD = r"""
import sys, hashlib
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import moldyn_b200 as md
from helpers import liquid, gas, to_gpu_state
for name, o, T in (("liquid", liquid(12), 120.0), ("gas", gas(16), 300.0)):
    for exact in (False, True):
        st = to_gpu_state(md, o)
        with md.Solver(exact=exact) as s:
            s.upload(st, with_forces=False)
            s.update_force()
            for k in (1, 37, 200):
                s.step(k, 0.002, thermostat=(md.Thermostat.Berendsen(10.0), T),
                       barostat=(md.Barostat.Berendsen(1.0, 5.0), 1.01325))
            s.download(st)
            stats = s.stats()
        h = hashlib.sha256()
        for a in (st.position, st.velocity, st.force, st.potential, st.temp, st.boundary_box):
            h.update(np.ascontiguousarray(a).tobytes())
        print(name, exact, h.hexdigest(), stats["graph_launches"] > 0, stats["rebuilds"])
"""


def test_programmatic_dependent_launch_is_bit_identical():
    """MOLDYN_B200_PDL=1 turns the kernel → kernel edges of the step graph into programmatic ones (the kernels wait with
    griddepcontrol.wait before reading anything; k_kick_drift fetches its positions ahead of the wait).  Scheduling only:
    NPT trajectories through the graph loop must not change by a bit."""
    import os
    import subprocess
    import sys
    tests = os.path.dirname(os.path.abspath(__file__))
    code = _PDL_CHILD.format(root=os.path.dirname(tests), tests=tests)
    outs = []
    for pdl in ("0", "1", "2"):  # 2 = + early-start k_kick_drift (ticket / sequence-number protocol)
        env = dict(os.environ, MOLDYN_B200_PDL=pdl)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines())
    assert len(outs[0]) == 4 and outs[0] == outs[1] == outs[2]
    assert all(line.split()[3] == "True" for line in outs[0])  # the graph loop ran
