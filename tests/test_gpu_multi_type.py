"""GPU parity for States with several particle types: the CUDA path through the C ABI against the oracle's
restatement of what the reference does with T > 1 types (pinned on CPU in test_multi_type_cpu.py).

  * MD_FORCE_EXACT: per-atom F / U / W and NVE trajectories BIT-IDENTICAL to the oracle, in the reference's one-sided
    cross-type accumulation (potential.rs:168-176) and in the symmetric variant;
  * MD_FORCE_FAST: 1e-10 relative to the atom's Σ|f_ij| scale; trajectories 1e-8 over 100 steps;
  * thermostat + barostat: the last type's lambda / myu for every type, the box scaled once per type (integrator.rs:18-27,
    54-58), per-type macro parameters.
"""
import numpy as np
import pytest

import moldyn_b200 as md
from oracle import oracle as orc

from test_multi_type_cpu import mixture, table3

pytestmark = pytest.mark.gpu

DT = 0.002


@pytest.fixture(scope="module", params=["exact", "fast"])
def mode(request):
    return request.param


def big_mixture(seed=5):
    """648 atoms, three types interleaved in space, liquid-like density: ~40 partners inside the largest cutoff."""
    return mixture(n_side=9, cell=0.42, counts=(300, 280, 149), masses=(66.335, 20.18, 131.29), temperature=140.0, seed=seed,
                   jitter=0.04)


def to_md(st):
    m = md.MultiState(st.pos, st.vel, st.counts, st.masses, st.box)
    m.force[:] = st.force
    m.potential[:] = st.pot
    m.temp[:] = st.vir
    return m


def set_table(s, tab):
    for (a, b), lj in tab.entries.items():
        s.set_potential_pair(a, b, md.Potential(lj.sigma, lj.eps, lj.r_cut, lj.u_cut))


def scale_of(st):
    """Σ_j |f_ij| per atom with the largest potential of the table: the scale FAST-mode errors are measured against."""
    from helpers import force_scale
    one = orc.State(st.pos, st.vel, 1.0, st.box)
    return force_scale(orc.LennardJones(0.37, 2.0), one) + 1.0


@pytest.mark.parametrize("symmetric", [False, True])
def test_update_force(mode, symmetric):
    tab = table3()
    st = big_mixture()
    orc.update_force_multi(tab, st, symmetric=symmetric)
    g = to_md(st)
    with md.Solver(exact=(mode == "exact")) as s:
        set_table(s, tab)
        s.set_cross_type_mode(symmetric)
        s.upload_typed(g, with_forces=False)
        s.update_force()
        s.download(g)
        st_stats = s.stats()
    assert st_stats["nbr_mean"] > 20
    if mode == "exact":
        assert np.array_equal(g.force, st.force) and np.array_equal(g.potential, st.pot) and np.array_equal(g.temp, st.vir)
    else:
        sc = scale_of(st)
        assert (np.abs(g.force - st.force).max(axis=1) <= 1e-10 * sc).all()
        assert (np.abs(g.potential - st.pot) <= 1e-10 * sc).all() and (np.abs(g.temp - st.vir) <= 1e-10 * sc).all()
    if not symmetric:
        # the reference's quirk is visible: momentum is not conserved by the one-sided accumulation
        assert np.abs(st.force.sum(axis=0)).max() > 1.0


@pytest.mark.parametrize("symmetric", [False, True])
def test_nve_trajectory_100_steps(mode, symmetric):
    tab = table3()
    st = big_mixture(seed=8)
    orc.update_force_multi(tab, st, symmetric=symmetric)
    g = to_md(st)
    with md.Solver(exact=(mode == "exact"), skin=0.05) as s:   # narrow skin: several list rebuilds inside the run
        set_table(s, tab)
        s.set_cross_type_mode(symmetric)
        s.upload_typed(g)
        for k in (1, 37, 62):
            s.step(k, DT)
        s.download(g)
        stats = s.stats()
    orc.step_multi(tab, st, DT, symmetric=symmetric, n_steps=100)
    assert stats["rebuilds"] >= 3
    if mode == "exact":
        assert np.array_equal(g.position, st.pos) and np.array_equal(g.velocity, st.vel)
        assert np.array_equal(g.force, st.force) and np.array_equal(g.temp, st.vir)
    else:
        assert np.abs(g.position - st.pos).max() <= 1e-8 and np.abs(g.velocity - st.vel).max() <= 1e-8


@pytest.mark.parametrize("symmetric", [False, True])
def test_npt_trajectory_and_per_type_macro(mode, symmetric):
    tab = table3()
    st = big_mixture(seed=11)
    orc.update_force_multi(tab, st, symmetric=symmetric)
    g = to_md(st)
    th_o, ba_o = orc.Thermostat(1, 0.5, 200.0), orc.Barostat(1.0e-3, 2.0, 1.0)
    th, ba = (md.Thermostat.Berendsen(0.5), 200.0), (md.Barostat.Berendsen(1.0e-3, 2.0), 1.0)
    with md.Solver(exact=(mode == "exact")) as s:
        set_table(s, tab)
        s.set_cross_type_mode(symmetric)
        s.upload_typed(g)
        box0 = st.box.copy()
        s.step(1, DT, thermostat=th, barostat=ba)
        orc.step_multi(tab, st, DT, th_o, ba_o, symmetric=symmetric)
        # one step: lambda and myu are those of the LAST type, the box is scaled once per type
        assert abs(th[0].lambda_ - th_o.lambda_) <= 1e-13 and abs(ba[0].myu - ba_o.myu) <= 1e-13
        s.download(g)
        assert np.abs(g.boundary_box / (box0 * ba_o.myu ** 3) - 1.0).max() <= 1e-13
        s.step(49, DT, thermostat=th, barostat=ba)
        orc.step_multi(tab, st, DT, th_o, ba_o, symmetric=symmetric, n_steps=49)
        s.download(g)
        assert np.abs(g.boundary_box / st.box - 1.0).max() <= 1e-11
        assert np.abs(g.position - st.pos).max() <= 1e-8 and np.abs(g.velocity - st.vel).max() <= 1e-8
        assert abs(th[0].lambda_ - th_o.lambda_) <= 1e-11 and abs(ba[0].myu - ba_o.myu) <= 1e-11
        for t in range(3):
            a, b = s.macro_type(t), orc.macro_type(st, t)
            for key in ("kinetic", "thermal", "potential", "temperature", "pressure"):
                assert abs(a[key] - b[key]) <= 1e-8 * max(1.0, abs(b[key])), (t, key, a[key], b[key])
            assert np.abs(a["vcom"] - b["vcom"]).max() <= 1e-10 and a["n"] == st.counts[t]
        with pytest.raises(md.MdError) as e:
            s.macro()
        assert e.value.code == 1


def test_boundaries_of_the_multi_type_path():
    tab = table3()
    st = mixture()
    g = to_md(st)
    with md.Solver(exact=True) as s:
        set_table(s, tab)
        s.upload_typed(g, with_forces=False)
        with pytest.raises(md.MdError) as e:   # thermostat.rs:59-65 with several types is not offered
            s.step(1, DT, thermostat=(md.Thermostat.NoseHoover(0.5), 200.0))
        assert e.value.code == 4
        with pytest.raises(md.MdError) as e:
            s.set_potential_pair(0, 8, md.Potential(0.3, 1.0, 0.75, 0.0))
        assert e.value.code == 1
    with pytest.raises(md.MdError):            # an empty type: the reference panics on particle_type[0]
        md.MultiState(st.pos, st.vel, [st.n, 0], [1.0, 1.0], st.box)
    # one type through the typed upload is the single-type path
    one = orc.argon_lattice(6, 0.5, 120.0, 3)
    lj = orc.LennardJones()
    orc.update_force(lj, one)
    g1 = md.MultiState(one.pos, one.vel, [one.n], [one.mass], one.box)
    with md.Solver(exact=True) as s:
        s.upload_typed(g1, with_forces=False)
        s.update_force()
        s.download(g1)
        assert s.macro()["n"] == one.n
    assert np.array_equal(g1.force, one.force)


def test_changing_pair_00_after_upload_changes_the_list_radius():
    """md_set_potential_lj is pair (0, 0) of the type table: a longer cutoff set after the upload must widen the lists."""
    tab = table3()
    st = big_mixture(seed=13)
    g = to_md(st)
    long00 = orc.LennardJones(0.3418, 1.712, r_cut=1.2)
    with md.Solver(exact=True) as s:
        set_table(s, tab)
        s.upload_typed(g, with_forces=False)
        s.update_force()
        s.set_potential(md.Potential(long00.sigma, long00.eps, long00.r_cut, long00.u_cut))
        s.update_force()
        s.download(g)
        assert s.stats()["persistent_loop"] == 0
    tab.set_potential(0, 0, long00)
    orc.update_force_multi(tab, st)
    assert np.array_equal(g.force, st.force) and np.array_equal(g.potential, st.pot)
