#!/usr/bin/env python
"""Writes tests/golden/oracle_fixtures.npz: many-body outputs of the CPU oracle on small seeded systems.

The reference's own tests pin only two-body cases (tests/golden/reference_kats.json); many-body forces and
thermostat/barostat trajectories are pinned by the restated oracle alone.  These vectors freeze the oracle's
answers (gcc -O2 -ffp-contract=off, no reassociation) so that a change of compiler, flags or source that moves a single bit
is caught on CPU by tests/test_oracle_golden.py::test_oracle_many_body_fixtures.

    python tests/golden/make_oracle_fixtures.py        # regenerate (only after a deliberate oracle change)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle as orc  # noqa: E402
from helpers import LONG_CUT, gas, liquid  # noqa: E402


def build():
    out = {}
    # 1. liquid, default cutoff: per-atom force / potential / virial (reference's Θ(N²) scan)
    o = liquid(6)
    lj = orc.LennardJones()
    orc.update_force(lj, o, mode="n2")
    out["liquid216_pos"], out["liquid216_box"] = o.pos.copy(), o.box.copy()
    out["liquid216_force"], out["liquid216_pot"], out["liquid216_vir"] = o.force.copy(), o.pot.copy(), o.vir.copy()
    # 2. the same system, 3.5 sigma cutoff (potentials.json entry of config C5)
    o2 = liquid(6)
    lj2 = orc.LennardJones(r_cut=LONG_CUT[0], u_cut=LONG_CUT[1])
    orc.update_force(lj2, o2, mode="n2")
    out["liquid216_long_force"], out["liquid216_long_vir"] = o2.force.copy(), o2.vir.copy()
    # 3. 50 NPT steps (Berendsen thermostat + barostat, README parameters) of the liquid
    o3 = liquid(6)
    orc.update_force(lj, o3, mode="n2")
    th = orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, 120.0)
    ba = orc.Barostat(1.0, 5.0, 1.01325)
    orc.step(lj, o3, 0.002, thermostat=th, barostat=ba, mode="n2", n_steps=50)
    out["liquid216_npt50_pos"], out["liquid216_npt50_vel"], out["liquid216_npt50_box"] = o3.pos.copy(), o3.vel.copy(), o3.box.copy()
    out["liquid216_npt50_lambda_myu"] = np.array([th.lambda_, ba.myu])
    # 4. 50 NVT steps of the dilute gas (README lattice, 6^3 atoms)
    o4 = gas(6)
    orc.update_force(lj, o4, mode="n2")
    th4 = orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, 300.0)
    orc.step(lj, o4, 0.002, thermostat=th4, mode="n2", n_steps=50)
    out["gas216_nvt50_pos"], out["gas216_nvt50_vel"] = o4.pos.copy(), o4.vel.copy()
    m = orc.macro(o4)
    out["gas216_nvt50_macro"] = np.array([m["kinetic"], m["thermal"], m["potential"], m["temperature"], m["pressure"]])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_fixtures.npz"), **build())
    print("wrote", os.path.join(HERE, "oracle_fixtures.npz"))
