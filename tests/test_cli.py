"""The moldyn_cli-compatible driver (moldyn_b200/host): flags of cli/src/args.rs, file formats of
core/src/save_data.rs / particles_database.rs / potential.rs.  Mirrors cli/src/tests.rs."""
import csv
import json
import os
import subprocess

import numpy as np
import pytest

from moldyn_b200 import build

from helpers import f8


@pytest.fixture(scope="module")
def cli():
    return build.build_cli()


def run(cli, *args):
    return subprocess.run([cli, *map(str, args)], check=True, capture_output=True, text=True)


def read_frame(d, n):
    with open(os.path.join(d, "data", f"{n}.csv")) as f:
        rows = list(csv.DictReader(f))
    pos = np.array([[float(r[f"position_{a}"]) for a in "xyz"] for r in rows])
    vel = np.array([[float(r[f"velocity_{a}"]) for a in "xyz"] for r in rows])
    with open(os.path.join(d, "bb.csv")) as f:
        bbs = [[float(r[a]) for a in "xyz"] for r in csv.DictReader(f)]
    return pos, vel, np.array(bbs[n]), [int(r["id"]) for r in rows]


def write_frame0(d, pos, vel, box, name="Argon", mass=66.335, radius=0.071):
    os.makedirs(os.path.join(d, "data"), exist_ok=True)
    with open(os.path.join(d, "data", "0.csv"), "w") as f:
        f.write("id,position_x,position_y,position_z,velocity_x,velocity_y,velocity_z\n")
        for p, v in zip(pos, vel):
            f.write("0," + ",".join(repr(float(x)) for x in (*p, *v)) + "\n")
    with open(os.path.join(d, "bb.csv"), "w") as f:
        f.write("x,y,z\n" + ",".join(repr(float(x)) for x in box) + "\n")
    with open(os.path.join(d, "db.csv"), "w") as f:
        f.write(f"id,name,mass,radius\n0,{name},{mass!r},{radius!r}\n")


def test_initialization(cli, tmp_path, kats):  # cli/src/tests.rs:9-33
    g = kats["initialization"]
    d = str(tmp_path / "run")
    run(cli, "-f", d, "initialize", "-t", "u", "-s", *g["grid"], "-n", "Argon", "-m", 66.335, "-r", 0.071,
        "-l", g["cell"], "-T", 273.15)
    pos, vel, box, ids = read_frame(d, 0)
    assert len(ids) == g["count"] and set(ids) == {0}
    assert list(box) == [g["box"]] * 3
    assert open(os.path.join(d, "db.csv")).read() == "id,name,mass,radius\n0,Argon,66.335,0.071\n"
    # z is the fastest index (solver/src/lib.rs:17-47) and the halves are antisymmetric (velocity.rs:12-28)
    assert list(pos[1]) == [0.0, 0.0, g["cell"]]
    assert np.array_equal(vel[500:], -vel[:500])
    out = run(cli, "-f", d, "particle-count").stdout
    assert out.strip() == "Particle count: 1000"


def test_initialize_fcc_and_seed(cli, tmp_path):
    d1, d2 = str(tmp_path / "a"), str(tmp_path / "b")
    for d in (d1, d2):
        run(cli, "-f", d, "initialize", "-t", "fcc", "-s", "2 3 4", "-n", "Cu", "-m", 105.5, "-r", 0.128,
            "-l", 0.3615, "-T", 300, "--seed", 7)
    a, b = read_frame(d1, 0), read_frame(d2, 0)
    assert a[0].shape == (96, 3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert list(a[0][1]) == [0.0, 0.5 * 0.3615, 0.5 * 0.3615]


def test_frame_text_is_shortest_roundtrip(cli, tmp_path):
    """Floats are written like the reference's csv/ryu writer: shortest digits that round-trip, '1.0' not '1'."""
    d = str(tmp_path / "run")
    run(cli, "-f", d, "initialize", "-t", "u", "-s", "2", "2", "2", "-n", "Argon", "-m", 66.335, "-r", 0.071,
        "-l", 3.338339, "-T", 273.15, "--seed", 3)
    text = open(os.path.join(d, "data", "0.csv")).read().splitlines()
    assert text[0] == "id,position_x,position_y,position_z,velocity_x,velocity_y,velocity_z"
    assert text[1].startswith("0,0.0,0.0,0.0,")
    for line in text[1:]:
        for tok in line.split(",")[1:]:
            assert repr(float(tok)) == tok or tok.endswith(".0") or "e" in tok, tok
            assert float(tok) == float(repr(float(tok)))


def test_potentials_file_commands(cli, tmp_path):  # cli/src/commands.rs:21-41
    d = str(tmp_path / "run")
    run(cli, "-f", d, "generate-default-potentials")
    data = json.load(open(os.path.join(d, "potentials.json")))
    assert data == {"0,0": {"LennardJones": {"sigma": 0.3418, "eps": 1.712, "r_cut": 0.8545,
                                            "u_cut": -0.027934517624831987}}}
    run(cli, "-f", d, "set-potential", "-i", "0", "1", "-p", "lennard-jones", "--params", "0.34", "1.7")
    data = json.load(open(os.path.join(d, "potentials.json")))
    assert set(data) == {"0,0", "0,1"} and data["0,1"]["LennardJones"]["r_cut"] == 0.34 * 2.5


def test_async_frame_writer(cli, tmp_path):
    """`solve` hands frames to a writer thread.  `copy-frames` drives the same FrameWriter without a GPU:
    frames must land on disk byte-identical to the synchronous writer's, bb.csv must get one row per frame, in order."""
    d = str(tmp_path / "run")
    run(cli, "-f", d, "initialize", "-t", "u", "-s", 6, 6, 6, "-n", "Argon", "-m", 66.335, "-r", 0.071,
        "-l", 3.338339, "-T", 273.15, "--seed", 3)
    frame0 = open(os.path.join(d, "data", "0.csv")).read()
    run(cli, "-f", d, "copy-frames", "-s", 0, "-c", 7)
    for k in range(1, 8):
        assert open(os.path.join(d, "data", f"{k}.csv")).read() == frame0
    with open(os.path.join(d, "bb.csv")) as f:
        rows = [[float(r[a]) for a in "xyz"] for r in csv.DictReader(f)]
    assert len(rows) == 8
    assert [r[0] - rows[0][0] for r in rows] == [float(k) for k in range(8)]
    # a second run overwrites rows 1..3 in place (save_data.rs:170-176 semantics) and keeps the rest
    run(cli, "-f", d, "copy-frames", "-s", 0, "-c", 3)
    with open(os.path.join(d, "bb.csv")) as f:
        assert len(list(csv.DictReader(f))) == 8


def test_solve_without_gpu_fails_loudly(cli, tmp_path, kats):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    k = kats["two_body"]
    d = str(tmp_path / "run")
    write_frame0(d, k["pos"], k["vel"], k["box"])
    r = subprocess.run([cli, "-f", d, "solve", "-s", "0", "-i", "verlet-method", "-c", "3", "-t", "0.002"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_solvation(cli, tmp_path, kats):  # cli/src/tests.rs:35-98
    k = kats["two_body"]
    d = str(tmp_path / "run")
    write_frame0(d, k["pos"], k["vel"], k["box"])
    run(cli, "-f", d, "--time", "solve", "-s", "0", "-i", "verlet-method", "-c", "3", "-t", "0.002")
    assert sorted(os.listdir(os.path.join(d, "data"))) == ["0.csv", "1.csv", "2.csv", "3.csv"]
    for g in k["steps"]:
        pos, vel, box, _ = read_frame(d, g["step"])
        assert [f8(x) for x in pos[0]] == g["pos1"] and [f8(x) for x in pos[1]] == g["pos2"]
        assert [f8(x) for x in vel[0]] == g["vel1"] and [f8(x) for x in vel[1]] == g["vel2"]
        assert list(box) == k["box"]


@pytest.mark.gpu
def test_solve_nvt_npt_frames_match_oracle(cli, tmp_path):
    """README-style run (C1 config): thermostat + barostat flags, --frames-per-save, then solve-macro-parameters."""
    from oracle import oracle as orc
    o = orc.argon_lattice(6, orc.GAS_CELL, 273.15, seed=9)
    rng = np.random.default_rng(2)
    o.pos += rng.uniform(-1.4, 1.4, o.pos.shape)
    orc.apply_boundary_conditions(o)
    d = str(tmp_path / "run")
    write_frame0(d, o.pos, o.vel, o.box)
    run(cli, "-f", d, "--frames-per-save", "10", "solve", "-s", "0", "-c", "40", "-t", "0.002", "-i", "verlet-method",
        "--thermostat", "berendsen", "--thermostat-params", "10", "-T", "300",
        "--barostat", "berendsen", "--barostat-params", "1 5", "-P", "1.01325")
    assert sorted(os.listdir(os.path.join(d, "data")), key=lambda s: int(s[:-4])) == [f"{i}.csv" for i in range(5)]
    lj = orc.LennardJones()
    orc.update_force(lj, o)
    th, ba = orc.Thermostat(orc.Thermostat.BERENDSEN, 10.0, 300.0), orc.Barostat(1.0, 5.0, 1.01325)
    for frame in range(1, 5):
        orc.step(lj, o, 0.002, thermostat=th, barostat=ba, n_steps=10)
        pos, vel, box, _ = read_frame(d, frame)
        assert np.abs(box / o.box - 1).max() < 1e-12
        assert np.abs(vel - o.vel).max() < 1e-9
        dx = np.abs(pos - o.pos)
        assert np.minimum(dx, np.abs(dx - o.box)).max() < 1e-9
    run(cli, "-f", d, "solve-macro-parameters", "-A")
    rows = list(csv.DictReader(open(os.path.join(d, "macro.csv"))))
    assert [int(r["iteration"]) for r in rows] == list(range(5))
    last = orc.State(*read_frame(d, 4)[:2], orc.ARGON_MASS, read_frame(d, 4)[2])
    orc.update_force(lj, last)
    m = orc.macro(last)
    for key, col in (("kinetic", "kinetic_energy"), ("thermal", "thermal_energy"), ("potential", "potential_energy"),
                     ("temperature", "temperature"), ("pressure", "pressure")):
        assert abs(float(rows[4][col]) - m[key]) <= 1e-10 * max(1.0, abs(m[key])), key
    assert abs(float(rows[4]["unit_kinetic_energy"]) - m["kinetic"] / last.n) <= 1e-12


@pytest.mark.gpu
def test_solve_two_particle_types(cli, tmp_path):
    """A State with two particle types through `moldyn_cli solve` (typed upload, the type-pair potentials of potentials.json,
    the reference's one-sided cross-type forces and last-type thermostat) against the oracle's multi-type step; then
    solve-macro-parameters, which the reference evaluates for particle type 0 (cli/src/commands.rs:237-264)."""
    from oracle import oracle as orc
    from test_multi_type_cpu import mixture
    o = mixture(n_side=6, cell=0.45, counts=(120, 96), masses=(66.335, 20.18), temperature=150.0, seed=4)
    d = str(tmp_path / "run")
    os.makedirs(os.path.join(d, "data"))
    with open(os.path.join(d, "data", "0.csv"), "w") as f:
        f.write("id,position_x,position_y,position_z,velocity_x,velocity_y,velocity_z\n")
        for t, p, v in zip(o.types(), o.pos, o.vel):
            f.write(f"{t}," + ",".join(repr(float(x)) for x in (*p, *v)) + "\n")
    with open(os.path.join(d, "bb.csv"), "w") as f:
        f.write("x,y,z\n" + ",".join(repr(float(x)) for x in o.box) + "\n")
    with open(os.path.join(d, "db.csv"), "w") as f:
        f.write("id,name,mass,radius\n0,Argon,66.335,0.071\n1,Neon,20.18,0.038\n")
    tab = orc.PotentialTable(2)
    tab.set_potential(0, 1, orc.LennardJones(0.31, 1.1))
    cross = tab.get_potential(0, 1)
    dflt = orc.LennardJones()
    with open(os.path.join(d, "potentials.json"), "w") as f:
        json.dump({"0,0": {"LennardJones": {"sigma": dflt.sigma, "eps": dflt.eps, "r_cut": dflt.r_cut, "u_cut": dflt.u_cut}},
                   "0,1": {"LennardJones": {"sigma": cross.sigma, "eps": cross.eps, "r_cut": cross.r_cut,
                                            "u_cut": cross.u_cut}}}, f)
    run(cli, "-f", d, "--frames-per-save", "10", "solve", "-s", "0", "-c", "20", "-t", "0.002", "-i", "verlet-method",
        "--thermostat", "berendsen", "--thermostat-params", "0.5", "-T", "200", "-p")
    orc.update_force_multi(tab, o)
    th = orc.Thermostat(orc.Thermostat.BERENDSEN, 0.5, 200.0)
    for frame in (1, 2):
        orc.step_multi(tab, o, 0.002, th, n_steps=10)
        pos, vel, box, ids = read_frame(d, frame)
        assert ids == list(o.types())
        assert np.abs(vel - o.vel).max() < 1e-9
        dx = np.abs(pos - o.pos)
        assert np.minimum(dx, np.abs(dx - o.box)).max() < 1e-9
    run(cli, "-f", d, "solve-macro-parameters", "-A", "--use-potentials")
    rows = list(csv.DictReader(open(os.path.join(d, "macro.csv"))))
    pos, vel, box, _ = read_frame(d, 2)
    last = orc.MultiState(pos, vel, o.counts, o.masses, box)
    orc.update_force_multi(tab, last)
    m = orc.macro_type(last, 0)
    for key, col in (("kinetic", "kinetic_energy"), ("thermal", "thermal_energy"), ("potential", "potential_energy"),
                     ("temperature", "temperature"), ("pressure", "pressure")):
        assert abs(float(rows[2][col]) - m[key]) <= 1e-9 * max(1.0, abs(m[key])), key
