"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, its host-arithmetic entry points reproduce the reference's goldens, and the product never
reaches into oracle/ nor falls back to the CPU."""
import ctypes as C
import os
import re

import pytest

import moldyn_b200 as md
from moldyn_b200 import _ffi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "moldyn_b200.h")).read()
    return sorted(set(re.findall(r"MD_API\s+[\w\s\*]+?\b(md_\w+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build_library()
    lib = C.CDLL(_ffi.library_path())
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_ffi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_version_and_struct_layouts(tmp_path):
    """ctypes mirrors must have exactly the layout gcc gives the header's structs."""
    import subprocess
    assert b"sm_100a" in _ffi.lib().md_version()
    src = tmp_path / "sizes.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "moldyn_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(md_config), sizeof(md_thermostat),'
        ' sizeof(md_barostat), sizeof(md_macro_out), sizeof(md_stats), offsetof(md_macro_out, n),'
        ' offsetof(md_stats, skin), offsetof(md_config, skin)); return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_ffi.Config), C.sizeof(_ffi.ThermostatC), C.sizeof(_ffi.BarostatC), C.sizeof(_ffi.MacroOut),
            C.sizeof(_ffi.Stats), _ffi.MacroOut.n.offset, _ffi.Stats.skin.offset, _ffi.Config.skin.offset]
    assert got == want


def test_scalar_potential_entry_points(kats):  # solver/src/lib.rs:77-89
    p = md.Potential.new_lennard_jones(kats["argon"]["sigma"], kats["argon"]["eps"])
    assert p.r_cut == 0.3418 * 2.5
    assert p.u_cut == -0.027934517624831987
    u, f = p.get_potential_and_force(kats["lennard_jones"]["r"])
    assert f"{u:.8f}" == kats["lennard_jones"]["potential"]
    assert f"{f:.8f}" == kats["lennard_jones"]["force"]
    assert p.get_potential_and_force(p.r_cut + 1e-9) == (0.0, 0.0)
    assert p.get_radius_cut() == p.r_cut


def test_scalar_potential_matches_oracle_bitwise():
    from oracle import oracle as orc
    o = orc.LennardJones()
    p = md.Potential.new_lennard_jones(0.3418, 1.712)
    for r in (0.3, 0.3418, 0.37, 0.5, 0.8, 0.8545, 0.9):
        assert p.get_potential_and_force(r) == o.get_potential_and_force(r)


def test_potentials_database_json_roundtrip(tmp_path):  # potential.rs:104-154
    db = md.PotentialsDatabase()
    assert db.get_potential(0, 0) is db.default_potential
    db.set_potential(1, 0, md.Potential(0.3418, 1.712, 1.1963, -0.003723224030513348))
    assert db.get_potential(0, 1).r_cut == 1.1963
    db.save_potentials_to_file(str(tmp_path))
    text = open(tmp_path / "potentials.json").read()
    assert '"0,1"' in text and '"LennardJones"' in text
    db2 = md.PotentialsDatabase()
    db2.load_potentials_from_file(str(tmp_path))
    assert db2.get_potential(1, 0).u_cut == -0.003723224030513348


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(md.MdError) as e:
        md.Solver()
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    pat = re.compile(r"oracle")
    for base, _, files in os.walk(os.path.join(ROOT, "moldyn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(base, f)).read()
                assert not pat.search(text), os.path.join(base, f)
    assert not pat.search(open(os.path.join(ROOT, "include", "moldyn_b200.h")).read())
