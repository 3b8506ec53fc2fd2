"""The Rust -sys crate (ffi/moldyn-b200-sys) is source a maintainer compiles where cargo exists; here its `extern "C"` block
and #[repr(C)] structs are diffed against include/moldyn_b200.h so the two cannot drift apart."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C2RUST = {
    "int": "c_int", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "uint8_t": "u8", "double": "f64", "void": "c_void",
    "char": "c_char", "md_ctx": "md_ctx", "md_config": "md_config", "md_thermostat": "md_thermostat",
    "md_barostat": "md_barostat", "md_macro_out": "md_macro_out", "md_stats": "md_stats",
}


def c_type_to_rust(t):
    t = t.strip()
    const = "const " in t
    base = t.replace("const", "").replace("*", "").strip()
    stars = t.count("*")
    r = C2RUST[base]
    for _ in range(stars):
        r = ("*const " if const else "*mut ") + r
    return r


def header_functions():
    src = open(os.path.join(ROOT, "include", "moldyn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"MD_API\s+([\w\s\*]+?)\s*\b(md_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                arr = re.match(r"(.*?)(\w+)\s*\[\w*\]$", a)      # `const double box[3]` decays to a pointer
                if arr:
                    params.append(c_type_to_rust(arr.group(1) + " *"))
                else:
                    params.append(c_type_to_rust(re.match(r"(.*?)(\w+)$", a).group(1)))
        out[name] = (None if ret == "void" else c_type_to_rust(ret), params)
    return out


def rust_functions():
    src = open(os.path.join(ROOT, "ffi", "moldyn-b200-sys", "src", "lib.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    out = {}
    for m in re.finditer(r"pub fn (md_\w+)\((.*?)\)(?:\s*->\s*([^;]+))?;", block, flags=re.S):
        params = [p.split(":", 1)[1].strip() for p in m.group(2).split(",") if p.strip()]
        out[m.group(1)] = (m.group(3).strip() if m.group(3) else None, params)
    return out


def test_extern_block_matches_the_header():
    h, r = header_functions(), rust_functions()
    assert len(h) >= 29
    assert set(h) == set(r), (sorted(set(h) - set(r)), sorted(set(r) - set(h)))
    for name in h:
        assert h[name] == r[name], (name, h[name], r[name])


def test_repr_c_structs_match_the_header():
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "moldyn_b200.h")).read(), flags=re.S)
    rs = open(os.path.join(ROOT, "ffi", "moldyn-b200-sys", "src", "lib.rs")).read()
    for name in ("md_config", "md_thermostat", "md_barostat", "md_macro_out", "md_stats"):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(\w+)\s+(\w+)(?:\[(\d+)\])?$", decl)
            ty = C2RUST[m.group(1)]
            c_fields.append((m.group(2), f"[{ty}; {m.group(3)}]" if m.group(3) else ty))
        rbody = re.search(r"pub struct %s \{(.*?)\n\}" % name, rs, flags=re.S).group(1)
        r_fields = [(m.group(1).rstrip("_") if m.group(1) == "box_" else m.group(1), m.group(2).strip())
                    for m in re.finditer(r"pub (\w+): ([^,]+),", rbody)]
        assert c_fields == r_fields, (name, c_fields, r_fields)


def test_shim_forwards_the_reference_signatures():
    shim = open(os.path.join(ROOT, "ffi", "shim", "gpu.rs")).read()
    for needle in ("pub fn update_force(potentials_database: &PotentialsDatabase, state: &mut State)",
                   "barostat: &mut Option<(&mut Barostat, f64)>, thermostat: &mut Option<(&mut Thermostat, f64)>",
                   "sys::md_upload_state_typed(", "sys::md_set_potential_pair(", "sys::md_download_state(", "sys::md_step(", "p.temp = self.vir[i]"):
        assert needle in shim, needle
    assert "/* " not in shim.split("impl GpuSession")[1].split("impl Drop")[0]  # upload/download have real bodies
