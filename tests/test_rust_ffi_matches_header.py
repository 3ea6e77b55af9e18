"""The Rust crate (`mfcc-rust_b200/speechsauce-cuda`) cannot be compiled in this image (no Rust toolchain), so
its `extern "C"` block is checked textually against include/speechsauce_cuda.h: the same functions, the same
number and kind of arguments in the same order, the same return kinds, `#[repr(C)]` structs with the same
fields, and the same constants.  A drift between the two would be undefined behaviour at the first call."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "speechsauce_cuda.h").read_text()
FFI = (ROOT / "mfcc-rust_b200" / "speechsauce-cuda" / "src" / "ffi.rs").read_text()
RUST_SRC = "\n".join(p.read_text() for p in sorted((ROOT / "mfcc-rust_b200" / "speechsauce-cuda" / "src").glob("*.rs")))

C_KIND = {"int": "i32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "float": "f32", "size_t": "usize", "void": "void"}
RUST_KIND = {"c_int": "i32", "i32": "i32", "u32": "u32", "i64": "i64", "f32": "f32", "usize": "usize"}


def strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def c_type_kind(t):
    """'const float*' -> 'ptr', 'int64_t' -> 'i64' ..."""
    t = t.strip()
    if "*" in t:
        return "ptr"
    t = t.replace("const", "").strip()
    return C_KIND[t]


def rust_type_kind(t):
    t = t.strip()
    if t.startswith("*"):
        return "ptr"
    return RUST_KIND[t]


def c_functions():
    src = strip_c_comments(HEADER)
    out = {}
    for ret, name, args in re.findall(r"^\s*([A-Za-z_][\w\s\*]*?)\s*\b(ss_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src, flags=re.M):
        kinds = []
        if args.strip() != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"(.*?)([A-Za-z_]\w*)$", a)          # strip the parameter name
                kinds.append(c_type_kind(m.group(1)))
        out[name] = (c_type_kind(ret), kinds)
    return out


def rust_functions():
    block = FFI[FFI.index('extern "C" {'):]
    out = {}
    for name, args, ret in re.findall(r"pub fn (ss_[a-z0-9_]+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        kinds = [rust_type_kind(a.split(":", 1)[1]) for a in args.split(",") if a.strip()]
        out[name] = (rust_type_kind(ret) if ret else "void", kinds)
    return out


def c_structs():
    src = strip_c_comments(HEADER)
    out = {}
    for body, name in re.findall(r"typedef struct\s*\{([^}]*)\}\s*(\w+)\s*;", src):
        fields = []
        for line in body.split(";"):
            line = line.strip()
            if line:
                m = re.match(r"(.*?)([A-Za-z_]\w*)$", line)
                fields.append((m.group(2), c_type_kind(m.group(1))))
        out[name] = fields
    return out


def rust_structs():
    out = {}
    for name, body in re.findall(r"#\[repr\(C\)\]\s*(?:#\[derive\([^)]*\)\]\s*)?pub struct (\w+)\s*\{([^}]*)\}", FFI):
        fields = [(f.split(":")[0].replace("pub", "").strip(), rust_type_kind(f.split(":", 1)[1]))
                  for f in body.split(",") if ":" in f and not f.strip().startswith("_private")]
        out[name] = fields
    return out


def test_every_header_function_is_declared_in_ffi_rs_with_the_same_signature():
    c, r = c_functions(), rust_functions()
    assert len(c) >= 45
    assert set(c) == set(r), f"only in the header: {sorted(set(c) - set(r))}; only in ffi.rs: {sorted(set(r) - set(c))}"
    for name in sorted(c):
        assert c[name] == r[name], f"{name}: header {c[name]} != ffi.rs {r[name]}"


def test_repr_c_structs_have_the_header_layout():
    c, r = c_structs(), rust_structs()
    for name in ("ss_config", "ss_plan_info", "ss_feature_outputs"):
        assert c[name] == r[name], f"{name}: header {c[name]} != ffi.rs {r[name]}"
    assert re.search(r"#\[repr\(C\)\]\s*pub struct ss_plan\s*\{\s*_private: \[u8; 0\],?\s*\}", FFI)  # opaque


def test_constants_agree():
    for name, val in re.findall(r"#define\s+(SS_(?:FLAG|CMVN)_\w+)\s+(0x[0-9a-fA-F]+|\d+)u?", HEADER):
        m = re.search(rf"pub const {name}: \w+ = (0x[0-9a-fA-F]+|\d+);", FFI)
        assert m, f"{name} is not defined in ffi.rs"
        assert int(m.group(1), 0) == int(val, 0), name
    for name, val in re.findall(r"\b(SS_(?:OK|ERR_\w+))\s*=\s*(-?\d+)", strip_c_comments(HEADER)):
        m = re.search(rf"\b{name}\b[^=\n]*=\s*(-?\d+)", RUST_SRC)
        if m:  # the crate turns every non-zero status into a panic; where it names a code the value must agree
            assert int(m.group(1)) == int(val), name


def test_build_rs_compiles_every_cuda_source_for_sm_100a():
    build_rs = (ROOT / "mfcc-rust_b200" / "speechsauce-cuda" / "build.rs").read_text()
    assert "sm_100a" in build_rs
    for src in sorted((ROOT / "mfcc-rust_b200" / "csrc").glob("*.cu")) + sorted((ROOT / "mfcc-rust_b200" / "csrc").glob("*.cpp")):
        assert src.name in build_rs, f"{src.name} is missing from build.rs"
