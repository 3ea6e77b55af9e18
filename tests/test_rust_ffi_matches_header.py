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


# The reference's public signatures for the hot path (parameter names dropped), keyed by
# module::function -> (parameter types, return type).  `lmfe` / `extract_derivative_feature` /
# `log_power_spectrum` are private in the reference and public here (D3); `ssc` is new (D4).
REFERENCE_SIGNATURES = {
    "feature::mfcc": (["ArrayView1<f32>", "&SpeechConfig"], "Array2<f32>"),                                  # feature.rs:99
    "feature::mfe": (["ArrayView1<f32>", "&SpeechConfig"], "(Array2<f32>, Array1<f32>)"),                    # feature.rs:200
    "feature::lmfe": (["ArrayView1<f32>", "&SpeechConfig"], "Array2<f32>"),                                  # feature.rs:242
    "feature::ssc": (["ArrayView1<f32>", "&SpeechConfig"], "Array2<f32>"),                                   # README.md:15
    "feature::mel_spectrogram1": (["ArrayView1<f32>", "&SpeechConfig"], "Array2<f32>"),                      # feature.rs:151
    "feature::mel_spectrogram2": (["ArrayView2<f32>", "&SpeechConfig"], "Array3<f32>"),                      # feature.rs:163
    "feature::extract_derivative_feature": (["Array2<f32>"], "Array3<f32>"),                                 # feature.rs:253
    "processing::preemphasis": (["Array1<f32>", "isize", "f32"], "Array1<f32>"),                             # processing.rs:31
    "processing::stack_frames": (["ArrayView1<f32>", "usize", "f32", "f32", "Option<fn(usize) -> Array2<f32>>", "bool"],
                                 "Array2<f32>"),                                                             # processing.rs:65
    "processing::power_spectrum": (["Array2<f32>", "usize"], "Array2<f32>"),                                 # processing.rs:179
    "processing::log_power_spectrum": (["Array2<f32>", "usize", "bool"], "Array2<f32>"),                     # processing.rs:194
    "processing::derivative_extraction": (["&Array2<f32>", "usize"], "Array2<f32>"),                         # processing.rs:222
    "processing::cmvn": (["ArrayView2<f32>", "bool"], "Array2<f32>"),                                        # processing.rs:265
    "processing::cmvnw": (["Array2<f32>", "usize", "bool"], "Array2<f32>"),                                  # processing.rs:315
    "functions::stft2": (["ArrayView2<f32>", "&SpeechConfig"], "Array3<Complex32>"),                         # functions.rs:86
    "functions::stft1": (["ArrayView1<f32>", "&SpeechConfig"], "Array2<Complex32>"),                         # functions.rs:199
}


def _split_params(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "<(":
            depth += 1
        elif ch in ">)" and not cur.endswith("-"):     # "->" is not a closing bracket
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_crate_keeps_the_reference_public_signatures():
    src_dir = ROOT / "mfcc-rust_b200" / "speechsauce-cuda" / "src"
    norm = lambda t: re.sub(r"\s+", " ", t.strip())
    for key, (params, ret) in REFERENCE_SIGNATURES.items():
        module, fn = key.split("::")
        text = (src_dir / f"{module}.rs").read_text()
        m = re.search(rf"pub fn {fn}\s*\(([^{{]*)\{{", text)
        assert m, f"{key} is missing from the crate"
        head, _, got_ret = m.group(1).rpartition(") ->")           # the last ") ->" closes the parameter list
        got = [norm(p.split(":", 1)[1]) for p in _split_params(head)]
        assert got == params, f"{key}: parameters {got} != {params}"
        assert norm(got_ret) == ret, f"{key}: returns {norm(got_ret)} != {ret}"
    cfg = (src_dir / "config.rs").read_text()
    # SpeechConfig::new's nine positional arguments (config.rs:140-150) and the builder's setters (config.rs:35-96)
    m = re.search(r"impl SpeechConfig \{.*?pub fn new\((.*?)\)\s*->\s*Self", cfg, flags=re.S)
    names = [p.split(":")[0].strip() for p in _split_params(m.group(1))]
    assert names == ["sample_rate", "fft_points", "frame_length", "frame_stride", "num_cepstral", "num_filters",
                     "low_frequency", "high_frequency", "dc_elimination"]
    for setter in ("high_freq", "low_freq", "dc_elimination", "num_cepstral", "frame_stride", "frame_length", "fft_points", "build"):
        assert re.search(rf"pub fn {setter}\(", cfg), f"SpeechConfigBuilder::{setter} is missing"
