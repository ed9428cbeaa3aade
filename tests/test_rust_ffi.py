"""CPU: the authored Rust side of the boundary (ffi/llkv-gpu-sys, ffi/llkv-gpu) cannot be compiled here (no Rust toolchain in
the image), so what CAN be checked is checked against the header the C compiler sees: every `#[repr(C)]` struct of the sys
crate has the header's fields in the header's order with the header's offsets and size (gcc's `offsetof`), and every tag
table of the flattening code (`EvalOp`, `OwnedOperator`, `Bound`, `ScalarExpr`, `Literal`, `AggregateKind`, PrimType,
`BinaryOp`, `CompareOp`) carries the header's values."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "llkv_gpu.h")
SYS_RS = os.path.join(ROOT, "ffi", "llkv-gpu-sys", "src", "lib.rs")
FLATTEN_RS = os.path.join(ROOT, "ffi", "llkv-gpu", "src", "flatten.rs")

RUST_TYPES = {"i8": (1, 1), "u8": (1, 1), "i16": (2, 2), "u16": (2, 2), "i32": (4, 4), "u32": (4, 4), "f32": (4, 4), "i64": (8, 8), "u64": (8, 8),
              "f64": (8, 8)}


def strip_comments(src):
    return re.sub(r"/\*.*?\*/", "", re.sub(r"//[^\n]*", "", src), flags=re.S)


def header_structs():
    """{name: [field names]} of the plain-data structs the header defines."""
    src = strip_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} (\w+);", src, flags=re.S):
        name, body = m.group(1), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if not decl.startswith(("const ", "unsigned ")) else decl.split(None, 2)[2]
            for n in names.split(","):
                fields.append(re.sub(r"\[[^\]]*\]", "", n.strip()).replace("*", "").strip())
        out[name] = fields
    return out


def c_layouts(structs):
    """sizeof / offsetof as gcc sees them."""
    lines = ['#include <stddef.h>', '#include <stdio.h>', f'#include "{HEADER}"', "int main(void) {"]
    for name, fields in structs.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'  printf("{name} {f} %zu\\n", offsetof({name}, {f}));')
    lines += ["  return 0;", "}"]
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "l.c"), os.path.join(d, "l")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-std=c11", src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    lay = {}
    for ln in out.splitlines():
        s, f, v = ln.split()
        lay.setdefault(s, {})[f] = int(v)
    return lay


def rust_structs():
    """{name: [(field, size, align)]} of the `#[repr(C)]` structs of the sys crate, nested structs resolved."""
    src = re.sub(r"//[^\n]*", "", open(SYS_RS).read())
    raw = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[derive\([^\]]*\)\]\s*)*pub struct (\w+) \{(.*?)\n\}", src, flags=re.S):
        fields = re.findall(r"pub (\w+): ([^,\n]+),", m.group(2))
        raw[m.group(1)] = fields
    done = {}

    def size_align(ty):
        ty = ty.strip()
        arr = re.fullmatch(r"\[(.+); (\d+)\]", ty)
        if arr:
            s, a = size_align(arr.group(1))
            return s * int(arr.group(2)), a
        if ty in RUST_TYPES:
            return RUST_TYPES[ty]
        if ty.startswith("*"):
            return 8, 8
        fields = layout(ty)
        return fields["__size"], fields["__align"]

    def layout(name):
        if name in done:
            return done[name]
        off, align, out = 0, 1, {}
        for f, ty in raw[name]:
            s, a = size_align(ty)
            off = (off + a - 1) // a * a
            out[f.rstrip("_") if f == "type_" else f] = off
            off += s
            align = max(align, a)
        out["__size"] = (off + align - 1) // align * align
        out["__align"] = align
        done[name] = out
        return out

    return {n: layout(n) for n in raw if raw[n]}


def test_repr_c_structs_match_the_header_field_for_field():
    hs = header_structs()
    assert {"llkv_literal", "llkv_scalar_node", "llkv_eval_op", "llkv_agg_spec", "llkv_agg_value", "llkv_group_key", "llkv_run_info",
            "llkv_chunk_metadata", "llkv_column_descriptor", "llkv_range_bound", "llkv_debug_column"} <= set(hs)
    c = c_layouts(hs)
    r = rust_structs()
    for name, fields in hs.items():
        assert name in r, f"ffi/llkv-gpu-sys does not define {name}"
        assert r[name]["__size"] == c[name]["size"], (name, r[name]["__size"], c[name]["size"])
        rust_fields = [f for f in r[name] if not f.startswith("__")]
        assert rust_fields == fields, (name, rust_fields, fields)
        for f in fields:
            assert r[name][f] == c[name][f], (name, f, r[name][f], c[name][f])


def header_enums():
    src = strip_comments(open(HEADER).read())
    vals = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", src, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = (x.strip() for x in item.split("="))
                nxt = int(v, 0)
            else:
                k = item
            vals[k] = nxt
            nxt += 1
    return vals


def test_flattening_tables_carry_the_header_values():
    h = header_enums()
    src = open(FLATTEN_RS).read()
    consts = dict((k, int(v)) for k, v in re.findall(r"pub const (\w+): i32 = (-?\d+);", src))
    checked = 0
    for k, v in h.items():
        if not k.startswith("LLKV_") or k.startswith(("LLKV_ERR", "LLKV_OK", "LLKV_BIN", "LLKV_CMP", "LLKV_EXPR")):
            continue
        short = k[len("LLKV_"):]
        assert short in consts, f"flatten.rs lacks {short}"
        assert consts[short] == v, (k, consts[short], v)
        checked += 1
    assert checked >= 60
    # the match arms of binary_op_code / compare_op_code, in the reference's variant order (llkv-expr/src/expr.rs:311-349)
    names = {"Add": "ADD", "Subtract": "SUB", "Multiply": "MUL", "Divide": "DIV", "Modulo": "MOD", "And": "AND", "Or": "OR",
             "BitwiseShiftLeft": "SHL", "BitwiseShiftRight": "SHR"}
    for variant, code in re.findall(r"BinaryOp::(\w+) => (\d+),", src):
        assert h["LLKV_BIN_" + names[variant]] == int(code), variant
    cmp_names = {"Eq": "EQ", "NotEq": "NE", "Lt": "LT", "LtEq": "LE", "Gt": "GT", "GtEq": "GE"}
    for variant, code in re.findall(r"CompareOp::(\w+) => (\d+),", src):
        assert h["LLKV_CMP_" + cmp_names[variant]] == int(code), variant
    path = open(os.path.join(ROOT, "ffi", "llkv-gpu", "src", "path.rs")).read()
    assert re.search(r"EXPR_ARROW: i32 = 0;", path) and re.search(r"EXPR_EXACT: i32 = 1;", path)
    assert h["LLKV_EXPR_ARROW"] == 0 and h["LLKV_EXPR_EXACT"] == 1


def test_error_codes_follow_the_reference_variant_order():
    """include/llkv_gpu.h numbers its status codes like llkv_result::Error's variants; the safe crate's `check` maps them back."""
    h = header_enums()
    lib = open(os.path.join(ROOT, "ffi", "llkv-gpu", "src", "lib.rs")).read()
    arms = dict((int(c), v) for c, v in re.findall(r"(\d+) => Error::(\w+)", lib))
    want = {h["LLKV_ERR_IO"]: "Io", h["LLKV_ERR_ARROW"]: "Arrow", h["LLKV_ERR_INVALID_ARGUMENT"]: "InvalidArgumentError", h["LLKV_ERR_NOT_FOUND"]: "NotFound",
            h["LLKV_ERR_EXPR_CAST"]: "ExprCast", h["LLKV_ERR_PREDICATE_BUILD"]: "PredicateBuild"}
    for code, variant in want.items():
        assert arms.get(code) == variant, (code, variant, arms.get(code))
