"""CPU: host-side logic — program flattening, the C ABI surface, loud failure without a GPU, workload builders."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from llkv_b200 import ffi, tpch
from llkv_b200.expr import (Bound, Expr, Operator, ProgramCompiler, ScalarExpr, flatten_aggregates, pred)
from llkv_b200.table import HostColumn, LlkvError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "llkv_gpu.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(llkv_gpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from llkv_b200 import gpu
    lib = C.CDLL(gpu.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/llkv_gpu.h but not exported"
    assert gpu.load().llkv_gpu_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    from llkv_b200 import gpu
    if gpu.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(LlkvError) as e:
        gpu.Context(0)
    assert e.value.code == ffi.ERR_IO and "no CPU fallback" in e.value.message


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "rust-llkv_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "llkv_oracle" not in text and "import oracle" not in text and "from oracle" not in text, os.path.join(base, f)


def test_struct_layouts_match_the_header():
    # sizes the C compiler gives the header's structs (the oracle library is built from the same header)
    assert C.sizeof(ffi.Literal) == 24
    assert C.sizeof(ffi.ScalarNode) == 56
    assert C.sizeof(ffi.EvalOp) == 56
    assert C.sizeof(ffi.AggSpec) == 16
    assert C.sizeof(ffi.AggValue) == 24
    assert C.sizeof(ffi.GroupKey) == 16


def test_same_field_ands_fuse():
    # gather_fused (llkv-compute/src/program.rs:415-439): every child a Pred on one field -> FusedAnd
    cp = ProgramCompiler(tpch.between_filter(1, 5, 9)).compile()
    assert [o.tag for o in cp.ops] == [ffi.EV_FUSED_AND, ffi.EV_FILTER_ITEM, ffi.EV_FILTER_ITEM]
    assert cp.ops[0].child_count == 2 and cp.ops[0].field_id == 1
    cp = ProgramCompiler(tpch.q6_filter()).compile()
    assert [o.tag for o in cp.ops] == [ffi.EV_PUSH_PREDICATE] * 3 + [ffi.EV_AND]
    assert cp.ops[0].operator_tag == ffi.OP_RANGE and cp.ops[0].lower_kind == ffi.BOUND_INCLUDED and cp.ops[0].upper_kind == ffi.BOUND_EXCLUDED
    assert cp.ops[3].child_count == 3


def test_postfix_order_and_node_pool():
    e = Expr.Not(Expr.Or([pred(1, Operator.Equals(1)), Expr.Compare(ScalarExpr.Column(1) + 2, 4, ScalarExpr.Column(2))]))
    cp = ProgramCompiler(e).compile()
    assert [o.tag for o in cp.ops] == [ffi.EV_PUSH_PREDICATE, ffi.EV_PUSH_COMPARE, ffi.EV_OR, ffi.EV_NOT]
    nodes = cp.pool.nodes
    cmp_op = cp.ops[1]
    assert nodes[cmp_op.expr_left].tag == ffi.SE_BINARY and nodes[cmp_op.expr_right].tag == ffi.SE_COLUMN
    assert nodes[nodes[cmp_op.expr_left].left].field_id == 1
    with pytest.raises(ValueError):
        ProgramCompiler(Expr.And([])).compile()


def test_q1_plan_shape():
    aggs, n, nodes, n_nodes = flatten_aggregates(tpch.q1_aggregates())
    assert n == 8 and aggs[7].expr_root == -1 and aggs[3].data_type == ffi.PT_DECIMAL128 and aggs[3].scale == 6
    assert n_nodes > 10


def test_lineitem_generator_is_deterministic_and_in_range():
    a = tpch.lineitem_arrays(10_000, seed=6)
    b = tpch.lineitem_arrays(10_000, seed=6)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert a["quantity"].min() >= 100 and a["quantity"].max() <= 5000
    assert a["discount"].min() == 0 and a["discount"].max() == 10
    assert a["extendedprice"].max() <= 50 * 210_000
    assert set(np.unique(a["returnflag"])) <= {ord("A"), ord("N"), ord("R")}
    sel = ((a["shipdate"] >= tpch.date32(1994, 1, 1)) & (a["shipdate"] < tpch.date32(1995, 1, 1)) & (a["discount"] >= 5)
           & (a["discount"] <= 7) & (a["quantity"] < 2400)).mean()
    assert 0.01 < sel < 0.03  # Q6 selectivity ~1.9 %


def test_shard_ranges_cover_the_table_once():
    for n in (0, 1, 131_072, 1_000_003, 59_986_052):
        for w in (1, 2, 3, 8):
            spans = [tpch.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            for a, b in spans[:-1]:
                assert a % 131_072 == 0 and (b % 131_072 == 0 or b == n)


def test_serialize_rejects_nulls_and_varlen():
    col = HostColumn.utf8(1, ["a", None])
    with pytest.raises(ValueError):
        col.serialize()


def test_rust_sys_crate_declares_every_function_of_the_header():
    """ffi/llkv-gpu-sys is authored, not compiled here: at least every entry point of include/llkv_gpu.h must be declared."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rs = open(os.path.join(root, "ffi", "llkv-gpu-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (llkv_gpu_[a-z0-9_]+)\s*\(", rs))
    missing = [s for s in header_symbols() if s not in declared]
    assert not missing, missing


def test_safe_rust_crate_only_calls_declared_sys_functions():
    """ffi/llkv-gpu (authored, not compiled here) may only use entry points and types the sys crate declares."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys_rs = open(os.path.join(root, "ffi", "llkv-gpu-sys", "src", "lib.rs")).read()
    src_dir = os.path.join(root, "ffi", "llkv-gpu", "src")
    safe_rs = "".join(open(os.path.join(src_dir, f)).read() for f in sorted(os.listdir(src_dir)) if f.endswith(".rs"))
    declared = set(re.findall(r"pub fn (llkv_gpu_[a-z0-9_]+)\s*\(", sys_rs)) | set(re.findall(r"pub struct (llkv_[a-z0-9_]+)", sys_rs))
    used = set(re.findall(r"sys::(llkv_[a-z0-9_]+)", safe_rs))
    assert used and not (used - declared), sorted(used - declared)


def test_narrowing_loops_agree_in_every_instruction_set_form():
    """upload.cpp narrows Arrow Decimal128 values to i64 / i32 on the host before the DMA, with a fit check.  The loop is
    dispatched on the CPU (SSE2 / AVX2 / AVX-512): every form this machine can run must write the same bytes and return the
    same verdict as a numpy restatement — for every length around the vector widths, unaligned sources, both signs, values
    at the edges of the narrow type, and a misfit at every position of a vector."""
    from llkv_b200 import gpu
    lib = C.CDLL(gpu.LIB_PATH)
    fn = lib.llkv_internal_narrow_d128
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
    rng = np.random.default_rng(12)
    forms = [isa for isa in (0, 1, 2, 3) if fn(4, isa, None, None, 0) >= 0]
    assert 0 in forms and 1 in forms

    def check(width, lo, hi):
        n = lo.size
        raw = np.zeros(2 * n + 1, dtype=np.int64)  # (+1: an unaligned view below)
        src = raw[:2 * n].reshape(n, 2) if n else raw[:0].reshape(0, 2)
        src[:, 0], src[:, 1] = lo, hi
        narrow_ok = (hi == (lo >> 63)) & ((lo == lo.astype(np.int32)) if width == 4 else True)
        want_ok = bool(np.all(narrow_ok))
        want = lo.astype(np.int32) if width == 4 else lo.copy()
        # an 8-byte-aligned but not 16/32/64-byte-aligned source, as ARR0 payloads are (24-byte header)
        shifted = np.zeros(2 * n + 3, dtype=np.int64)
        shifted[1:2 * n + 1] = src.reshape(-1)
        for isa in forms:
            for buf, off in ((src, 0), (shifted, 8)):
                out = np.full(n + 2, 0x5A5A5A5A, dtype=np.int32 if width == 4 else np.int64)
                rc = fn(width, isa, buf.ctypes.data + off, out.ctypes.data, n)
                assert rc == int(want_ok), (width, isa, n, off)
                assert np.array_equal(out[:n], want) and np.all(out[n:] == 0x5A5A5A5A), (width, isa, n, off)

    for width in (4, 8):
        lim = 2**31 if width == 4 else 2**63
        for n in list(range(0, 70)) + [127, 128, 129, 1000, 65536 + 5]:
            lo = rng.integers(-lim, lim, n, dtype=np.int64)
            if n:
                lo[rng.integers(0, n, min(n, 4))] = [lim - 1, -lim, 0, -1][:min(n, 4)]
            check(width, lo, lo >> 63)
        # one misfit at every position of a 64-value block: a wrong high word, a value outside the narrow type
        base = rng.integers(-1000, 1000, 64, dtype=np.int64)
        for pos in range(64):
            for bad_hi in (1, -2, 2**62, (base[pos] >> 63) ^ 1):
                hi = base >> 63
                hi[pos] = bad_hi
                check(width, base.copy(), hi)
            if width == 4:
                for v in (2**31, -2**31 - 1, 2**40, -2**62):
                    lo = base.copy()
                    lo[pos] = v
                    check(width, lo, lo >> 63)
