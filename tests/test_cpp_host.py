"""The C++ host mirror (rust-llkv_b200/host/llkv_gpu.hpp): the compiled-language host side above the C ABI.
CPU: it flattens the benchmark plans and a kitchen-sink predicate to exactly the bytes the Python mirror produces.
GPU: a C++ program drives Q6 end to end through Context / Column / Program / Aggregation and checks the result."""
import ctypes as C
import os
import subprocess

import pytest

from llkv_b200 import tpch
from llkv_b200.expr import (AggregateKind, AggregateSpec, Bound, CompareOp, DataType, Expr, Literal, Operator, ProgramCompiler, ScalarExpr,
                            flatten_aggregates, pred)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def build(src, out, link=False):
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", os.path.join(CPP, src), "-o", out]
    if link:
        lib = os.path.join(ROOT, "rust-llkv_b200", "csrc")
        cmd += ["-L" + lib, "-lllkv_gpu", "-Wl,-rpath," + lib]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def raw(arr, n):
    return bytes(C.string_at(C.addressof(arr), C.sizeof(arr._type_) * n)).hex() if n else ""


def py_program(e):
    cp = ProgramCompiler(e).compile()
    ops, n_ops, lits, n_lits, nodes, n_nodes, roots, n_roots = cp.c_arrays()
    return {"ops": (n_ops, raw(ops, n_ops)), "literals": (n_lits, raw(lits, n_lits)), "nodes": (n_nodes, raw(nodes, n_nodes)),
            "list_roots": (n_roots, raw(roots, n_roots))}


def py_aggs(specs):
    aggs, n_aggs, nodes, n_nodes = flatten_aggregates(specs)
    return {"specs": (n_aggs, raw(aggs, n_aggs)), "nodes": (n_nodes, raw(nodes, n_nodes))}


def test_cpp_mirror_flattens_like_the_python_mirror(tmp_path):
    exe = str(tmp_path / "dump")
    build("host_mirror_dump.cpp", exe)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.splitlines()
    got, cur = {}, None
    for ln in out:
        f = ln.split(" ")
        if f[0] in ("program", "aggregates"):
            cur = got.setdefault((f[0], f[1]), {})
        else:
            cur[f[0]] = (int(f[1]), f[2] if len(f) > 2 else "")
    mixed = Expr.Not(Expr.Or([
        Expr.And([pred(1, Operator.GreaterThanOrEquals(-500)), pred(1, Operator.LessThanOrEquals(500))]),
        Expr.And([pred(2, Operator.In([1, 2, 3, -7])), pred(10, Operator.Equals("N")), pred(3, Operator.GreaterThan(-50.0)), pred(1, Operator.IsNotNull)]),
        Expr.Compare(ScalarExpr.Column(1) * 2, CompareOp.LtEq, ScalarExpr.Column(4)),
        Expr.InList(ScalarExpr.Column(2), [5, Literal.Null(), ScalarExpr.Column(1)], negated=True),
        Expr.IsNull(ScalarExpr.Column(1) + ScalarExpr.Column(2)),
        Expr.Compare(ScalarExpr.Cast(ScalarExpr.Column(5), DataType.Float64), CompareOp.Gt, ScalarExpr.Literal(Literal.Decimal128(10**9, 4))),
        Expr.Literal(False)]))
    want = {("program", "q6"): py_program(tpch.q6_filter()), ("aggregates", "q6"): py_aggs(tpch.q6_aggregates()),
            ("program", "q1"): py_program(tpch.q1_filter()), ("aggregates", "q1"): py_aggs(tpch.q1_aggregates()),
            ("program", "mixed"): py_program(mixed)}
    assert set(got) == set(want)
    for k in want:
        assert got[k] == want[k], k


@pytest.mark.gpu
def test_cpp_host_runs_q6_on_the_gpu(tmp_path):
    exe = str(tmp_path / "q6")
    build("q6_cpp_host.cpp", exe, link=True)
    r = subprocess.run([exe, "1000003"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok ") and "lean=1" in r.stdout and "specialised=1" in r.stdout, r.stdout


def test_cpp_descriptor_walk_matches_the_python_mirror(tmp_path):
    """llkv::walk_descriptor (C++ mirror) over an in-memory pager: same surviving chunks as llkv_b200.metadata and the
    oracle's rule, with and without a pruning range."""
    from llkv_b200 import metadata
    from oracle import metadata as om
    exe = str(tmp_path / "walk")
    build("descriptor_walk.cpp", exe, link=True)
    metas = [(1000 + i, 0, 4096, 65560, om.sortable_u64(om.INT64, i * 10 - 700), om.sortable_u64(om.INT64, i * 10 - 691), 0, 10) for i in range(140)]
    pages = om.descriptor_pages(metas, [500, 501, 502])
    pager = {499: om.descriptor_bytes((1 << 32) | 7, 500, 502, 140 * 4096, 140, data_type_code=6), **dict(pages)}
    lines = [f"{k} {v.hex()}" for k, v in pager.items()]
    m64 = (1 << 64) - 1
    queries = [(2, 0, 2, 0), (0, (-95) & m64, 1, 205), (1, 9, 2, 0), (0, 5000, 0, 6000)]  # (lower kind, bits, upper kind, bits)
    lines += [f"walk 499 {om.INT64} {lk} {lb} {uk} {ub}" for lk, lb, uk, ub in queries]
    env = dict(os.environ)
    out = subprocess.run([exe], input="\n".join(lines) + "\n", check=True, capture_output=True, text=True, env=env).stdout.splitlines()
    assert len(out) == len(queries)
    for (lk, lb, uk, ub), ln in zip(queries, out):
        f = ln.split()
        assert f[:4] == ["desc", str(140 * 4096), "140", "6"], ln
        lower = None if lk == 2 else (lk, lb)
        upper = None if uk == 2 else (uk, ub)
        _, want, _ = metadata.walk_descriptor(lambda pks: [pager[k] for k in pks], 499, om.INT64, lower, upper)
        assert [int(x) for x in f[4:]] == [c.chunk_pk for c in want], (lk, lb, uk, ub)
    assert out[0].count(" ") == 3 + 140 and out[3].split()[4:] == []  # no range: every chunk; a range above all values: none
