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
