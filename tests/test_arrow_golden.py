"""The arithmetic the reference delegates to arrow-arith / arrow-cast 57.1 (Decimal128 result types, the rescale back to the
preferred type with half-away-from-zero rounding, safe-cast NULLs, checked integer kernels, divisor zeros -> NULL), pinned by
tests/golden/arrow_known_answers.json (derivation and cross-checks: tests/golden/make_arrow_golden.py).  Every vector runs
through the oracle (CPU) and through the CUDA path (`-m gpu`): the expression is observed row by row through MIN(expr) over
single-row ranges, the same evaluator the ungrouped aggregates use (llkv-compute/src/eval.rs:565-614)."""
import json
import os

import numpy as np
import pytest

import util
from llkv_b200 import ffi
from llkv_b200.expr import AggregateKind, AggregateSpec, DataType
from llkv_b200.table import HostColumn, HostTable, LlkvError, decimal_array
from oracle import oracle

CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "arrow_known_answers.json")))["cases"]


def table_of(case) -> HostTable:
    t = HostTable(1)
    for c in case["columns"]:
        if c["type"] == "Decimal128":
            t.add(HostColumn(c["field"], DataType.Decimal128(c["precision"], c["scale"]), decimal_array(c["values"])))
        else:
            t.add(HostColumn(c["field"], DataType.Int64, np.asarray(c["values"], dtype=np.int64)))
    return t


def spec_type(case):
    st = case["spec_type"]
    return DataType.Decimal128(st["precision"], st["scale"]) if st["type"] == "Decimal128" else DataType.Int64


def check_case(case, aggregate):
    """`aggregate(table, specs, row_begin, row_end)` -> finalized rows, for the oracle or the GPU."""
    t = table_of(case)
    e = util.sexpr_from_json(case["expr"])
    st = spec_type(case)
    specs = [AggregateSpec("v", AggregateKind.Min(e, st)), AggregateSpec("n", AggregateKind.Count(e))]
    if "expect_error" in case:
        with pytest.raises(LlkvError) as err:
            aggregate(t, specs, 0, t.n_rows)
        assert err.value.code == case["expect_error"]["code"], err.value
        assert case["expect_error"]["contains"] in err.value.message, err.value.message
        return
    for i, want in enumerate(case["expect"]):
        v, n = aggregate(t, specs, i, i + 1)[0][1]
        assert n.value == (0 if want is None else 1), (case["name"], i)
        assert v.value == want, (case["name"], i, v, want)
        if want is not None and st.type == ffi.PT_DECIMAL128:
            assert v.type == ffi.PT_DECIMAL128 and v.scale == st.scale
    # and the whole column at once: SUM over the non-NULL results
    ok = [w for w in case["expect"] if w is not None]
    if st.type == ffi.PT_INT64:  # (SumInt64 errors at the first prefix that leaves i64: not what these vectors are about)
        run = 0
        for w in ok:
            run += w
            if not (-2**63 <= run < 2**63):
                return
    got = aggregate(t, [AggregateSpec("s", AggregateKind.Sum(e, st)), AggregateSpec("n", AggregateKind.Count(e))], 0, t.n_rows)[0][1]
    assert got[1].value == len(ok)
    if st.type == ffi.PT_DECIMAL128 or ok:
        assert got[0].value == sum(ok)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_arrow_known_answers_oracle(case):
    check_case(case, lambda t, specs, lo, hi: oracle.aggregate(t, None, specs, row_begin=lo, row_end=hi))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_arrow_known_answers_gpu(gpu_ctx, case):
    from llkv_b200 import gpu
    holder = {}

    def aggregate(t, specs, lo, hi):
        if "dt" not in holder:
            holder["dt"] = gpu.DeviceTable.from_host(gpu_ctx, t)
        return holder["dt"].aggregate(None, specs, row_begin=lo, row_end=hi)

    try:
        check_case(case, aggregate)
    finally:
        if "dt" in holder:
            holder["dt"].destroy()
