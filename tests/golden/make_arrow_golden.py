#!/usr/bin/env python
"""Writes tests/golden/arrow_known_answers.json: known answers for the arithmetic the reference delegates to the arrow crates
(arrow-arith / arrow-cast 57.1.0, pinned in /root/reference/Cargo.toml:40-45 but NOT vendored under /root/reference, and
there is no network here, so no vector can be copied from the crates' own test files).

What each vector rests on instead, stated per case in "source" / "rule":
  * the rule as arrow-rs publishes it — arrow-arith/src/numeric.rs `decimal_op` (result precision / scale of add, sub, mul:
    "max(s1, s2) + max(p1 - s1, p2 - s2) + 1", "p1 + p2 + 1" / "s1 + s2"), the checked integer kernels of the same file
    (`add`, `sub`, `mul`, `div`, `rem`: ArrowError::ArithmeticOverflow / DivideByZero), arrow-cast/src/cast/decimal.rs
    (decimal -> decimal to a smaller scale divides and rounds half away from zero; a value that does not fit the target
    precision is NULL under the default `safe` cast options) — and the reference's own call sites
    (llkv-compute/src/kernels.rs:98-177 `compute_binary`: operands coerced to the common type, zeros of a divisor become NULL,
    arrow errors become Error::Internal; llkv-compute/src/eval.rs:565-614: the result is cast to the preferred type);
  * every expected value is computed here TWICE, independently of the oracle: with Python's `decimal` module
    (ROUND_HALF_UP = ties away from zero) and with pyarrow (Arrow C++, same Decimal128 type rules; `pc.round(...,
    round_mode="half_towards_infinity")` is the tie-breaking rule in question; `*_checked` kernels overflow the same way).
    The script asserts that both agree before it writes a vector.

Run from the repo root:  python tests/golden/make_arrow_golden.py
"""
import decimal
import json
import os
from decimal import Decimal

import pyarrow as pa
import pyarrow.compute as pc

decimal.getcontext().prec = 80
HERE = os.path.dirname(os.path.abspath(__file__))
I64_MAX, I64_MIN = 2**63 - 1, -2**63


def dec(raw, scale):
    return Decimal(raw).scaleb(-scale)


def fits(raw, precision):
    return abs(raw) < 10 ** precision


def rescale_half_away(raw, from_scale, to_scale):
    """arrow-cast decimal -> decimal: the raw integer at the new scale (to_scale < from_scale rounds half away from zero)."""
    if to_scale >= from_scale:
        return raw * 10 ** (to_scale - from_scale)
    q = dec(raw, from_scale).quantize(Decimal(1).scaleb(-to_scale), rounding=decimal.ROUND_HALF_UP)
    return int(q.scaleb(to_scale))


def pa_round_half_away(raw, p, s, to_scale):
    arr = pa.array([dec(raw, s)], pa.decimal128(p, s))
    out = pc.round(arr, to_scale, round_mode="half_towards_infinity")[0].as_py()
    return int(out.scaleb(to_scale))


def common_decimal(p1, s1, p2, s2):
    """coerce_decimals (llkv-compute/src/kernels.rs:179-242): scale max(s1, s2), integer digits max(p1 - s1, p2 - s2)."""
    s = max(s1, s2)
    return min(38, max(p1 - s1, p2 - s2) + s), s


def decimal_cases():
    cases = []

    def mul_case(name, p, s, pairs, note):
        rp, rs = min(38, 2 * p + 1), 2 * s       # numeric::mul result type
        tp, ts = common_decimal(p, s, p, s)       # preferred type the evaluator casts back to: the operands' common type
        a = pa.array([dec(x, s) for x, _ in pairs], pa.decimal128(p, s))
        b = pa.array([dec(y, s) for _, y in pairs], pa.decimal128(p, s))
        prod = pc.multiply(a, b)
        assert prod.type == pa.decimal128(rp, rs), (prod.type, rp, rs)
        expect = []
        for (x, y), pv in zip(pairs, prod):
            raw = x * y
            assert int(pv.as_py().scaleb(rs)) == raw
            r = rescale_half_away(raw, rs, ts)
            assert r == pa_round_half_away(raw, rp, rs, ts), (x, y, r)
            expect.append(r if fits(r, tp) else None)
        cases.append({"name": name, "source": "arrow-arith numeric::mul -> Decimal128(min(38, p1+p2+1), s1+s2); llkv-compute/src/eval.rs:587-591 casts to the "
                                              "preferred type; arrow-cast decimal->decimal rounds half away from zero, safe cast: precision overflow -> NULL",
                      "rule": note, "columns": [{"field": 1, "type": "Decimal128", "precision": p, "scale": s, "values": [x for x, _ in pairs]},
                                                {"field": 2, "type": "Decimal128", "precision": p, "scale": s, "values": [y for _, y in pairs]}],
                      "expr": {"bin": [{"col": 1}, "mul", {"col": 2}]}, "intermediate_type": [rp, rs],
                      "spec_type": {"type": "Decimal128", "precision": tp, "scale": ts}, "expect": expect})

    mul_case("decimal_mul_rescale_rounds_half_away_from_zero", 15, 2,
             [(12345, 5), (105, 50), (-105, 50), (1, 49), (1, 50), (-1, 50), (1, -50), (-1, -50), (999, 15), (-999, 15), (210000000, 10), (0, 7)],
             "(15,2) x (15,2) -> (31,4) -> (15,2): 1.05 x 0.50 = 0.5250 -> 0.53, -1.05 x 0.50 -> -0.53, 0.01 x 0.49 -> 0.00, 0.01 x 0.50 -> 0.01")
    mul_case("decimal_mul_precision_overflow_is_null_under_safe_cast", 5, 2,
             [(99999, 99999), (-99999, 99999), (31623, 31622), (31623, 31623), (100, 99999), (10000, 1000)],
             "(5,2) x (5,2) -> (11,4) -> (5,2): 999.99 x 999.99 = 999980.0001 needs 8 digits at scale 2 -> NULL; 316.23 x 316.22 = 99998.2... -> NULL, "
             "1.00 x 999.99 fits")

    def addsub_case(name, op, t1, t2, pairs, note):
        (p1, s1), (p2, s2) = t1, t2
        tp, ts = common_decimal(p1, s1, p2, s2)
        # both operands are cast to the common type first (compute_binary), then add/sub types the result from those
        rp, rs = min(38, max(tp - ts, tp - ts) + ts + 1), ts
        a = pa.array([dec(x, s1) for x, _ in pairs], pa.decimal128(p1, s1))
        b = pa.array([dec(y, s2) for _, y in pairs], pa.decimal128(p2, s2))
        res = (pc.add if op == "add" else pc.subtract)(a.cast(pa.decimal128(tp, ts)), b.cast(pa.decimal128(tp, ts)))
        assert res.type == pa.decimal128(rp, rs), (res.type, rp, rs)
        expect = []
        for (x, y), rv in zip(pairs, res):
            xa, ya = x * 10 ** (ts - s1), y * 10 ** (ts - s2)
            raw = xa + ya if op == "add" else xa - ya
            assert int(rv.as_py().scaleb(rs)) == raw
            expect.append(raw if fits(raw, tp) else None)
        cases.append({"name": name, "source": "arrow-arith numeric::add/sub -> Decimal128(min(38, max(s1,s2) + max(p1-s1, p2-s2) + 1), max(s1,s2)); "
                                              "llkv-compute/src/kernels.rs:179-242 coerce_decimals; eval.rs:587-591 cast back to the common type (safe)",
                      "rule": note, "columns": [{"field": 1, "type": "Decimal128", "precision": p1, "scale": s1, "values": [x for x, _ in pairs]},
                                                {"field": 2, "type": "Decimal128", "precision": p2, "scale": s2, "values": [y for _, y in pairs]}],
                      "expr": {"bin": [{"col": 1}, op, {"col": 2}]}, "intermediate_type": [rp, rs],
                      "spec_type": {"type": "Decimal128", "precision": tp, "scale": ts}, "expect": expect})

    addsub_case("decimal_add_aligns_scales_and_carries_one_digit", "add", (5, 2), (7, 4), [(123, 45678), (99999, 9999999), (-99999, -9999999), (99999, 1), (-500, 5000000)],
                "(5,2) + (7,4): common type (7,4), sum typed (8,4); 999.99 + 999.9999 = 1999.9899 needs 8 digits -> NULL after the cast back to (7,4)")
    addsub_case("decimal_sub_aligns_scales", "sub", (7, 4), (5, 2), [(45678, 123), (-9999999, 99999), (9999999, -99999), (10000, 100)],
                "(7,4) - (5,2): -999.9999 - 999.99 = -1999.9899 -> NULL at (7,4); 1.0000 - 1.00 = 0")
    return cases


def integer_cases():
    def col(field, values):
        return {"field": field, "type": "Int64", "values": values}

    cases = []
    for name, op, pairs, checked in (
            ("int64_add_overflow_is_an_error", "add", [(I64_MAX, 1)], pc.add_checked),
            ("int64_sub_overflow_is_an_error", "sub", [(I64_MIN, 1)], pc.subtract_checked),
            ("int64_mul_overflow_is_an_error", "mul", [(2**62, 2)], pc.multiply_checked)):
        try:
            checked(pa.array([pairs[0][0]], pa.int64()), pa.array([pairs[0][1]], pa.int64()))
            raise AssertionError("expected an overflow")
        except pa.ArrowInvalid:
            pass
        cases.append({"name": name, "source": "arrow-arith numeric::{add,sub,mul} are checked: ArrowError::ArithmeticOverflow; "
                                              "llkv-compute/src/kernels.rs:112-120 maps it to Error::Internal",
                      "rule": "integer overflow is an error, never a wrapped value", "columns": [col(1, [a for a, _ in pairs]), col(2, [b for _, b in pairs])],
                      "expr": {"bin": [{"col": 1}, op, {"col": 2}]}, "spec_type": {"type": "Int64"},
                      "expect_error": {"code": 8, "contains": "verflow"}})
    edge = [(I64_MAX, 0), (I64_MAX - 1, 1), (I64_MIN, 0), (-1, I64_MIN + 1), (3037000499, 3037000499), (-3037000499, 3037000499)]
    for op, f in (("add", lambda a, b: a + b), ("mul", lambda a, b: a * b)):
        pairs = [(a, b) for a, b in edge if I64_MIN <= f(a, b) <= I64_MAX]
        cases.append({"name": f"int64_{op}_at_the_edge_of_the_range", "source": "arrow-arith checked kernels: exact results inside i64",
                      "rule": "values up to the limits pass", "columns": [col(1, [a for a, _ in pairs]), col(2, [b for _, b in pairs])],
                      "expr": {"bin": [{"col": 1}, op, {"col": 2}]}, "spec_type": {"type": "Int64"}, "expect": [f(a, b) for a, b in pairs]})
    divs = [(7, 2), (-7, 2), (7, -2), (5, 0), (0, 0), (I64_MIN, 1), (I64_MAX, -1)]

    def trunc_div(a, b):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q

    cases.append({"name": "int64_div_truncates_and_zero_divisors_give_null", "source": "llkv-compute/src/kernels.rs:121-135: zeros of the divisor become NULL "
                                                                                       "(nullif) before numeric::div; Rust / arrow integer division truncates toward zero",
                  "rule": "7/2 = 3, -7/2 = -3, x/0 = NULL", "columns": [col(1, [a for a, _ in divs]), col(2, [b for _, b in divs])],
                  "expr": {"bin": [{"col": 1}, "div", {"col": 2}]}, "spec_type": {"type": "Int64"},
                  "expect": [None if b == 0 else trunc_div(a, b) for a, b in divs]})
    cases.append({"name": "int64_min_div_minus_one_overflows", "source": "arrow-arith numeric::div is checked: i64::MIN / -1 -> ArithmeticOverflow -> Error::Internal",
                  "rule": "the one quotient that does not fit", "columns": [col(1, [I64_MIN]), col(2, [-1])],
                  "expr": {"bin": [{"col": 1}, "div", {"col": 2}]}, "spec_type": {"type": "Int64"}, "expect_error": {"code": 8, "contains": "verflow"}})
    cases.append({"name": "int64_rem_by_zero_is_an_error", "source": "llkv-compute/src/kernels.rs:136-138: numeric::rem without the nullif -> ArrowError::DivideByZero "
                                                                     "-> Error::Internal",
                  "rule": "only Divide turns zero divisors into NULL", "columns": [col(1, [5]), col(2, [0])],
                  "expr": {"bin": [{"col": 1}, "mod", {"col": 2}]}, "spec_type": {"type": "Int64"}, "expect_error": {"code": 8, "contains": "ivide by zero"}})
    rems = [(7, 2), (-7, 2), (7, -2), (I64_MIN, -1)]

    def trunc_rem(a, b):
        return a - b * trunc_div(a, b)

    cases.append({"name": "int64_rem_takes_the_sign_of_the_dividend", "source": "Rust `%` / arrow-arith numeric::rem (wrapping for i64::MIN % -1 = 0)",
                  "rule": "-7 % 2 = -1, 7 % -2 = 1", "columns": [col(1, [a for a, _ in rems]), col(2, [b for _, b in rems])],
                  "expr": {"bin": [{"col": 1}, "mod", {"col": 2}]}, "spec_type": {"type": "Int64"}, "expect": [trunc_rem(a, b) for a, b in rems]})
    return cases


def main():
    doc = {"_comment": __doc__.strip().splitlines()[0] + " See tests/golden/make_arrow_golden.py for how every value was derived and cross-checked "
                       "(Python decimal + pyarrow %s); arrow-rs itself is not available offline." % pa.__version__,
           "cases": decimal_cases() + integer_cases()}
    path = os.path.join(HERE, "arrow_known_answers.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")
    print("wrote", path, len(doc["cases"]), "cases")


if __name__ == "__main__":
    main()
