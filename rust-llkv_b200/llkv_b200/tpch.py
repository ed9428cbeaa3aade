"""Synthetic workloads of BASELINE.json: lineitem (TPC-H shaped), the single-Int64 column, the high-cardinality table —
and the plans llkv-sql hands the executor for them.

`tpchgen` (llkv-tpch/src/lib.rs:27-30,363) is not vendored, so the generator is ours: seeded, vectorised numpy, value
ranges of the TPC-H spec.  Column dtypes are what llkv-plan gives SQL tables
(llkv-plan/src/translation/types.rs:22-44): DECIMAL(15,2) -> Decimal128(15,2), DATE -> Date32, CHAR(1) -> Utf8,
integers -> Int64; MVCC columns UInt64 (llkv-transaction/src/mvcc.rs:475-481).
"""
from __future__ import annotations

import datetime
from typing import List, Optional, Tuple

import numpy as np

from . import ffi
from .expr import (AggregateKind, AggregateSpec, Bound, DataType, Expr, Literal, Operator, ScalarExpr, pred)
from .table import HostColumn, HostTable, Snapshot, decimal_from_i64

ROWS_PER_SF = 6_000_000  # lineitem is ~6.0 M rows per scale factor (SF1 6 001 215)
LINEITEM_ROWS = {1: 6_001_215, 10: 59_986_052, 100: 600_037_902}

# FieldIds of the lineitem columns this path reads
L_QUANTITY, L_EXTENDEDPRICE, L_DISCOUNT, L_TAX, L_RETURNFLAG, L_LINESTATUS, L_SHIPDATE = 5, 6, 7, 8, 9, 10, 11
DEC_15_2 = DataType.Decimal128(15, 2)

TXN_ID_NONE = (1 << 64) - 1
TXN_ID_AUTO_COMMIT = 1


def date32(y: int, m: int, d: int) -> int:
    return (datetime.date(y, m, d) - datetime.date(1970, 1, 1)).days


def lineitem_rows(sf: float) -> int:
    return LINEITEM_ROWS.get(int(sf), int(sf * ROWS_PER_SF)) if sf == int(sf) else int(sf * ROWS_PER_SF)


def lineitem_arrays(n: int, seed: int = 1, with_q1: bool = True):
    """Raw numpy arrays of the lineitem columns (int64 raw decimals, int32 dates, uint8 flag bytes)."""
    rng = np.random.default_rng(seed)
    qty = rng.integers(1, 51, n, dtype=np.int64)
    part_price = rng.integers(90_000, 210_001, n, dtype=np.int64)  # 900.00 .. 2100.00
    out = {
        "quantity": qty * 100,
        "extendedprice": qty * part_price,
        "discount": rng.integers(0, 11, n, dtype=np.int64),
        "shipdate": rng.integers(date32(1992, 1, 2), date32(1998, 12, 1) + 1, n, dtype=np.int64).astype(np.int32),
    }
    if with_q1:
        out["tax"] = rng.integers(0, 9, n, dtype=np.int64)
        cutoff = date32(1995, 6, 17)
        ship = out["shipdate"]
        coin = rng.integers(0, 2, n, dtype=np.int64)
        flag = np.where(ship <= cutoff, np.where(coin == 0, ord("A"), ord("R")), ord("N")).astype(np.uint8)
        # the small N/F group of TPC-H: shipped just before the cutoff, received after it
        nf = (ship > cutoff - 30) & (ship <= cutoff) & (coin == 1)
        flag[nf] = ord("N")
        out["returnflag"] = flag
        out["linestatus"] = np.where(ship <= cutoff, ord("F"), ord("O")).astype(np.uint8)
    return out


def _utf8_single_char(field_id: int, codes: np.ndarray) -> HostColumn:
    n = codes.shape[0]
    return HostColumn(field_id, DataType.Utf8, np.arange(n + 1, dtype=np.int32), aux=np.ascontiguousarray(codes, dtype=np.uint8))


def mvcc_arrays(n: int, seed: int, snapshot_id: int = 100, active_txn: int = 77):
    """created_by / deleted_by for config 3: 1 % of the rows deleted by a committed txn <= snapshot, 0.5 % created
    after the snapshot, a few touched by one Active transaction."""
    rng = np.random.default_rng(seed + 1000)
    created = np.full(n, TXN_ID_AUTO_COMMIT, dtype=np.uint64)
    deleted = np.full(n, TXN_ID_NONE, dtype=np.uint64)
    u = rng.random(n)
    deleted[u < 0.01] = np.uint64(50)
    created[(u >= 0.01) & (u < 0.015)] = np.uint64(snapshot_id + 5)
    created[(u >= 0.015) & (u < 0.016)] = np.uint64(active_txn)
    deleted[(u >= 0.016) & (u < 0.017)] = np.uint64(active_txn)
    return created, deleted, Snapshot(txn_id=snapshot_id + 1, snapshot_id=snapshot_id, noncommitted=(active_txn, snapshot_id + 1))


def lineitem_table(n: int, seed: int = 1, with_q1: bool = True, with_mvcc: bool = False, table_id: int = 1):
    a = lineitem_arrays(n, seed, with_q1)
    t = HostTable(table_id)
    t.add(HostColumn(L_QUANTITY, DEC_15_2, decimal_from_i64(a["quantity"])))
    t.add(HostColumn(L_EXTENDEDPRICE, DEC_15_2, decimal_from_i64(a["extendedprice"])))
    t.add(HostColumn(L_DISCOUNT, DEC_15_2, decimal_from_i64(a["discount"])))
    t.add(HostColumn(L_SHIPDATE, DataType.Date32, a["shipdate"]))
    if with_q1:
        t.add(HostColumn(L_TAX, DEC_15_2, decimal_from_i64(a["tax"])))
        t.add(_utf8_single_char(L_RETURNFLAG, a["returnflag"]))
        t.add(_utf8_single_char(L_LINESTATUS, a["linestatus"]))
    snap = None
    if with_mvcc:
        c, d, snap = mvcc_arrays(n, seed)
        t.add_mvcc(c, d)
    return t, snap


# ------------------------------------------------------------------------------------------------ plans
def q6_filter() -> Expr:
    """l_shipdate >= DATE '1994-01-01' AND l_shipdate < DATE '1995-01-01' AND l_discount BETWEEN 0.05 AND 0.07 AND
    l_quantity < 24 — BETWEEN lowers to two comparisons (llkv-sql/src/sql_engine.rs:8818-8868), same-field ANDs fuse
    (llkv-compute/src/program.rs:415-439)."""
    return Expr.And([
        pred(L_SHIPDATE, Operator.Range(Bound.Included(Literal.Date32(date32(1994, 1, 1))), Bound.Excluded(Literal.Date32(date32(1995, 1, 1))))),
        pred(L_DISCOUNT, Operator.Range(Bound.Included(Literal.Decimal128(5, 2)), Bound.Included(Literal.Decimal128(7, 2)))),
        pred(L_QUANTITY, Operator.LessThan(Literal.Int128(24))),
    ])


def q6_aggregates() -> List[AggregateSpec]:
    """SUM(l_extendedprice * l_discount): arrow-mode product cast back to Decimal128(15,2) (SURVEY.md §8a note D1)."""
    e = ScalarExpr.Column(L_EXTENDEDPRICE) * ScalarExpr.Column(L_DISCOUNT)
    return [AggregateSpec("revenue", AggregateKind.Sum(e, DEC_15_2))]


def q1_filter() -> Expr:
    """l_shipdate <= DATE '1998-12-01' - INTERVAL '90' DAY (constant-folded to 1998-09-02)."""
    return pred(L_SHIPDATE, Operator.LessThanOrEquals(Literal.Date32(date32(1998, 9, 2))))


def q1_aggregates() -> List[AggregateSpec]:
    """TPC-H Q1 select list; GROUP BY expressions run in exact decimal mode (SURVEY.md §8a note D2)."""
    qty, price, disc, tax = (ScalarExpr.Column(c) for c in (L_QUANTITY, L_EXTENDEDPRICE, L_DISCOUNT, L_TAX))
    one = ScalarExpr.Literal(1)
    disc_price = price * (one - disc)
    charge = disc_price * (one + tax)
    d2, d4, d6 = DataType.Decimal128(38, 2), DataType.Decimal128(38, 4), DataType.Decimal128(38, 6)
    return [
        AggregateSpec("sum_qty", AggregateKind.Sum(qty, DEC_15_2)),
        AggregateSpec("sum_base_price", AggregateKind.Sum(price, DEC_15_2)),
        AggregateSpec("sum_disc_price", AggregateKind.Sum(disc_price, d4)),
        AggregateSpec("sum_charge", AggregateKind.Sum(charge, d6)),
        AggregateSpec("avg_qty", AggregateKind.Avg(qty, DEC_15_2)),
        AggregateSpec("avg_price", AggregateKind.Avg(price, DEC_15_2)),
        AggregateSpec("avg_disc", AggregateKind.Avg(disc, d2)),
        AggregateSpec("count_order", AggregateKind.CountStar()),
    ]


Q1_GROUP_BY = (L_RETURNFLAG, L_LINESTATUS)


# config 1: SELECT SUM(x) FROM t WHERE x BETWEEN a AND b
X_FIELD = 1


def int64_table(n: int, seed: int = 1, with_mvcc: bool = True, table_id: int = 1):
    rng = np.random.default_rng(seed)
    x = rng.integers(-10**9, 10**9 + 1, n, dtype=np.int64)
    t = HostTable(table_id).add(HostColumn(X_FIELD, DataType.Int64, x))
    snap = None
    if with_mvcc:
        t.add_mvcc(np.full(n, TXN_ID_AUTO_COMMIT, np.uint64), np.full(n, TXN_ID_NONE, np.uint64))
        snap = Snapshot(txn_id=TXN_ID_AUTO_COMMIT, snapshot_id=TXN_ID_AUTO_COMMIT)
    return t, snap


def between_filter(field: int, a: int, b: int) -> Expr:
    """x BETWEEN a AND b -> Expr::And[Pred(x >= a), Pred(x <= b)] (llkv-sql/src/sql_engine.rs:8818-8868)."""
    return Expr.And([pred(field, Operator.GreaterThanOrEquals(a)), pred(field, Operator.LessThanOrEquals(b))])


def sum_int64(field: int) -> List[AggregateSpec]:
    return [AggregateSpec("sum", AggregateKind.Sum(field, DataType.Int64))]


# config 4: high-cardinality GROUP BY
K_FIELD, V_FIELD = 1, 2


def highcard_table(n: int, n_keys: int, seed: int = 4, table_id: int = 1) -> HostTable:
    rng = np.random.default_rng(seed)
    k = rng.integers(0, n_keys, n, dtype=np.int64)
    v = rng.integers(0, 1001, n, dtype=np.int64)
    return HostTable(table_id).add(HostColumn(K_FIELD, DataType.Int64, k)).add(HostColumn(V_FIELD, DataType.Int64, v))


def highcard_aggregates() -> List[AggregateSpec]:
    return [AggregateSpec("s", AggregateKind.Sum(V_FIELD, DataType.Int64)), AggregateSpec("c", AggregateKind.CountStar())]


def shard_range(n_rows: int, world_size: int, rank: int, align: int = 131072) -> Tuple[int, int]:
    """Row range of `rank`: contiguous, aligned to chunk boundaries so every column of a shard covers the same rows
    (SURVEY.md §8e)."""
    chunks = (n_rows + align - 1) // align
    lo = chunks * rank // world_size * align
    hi = chunks * (rank + 1) // world_size * align
    return min(lo, n_rows), min(hi, n_rows)


# ------------------------------------------------------------------------------------------------ expected answers
# Plain numpy restatements of the two queries over the generator's arrays: what bench.py and the full-size tests check the
# device results against (exact integers; not the oracle, which replays the reference's algorithm row by row).
def visible_mask(created: np.ndarray, deleted: np.ndarray, snap: Snapshot) -> np.ndarray:
    """RowVersion::is_visible_for (llkv-transaction/src/mvcc.rs:282-334) with TxnIdManager::status (mvcc.rs:157-171:
    MAX -> none, 1 -> committed, listed -> not committed, unknown -> committed), vectorised."""
    txn, sid = np.uint64(snap.txn_id), np.uint64(snap.snapshot_id)
    none = np.uint64(TXN_ID_NONE)
    nc = np.asarray(sorted(set(int(x) for x in snap.noncommitted) - {TXN_ID_AUTO_COMMIT}), dtype=np.uint64)
    c_comm = (created != none) & ~np.isin(created, nc)
    d_comm = (deleted != none) & ~np.isin(deleted, nc)
    own = snap.txn_id != TXN_ID_AUTO_COMMIT
    own_c = (created == txn) if own else np.zeros(created.shape, bool)
    own_d = (deleted == txn) if own else np.zeros(created.shape, bool)
    others = c_comm & (created <= sid) & ((deleted == none) | (~own_d & (~d_comm | (deleted > sid))))
    return np.where(own_c, ~own_d, others)


def expected_q6(a) -> int:
    """Raw Decimal128(15,2) revenue: every product (scale 4) is rescaled to scale 2, half away from zero, then summed
    (SURVEY.md §8a note D1, "as written")."""
    m = ((a["shipdate"] >= date32(1994, 1, 1)) & (a["shipdate"] < date32(1995, 1, 1)) & (a["discount"] >= 5) & (a["discount"] <= 7)
         & (a["quantity"] < 2400))
    prod = a["extendedprice"][m] * a["discount"][m]
    return int(((prod + 50) // 100).sum())


def expected_q1_partials(a, created=None, deleted=None, snap: Optional[Snapshot] = None):
    """{(returnflag, linestatus): [sum_qty, sum_base_price, sum_disc_price, sum_charge, sum_disc, count, first_row]} as
    exact Python ints (partial states: add them across shards, then q1_rows_from_partials)."""
    sel = a["shipdate"] <= date32(1998, 9, 2)
    if snap is not None:
        sel &= visible_mask(created, deleted, snap)
    disc_price = a["extendedprice"] * (100 - a["discount"])  # scale 4, exact in int64
    charge = disc_price * (100 + a["tax"])                   # scale 6: < 1.4e11 per row
    out = {}
    code = a["returnflag"].astype(np.int64) * 256 + a["linestatus"]
    for c in np.unique(code[sel]):
        g = sel & (code == c)
        key = (chr(int(c) >> 8), chr(int(c) & 255))
        parts = np.nonzero(g)[0]
        # int64 sums stay exact: per group < 4e7 rows x 1.4e11
        out[key] = [int(a["quantity"][g].sum()), int(a["extendedprice"][g].sum()), int(disc_price[g].sum()), int(charge[g].sum()),
                    int(a["discount"][g].sum()), int(parts.size), int(parts[0])]
    return out


def add_partials(parts):
    out = {}
    for rank_base, p in parts:  # (row id of the shard's first row, partials)
        for k, v in p.items():
            if k not in out:
                out[k] = v[:6] + [v[6] + rank_base]
            else:
                cur = out[k]
                for i in range(6):
                    cur[i] += v[i]
                cur[6] = min(cur[6], v[6] + rank_base)
    return out


def q1_rows_from_partials(p):
    """Rows in first-appearance order, values in q1_aggregates() order (AVG = sum / count at the input scale, rounded half
    away from zero: llkv-aggregate/src/lib.rs:1720-1761)."""
    def avg(s, c):
        q, r = divmod(abs(s), c)
        q += 1 if 2 * r >= c else 0
        return q if s >= 0 else -q
    rows = []
    for k, v in sorted(p.items(), key=lambda kv: kv[1][6]):
        sq, sp, sdp, sc, sd, c, _ = v
        rows.append((k, [sq, sp, sdp, sc, avg(sq, c), avg(sp, c), avg(sd, c), c]))
    return rows


def splitmix64(i: np.ndarray) -> np.ndarray:
    """SplitMix64 of uint64 counters (the key stream of BASELINE.json configs[3]: key = SplitMix64(i) mod n_keys)."""
    with np.errstate(over="ignore"):
        z = i.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
