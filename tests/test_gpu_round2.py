"""GPU: round-2 additions to the boundary — host-narrowed uploads from page-locked memory, resident i32 decimals,
llkv_gpu_agg_execute (one call per step, replayed as a CUDA graph), llkv_gpu_column_flush, chunked Utf8 uploads — each
against the oracle / numpy on the same inputs, through the C ABI."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from llkv_b200 import ffi, tpch
from llkv_b200.expr import AggregateKind, AggregateSpec, DataType, ScalarExpr
from llkv_b200.table import HostColumn, HostTable, decimal_from_i64
from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEC = DataType.Decimal128(15, 2)


def pinned_decimal(gpu, values_i64: np.ndarray, hi_override=None):
    """Arrow Decimal128 values buffer (16 B per value) in page-locked memory."""
    wide = decimal_from_i64(values_i64)
    if hi_override is not None:
        for i, hi in hi_override.items():
            wide.reshape(-1, 2)[i, 1] = np.uint64(hi)
    buf, ptr = gpu.pinned_empty(wide.nbytes)
    buf[:] = wide.view(np.uint8).reshape(-1)
    return buf.view(wide.dtype).reshape(wide.shape), ptr


@pytest.mark.parametrize("share", [0, 50, 100], ids=["workers", "hybrid", "dma"])
@pytest.mark.parametrize("case", ["fits_i32", "fits_i64", "late_i64_value", "late_wide_value"])
def test_host_narrowed_upload_from_pinned_memory(gpu_ctx, case, share):
    """Decimal128 chunks appended from page-locked memory are narrowed by the host workers before the DMA; a chunk that
    stops fitting sends the column back to the Arrow layout without losing a value (llkv_gpu.h: llkv_gpu_ctx_set_upload_threads)."""
    from llkv_b200 import gpu
    n, chunk = 1_300_000, 65_536  # three 8 MiB blocks of Arrow bytes: at 50 % the middle one takes the copy engine
    rng = np.random.default_rng(11)
    v = rng.integers(-2_000_000, 2_000_000, n, dtype=np.int64)
    hi = None
    if case == "fits_i64":
        v = v * 10_000_000
    if case == "late_i64_value":
        v[1_000_000] = 1 << 40  # (in the middle block: found by the workers or by the device, depending on the share)
    if case == "late_wide_value":
        hi = {1_000_000: 5}  # a value that needs more than 64 bits
    view, ptr = pinned_decimal(gpu, v, hi)
    gpu_ctx.set_upload_threads(4)
    gpu_ctx.set_dma_share(share)  # percent of the Arrow bytes (8 MiB blocks) that cross as they lie and are narrowed on the device
    # rows of the DMA share: blocks of 2^19 values, `share` percent of them
    blocks = [(b + 1) * share // 100 > b * share // 100 for b in range((n * 16 >> 23) + 1)]

    def dma_rows(lo_row=0, hi_row=n):
        return sum(min(chunk, hi_row - lo) for lo in range(lo_row, hi_row, chunk) if blocks[(lo * 16) >> 23])
    col = HostColumn(1, DEC, view)
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(7, 1), col)
    try:
        base = view.ctypes.data
        for lo in range(0, n, chunk):
            dc.append_raw(base + lo * 16, min(chunk, n - lo), lo)
        dc.seal()
        moved = dc.h2d_bytes()
        d = dma_rows()
        if case == "fits_i32":
            assert moved == (n - d) * 4 + d * 16
        elif case == "fits_i64":
            assert moved == (n - d) * 8 + d * 16
        else:  # narrowed first, then everything since the last seal again in the Arrow layout
            assert moved == (n - d) * 4 + d * 16 + n * 16
        back = dc.read()
        assert np.array_equal(back.reshape(-1), view.reshape(-1))
        if case != "late_wide_value":
            t = HostTable(7).add(HostColumn(1, DEC, np.array(view)))
            dt = gpu.DeviceTable(gpu_ctx, 7)
            dt.columns[1] = dc
            dt.n_rows = n
            specs = [AggregateSpec("s", AggregateKind.Sum(1, DEC)), AggregateSpec("mn", AggregateKind.Min(1, DEC)),
                     AggregateSpec("mx", AggregateKind.Max(1, DEC))]
            flt = tpch.between_filter(1, -1_000_000, 1_500_000)
            got = dt.aggregate(flt, specs)
            util.assert_same_result(got, oracle.aggregate(t, flt, specs))
        # the same column again after clear(): the width that failed is not tried twice
        dc.clear()
        before = dc.h2d_bytes()
        for lo in range(0, n, chunk):
            dc.append_raw(base + lo * 16, min(chunk, n - lo), lo)
        dc.seal()
        again = dc.h2d_bytes() - before
        w = {"fits_i32": 4, "fits_i64": 8, "late_i64_value": 8, "late_wide_value": 8}[case]
        assert again == (n - d) * w + d * 16 + (n * 16 if case == "late_wide_value" else 0)
        assert np.array_equal(dc.read().reshape(-1), view.reshape(-1))
    finally:
        dc.destroy()
        gpu_ctx.set_upload_threads(-1)
        gpu_ctx.set_dma_share(-1)
        gpu.pinned_free(ptr)


def test_flush_lets_the_caller_reuse_a_pinned_buffer(gpu_ctx):
    """llkv_gpu_column_flush: afterwards a page-locked source may be overwritten (the coalesced copy has been issued and
    has drained); every batch must arrive as it was when it was appended."""
    from llkv_b200 import gpu
    n = 100_000
    buf, ptr = gpu.pinned_empty(n * 8)
    view = buf.view(np.int64)
    dc = gpu.DeviceColumn(gpu_ctx, gpu.logical_field_id(8, 1), HostColumn(1, DataType.Int64, view))
    try:
        want = []
        for batch in range(4):
            view[:] = np.arange(n, dtype=np.int64) * (batch + 1)
            want.append(view.copy())
            dc.append_raw(view.ctypes.data, n, batch * n)
            dc.flush()
        dc.seal()
        assert np.array_equal(dc.read(), np.concatenate(want))
    finally:
        dc.destroy()
        gpu.pinned_free(ptr)


def test_chunked_utf8_upload_moves_each_byte_once(gpu_ctx):
    from llkv_b200 import gpu
    rng = np.random.default_rng(5)
    words = ["A", "N", "R", "ab", "xyz", "", "hello", "seven77"]
    vals = [words[i] for i in rng.integers(0, len(words), 50_000)]
    col = HostColumn.utf8(1, vals)
    t = HostTable(9).add(col).add(HostColumn(2, DataType.Int64, rng.integers(0, 100, len(vals), dtype=np.int64)))
    dt = gpu.DeviceTable.from_host(gpu_ctx, t, chunk_rows=4096)
    try:
        moved = dt.columns[1].h2d_bytes()
        n_chunks = (len(vals) + 4095) // 4096
        assert moved == col.aux.nbytes + (len(vals) + n_chunks) * 4  # the data once + the offsets of every chunk
        specs = [AggregateSpec("c", AggregateKind.CountStar()), AggregateSpec("s", AggregateKind.Sum(2, DataType.Int64))]
        got = dt.aggregate(None, specs, group_by=(1,), cardinality_hint=8)
        util.assert_same_result(got, oracle.aggregate(t, None, specs, None, (1,), group_capacity=64))
    finally:
        dt.destroy()


@pytest.mark.parametrize("query", ["q6", "q1"])
def test_execute_replays_a_captured_graph_with_the_same_result(gpu_ctx, query):
    """llkv_gpu_agg_execute = reset + run in one call; from the fourth unchanged step on it replays a CUDA graph."""
    from llkv_b200 import gpu
    table, snap = tpch.lineitem_table(300_000, seed=9, with_q1=True, with_mvcc=True)
    dt = gpu.DeviceTable.from_host(gpu_ctx, table)
    try:
        if query == "q6":
            flt, specs, keys, hint, sn = tpch.q6_filter(), tpch.q6_aggregates(), (), 0, None
        else:
            flt, specs, keys, hint, sn = tpch.q1_filter(), tpch.q1_aggregates(), tpch.Q1_GROUP_BY, 4, snap
        want = oracle.aggregate(table, flt, specs, sn, keys, group_capacity=16)
        dt.set_snapshot(sn)
        prog = gpu.Program(gpu_ctx, flt)
        agg = gpu.Aggregation(dt, specs, keys, cardinality_hint=hint)
        for step in range(8):
            agg.execute(prog, sn is not None)
            util.assert_same_result(agg.finalize(16), want)
        info = agg.run_info()
        assert info.used_jit_kernel == 1 and info.graph_replays >= 3, (info.used_jit_kernel, info.graph_replays)
        # a change of the snapshot drops the graph: the result follows the new snapshot
        if sn is not None:
            from llkv_b200.table import Snapshot
            sn2 = Snapshot(txn_id=snap.txn_id, snapshot_id=40, noncommitted=snap.noncommitted)
            dt.set_snapshot(sn2)
            agg.execute(prog, True)
            util.assert_same_result(agg.finalize(16), oracle.aggregate(table, flt, specs, sn2, keys, group_capacity=16))
        # ... and with graphs switched off the same steps give the same answer
        gpu_ctx.set_graphs(0)
        dt.set_snapshot(sn)
        for step in range(3):
            agg.execute(prog, sn is not None)
            util.assert_same_result(agg.finalize(16), want)
        agg.destroy()
        prog.destroy()
    finally:
        gpu_ctx.set_graphs(1)
        dt.destroy()


def test_resident_decimals_are_four_bytes_when_they_fit(gpu_ctx):
    """Decimal128(15,2) columns whose values fit i32 are resident as 4 bytes per row (Q6: 16 B/row instead of 52)."""
    from llkv_b200 import gpu
    table, _ = tpch.lineitem_table(200_000, seed=3, with_q1=False)
    dt = gpu.DeviceTable.from_host(gpu_ctx, table)
    try:
        prog = gpu.Program(gpu_ctx, tpch.q6_filter())
        agg = gpu.Aggregation(dt, tpch.q6_aggregates())
        agg.run(prog)
        got = agg.finalize(1)
        info = agg.run_info()
        assert info.algorithmic_bytes_per_row == 52 and info.physical_bytes_per_row == 16 and info.used_fast_kernel == 1
        util.assert_same_result(got, oracle.aggregate(table, tpch.q6_filter(), tpch.q6_aggregates()))
        agg.destroy()
        prog.destroy()
    finally:
        dt.destroy()


def test_multi_gpu_merge_matches_the_oracle():
    """tools/multi_gpu_check.py under torchrun on every visible GPU (>= 2): row-range shards, partial states merged over the
    NVLink peer mailboxes / NCCL, every rank's merged Q6, Q1 and high-cardinality results against the oracle."""
    from llkv_b200 import gpu
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs on the box")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, NCCL_DEBUG="WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-4000:]
    assert f"multi-GPU merge ok on {world} ranks" in out.stdout, out.stdout[-4000:]
