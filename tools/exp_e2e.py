"""Experiment: Q6 end-to-end step time (host Arrow buffers -> result) against the DMA share of the hybrid upload.
python tools/exp_e2e.py [shares...]   (percent; -1 = automatic)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-llkv_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from llkv_b200 import gpu, tpch  # noqa: E402
from llkv_b200.table import HostColumn  # noqa: E402


def main():
    shares = [int(x) for x in sys.argv[1:]] or [-1, 0, 20, 30, 40, 50, 60, 100]
    n = tpch.lineitem_rows(10)
    ctx = gpu.Context(0, n_streams=4, pinned_bytes=64 << 20)
    cores = os.cpu_count() or 1
    threads = int(os.environ.get("WORKERS", max(0, min(32, cores - 1))))
    ctx.set_upload_threads(threads)
    cols, dcs, ptrs = {}, {}, []
    fields = [tpch.L_QUANTITY, tpch.L_EXTENDEDPRICE, tpch.L_DISCOUNT, tpch.L_SHIPDATE]
    full, _ = tpch.lineitem_table(n, seed=6, with_q1=False)
    dt = gpu.DeviceTable(ctx, 1)
    for f in fields:
        c = full.columns[f]
        buf, ptr = gpu.pinned_empty(c.values.nbytes)
        buf[:] = c.values.view(np.uint8).reshape(-1)
        ptrs.append(ptr)
        cols[f] = HostColumn(f, c.dtype, buf.view(c.values.dtype).reshape(c.values.shape))
        dcs[f] = gpu.DeviceColumn(ctx, gpu.logical_field_id(1, f), cols[f])
        dt.columns[f] = dcs[f]
    dt.n_rows = n
    prog = gpu.Program(ctx, tpch.q6_filter())
    agg = gpu.Aggregation(dt, tpch.q6_aggregates())
    want = tpch.expected_q6(tpch.lineitem_arrays(n, 6, False))

    chunk_bytes = int(os.environ.get("CHUNK_BYTES", 1 << 20))
    phases = [0.0, 0.0, 0.0]

    def step():
        t0 = time.perf_counter()
        for f in fields:
            dcs[f].clear()
            bench.upload_column(dcs[f], cols[f], chunk_bytes, 0)
        t1 = time.perf_counter()
        for f in fields:
            dcs[f].seal()
        t2 = time.perf_counter()
        agg.execute(prog, False, 0, n)
        r = agg.finalize(1)
        t3 = time.perf_counter()
        phases[0] += t1 - t0
        phases[1] += t2 - t1
        phases[2] += t3 - t2
        return r

    print(f"cores {cores}, workers {threads}")
    for sh in shares:
        ctx.set_dma_share(sh)
        for _ in range(2):
            r = step()
        moved0 = sum(dcs[f].h2d_bytes() for f in fields)
        phases[:] = [0.0, 0.0, 0.0]
        t0 = time.perf_counter()
        k = 4
        for _ in range(k):
            r = step()
        dt_s = (time.perf_counter() - t0) / k
        moved = (sum(dcs[f].h2d_bytes() for f in fields) - moved0) / k
        assert r[0][1][0].value == want, (r[0][1][0].value, want)
        print(f"share {sh:4d}: {dt_s * 1e3:7.2f} ms/step  {n / dt_s / 1e9:5.2f} G rows/s  h2d {moved / 1e9:5.2f} GB  (result checked)  append {phases[0] / k * 1e3:.1f} seal {phases[1] / k * 1e3:.1f} run {phases[2] / k * 1e3:.1f} ms")


if __name__ == "__main__":
    main()
