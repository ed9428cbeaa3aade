"""BASELINE.json configs[3] shape on one GPU, for profiling the partitioned GROUP BY: generates the two Int64 columns on the
device (key = SplitMix64(i) mod keys), runs the aggregate a few times and prints the kernel time.
  python tools/exp_highcard.py [rows] [keys] [partitioning mode]      (under ncu: add LLKV_GPU_NO_GRAPHS=1)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "rust-llkv_b200"), ROOT]
import numpy as np
import torch

from llkv_b200 import gpu, tpch
from llkv_b200.expr import DataType
from llkv_b200.table import HostColumn

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
keys = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = gpu.Context(0)
ctx.set_timing(True)
dev = torch.device("cuda", 0)


def shr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix(i):
    z = i + (-7046029254386353131)
    z = (z ^ shr(z, 30)) * (-4658895280553007687)
    z = (z ^ shr(z, 27)) * (-7723592293110705685)
    return z ^ shr(z, 31)


def umod(z, m):
    return (((z >> 1) & 0x7FFFFFFFFFFFFFFF) % m * 2 + (z & 1)) % m


dt = gpu.DeviceTable(ctx, 2)
cols = {}
piece = 1 << 26
for lo in range(0, rows, piece):
    m = min(piece, rows - lo)
    i = torch.arange(lo, lo + m, dtype=torch.int64, device=dev)
    k = umod(splitmix(i), keys)
    v = umod(splitmix(i + (1 << 40)), 1001)
    torch.cuda.synchronize()
    for fid, arr in ((tpch.K_FIELD, k), (tpch.V_FIELD, v)):
        if fid not in cols:
            cols[fid] = gpu.DeviceColumn(ctx, gpu.logical_field_id(2, fid), HostColumn(fid, DataType.Int64, np.zeros(0, np.int64)))
            cols[fid].reserve(rows)
            dt.columns[fid] = cols[fid]
        cols[fid].append_raw(arr.data_ptr(), m, lo)
        cols[fid].flush()
    del i, k, v
dt.n_rows = rows
dt.seal()
ctx.set_partitioning(mode)
if os.environ.get("TUNE"):
    ctas, block, stages, rpt = (int(x) for x in os.environ["TUNE"].split(","))
    ctx.set_tuning(ctas, block, stages, rpt, 0)
agg = gpu.Aggregation(dt, tpch.highcard_aggregates(), (tpch.K_FIELD,), cardinality_hint=keys)
for r in range(reps):
    agg.reset()
    agg.run(None, False, 0, rows)
    agg.reset()  # waits for the run
    info = agg.run_info()
    ctx.synchronize()
    agg2 = None
agg.reset()
agg.run(None, False, 0, rows)
g = agg.group_count()
info = agg.run_info()
print(f"rows={rows} keys={keys} mode={mode}: groups={g} kernel {info.last_kernel_ms:.3f} ms  {rows / info.last_kernel_ms / 1e6:.2f} Grows/s  "
      f"partitions={info.partitions} packed={info.packed_tuples} launches={info.kernel_launches} grid={info.grid} block={info.block} smem={info.smem_bytes}", flush=True)
