"""Per-instruction view of one launch of an `ncu --set full` capture: where the warp instructions and the stall samples are.

  python tools/ncu_hot.py <report.ncu-rep> [launch index, default last] [listing file]

Reads the source page of the report (`ncu -i … --page source --csv --print-source sass`: executed count and warp-stall
samples per SASS line) and prints
  * the instruction classes by execution count (every instruction of a straight-line region runs equally often, so a
    region shows up as "N instructions executed X times": the common path of a tile, a slow path taken by a share of the
    warps, the producer's loop), with their share of all issued instructions and of all stall samples;
  * the stall samples that sit on the first instruction behind a gap in the executed addresses of the hottest class
    (= behind a taken branch over code that did not run: instruction fetch and branch resolution);
  * the 25 SASS lines with the most samples.
With a listing file, every line is written as `offset executed samples sass` for reading next to `cuobjdump -sass`.
This is how the round-2 Q1 work was found (profiles/r02_summary.md, "Q1 after the late-round changes")."""
import csv
import subprocess
import sys
from collections import defaultdict


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tables, cur, header = [], None, None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1] if len(row) > 1 else "?", "rows": []}
            tables.append(cur)
            continue
        if cur is None or not row:
            continue
        if row[0] == "Address":
            header = {name: i for i, name in enumerate(row)}
            continue
        cur["rows"].append(row)
    return tables, header


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        return 1
    tables, header = load(sys.argv[1])
    if not tables:
        print("no source page in the report (was it captured with --set full?)")
        return 1
    which = int(sys.argv[2]) if len(sys.argv) > 2 else len(tables) - 1
    t = tables[which]
    i_exec, i_samp = header["Instructions Executed"], header["# Samples"]
    base = int(t["rows"][0][0], 16)
    lines = [(int(r[0], 16) - base, int(r[i_exec]), int(r[i_samp]), r[1].strip()) for r in t["rows"]]
    tot_i = sum(l[1] for l in lines) or 1
    tot_s = sum(l[2] for l in lines) or 1
    print(f"launch {which} of {len(tables)}: {t['name']}: {len(lines)} SASS lines, {tot_i} warp instructions, {tot_s} stall samples")
    classes = defaultdict(lambda: [0, 0])
    for _, ex, sm, _ in lines:
        c = classes[ex]
        c[0] += 1
        c[1] += sm
    print("\nexecuted x times | instructions | share of issued | share of samples")
    for ex, (n, sm) in sorted(classes.items(), key=lambda kv: -kv[0] * kv[1][0])[:14]:
        if ex:
            print(f"{ex:>16} | {n:>12} | {ex * n / tot_i:>15.3f} | {sm / tot_s:>16.3f}")
    hottest = max(classes.items(), key=lambda kv: kv[0] * kv[1][0])[0]
    band = [l for l in lines if 0.98 * hottest <= l[1] <= 1.02 * hottest]
    behind, prev = [], None
    for l in band:
        if prev is not None and l[0] - prev > 0x10:
            behind.append(l)
        prev = l[0]
    print(f"\ncommon path (executed ~{hottest} times): {len(band)} instructions, {sum(l[2] for l in band)} samples; "
          f"{sum(l[2] for l in behind)} of them on the {len(behind)} instructions right behind a taken branch:")
    for off, ex, sm, sass in behind:
        print(f"  {off:#07x} {sm:>6}  {sass}")
    print("\nmost-sampled lines:")
    for off, ex, sm, sass in sorted(lines, key=lambda l: -l[2])[:25]:
        print(f"  {off:#07x} x{ex:<9} {sm:>6}  {sass}")
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            for off, ex, sm, sass in lines:
                f.write(f"{off:05x} {ex:>10} {sm:>6} {sass}\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
