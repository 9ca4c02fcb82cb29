#!/usr/bin/env python
"""Turn `ncu -i <rep> --page raw --csv` into (a) a readable per-kernel summary for profiles/ and (b) the record
bench.py reads for `roofline.traffic` (profiles/ncu_gemm1.json), stamped with the hash of the kernel sources so that a
later change to the kernel invalidates the number instead of leaving it stale.

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv
  python scripts/ncu_extract.py gpurun_out/prof_raw.csv --summary profiles/r02_ncu_step.txt \
      --record profiles/ncu_gemm1.json --record-kernel gemm_tc2_kernel --record-index 0 --streams 4096 --chunk 64 --precision fp16
"""
import argparse
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["prego_b200/csrc/gemm_tc.cuh", "prego_b200/csrc/ptx.cuh", "prego_b200/csrc/api.cu"]
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "sm__cycles_elapsed.max",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0}


def load(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {}
    for i, n in enumerate(names):
        col.setdefault(n.split(".TriageCompute.")[-1] if ".TriageCompute." in n else n, i)
        col.setdefault(n, i)
    return names, units, col, rows[hdr + 2:]


def value(row, units, col, key):
    i = col.get(key)
    if i is None or i >= len(row) or row[i] in ("", "n/a", "no data"):
        return None, None
    try:
        return float(row[i].replace(",", "")), units[i]
    except ValueError:
        return row[i], units[i]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--summary")
    ap.add_argument("--title", default="")
    ap.add_argument("--record")
    ap.add_argument("--record-kernel", default="gemm_tc2_kernel")
    ap.add_argument("--record-index", type=int, default=0, help="which launch of that kernel in the capture (0 = first = GEMM1)")
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    names, units, col, rows = load(args.csv)
    kcol = col["Kernel Name"]
    lines = [args.title] if args.title else []
    seen = {}
    for r in rows:
        if len(r) <= kcol:
            continue
        kname = r[kcol]
        lines.append("  " + kname[:110])
        for k in KEYS:
            v, u = value(r, units, col, k)
            if v is not None:
                lines.append(f"    {k:<86}{v if isinstance(v, str) else format(v, 'f')} {u}")
        lines.append("")
        short = kname.split("(")[0].split("<")[0].split()[-1]
        idx = seen.get(short, 0)
        seen[short] = idx + 1
        if args.record and args.record_kernel in kname and idx == args.record_index:
            rd, ru = value(r, units, col, "dram__bytes_read.sum")
            wr, wu = value(r, units, col, "dram__bytes_write.sum")
            tp, _ = value(r, units, col, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")
            dur, du = value(r, units, col, "gpu__time_duration.sum")
            h = hashlib.sha256()
            for f in SOURCES:
                h.update(open(os.path.join(ROOT, f), "rb").read())
            rec = {"kernel": kname[:120], "capture": os.path.relpath(args.summary or args.csv, ROOT), "streams": args.streams, "chunk": args.chunk,
                   "precision": args.precision, "dram_bytes_read": rd * UNIT_SCALE.get(ru, 1.0), "dram_bytes_write": wr * UNIT_SCALE.get(wu, 1.0),
                   "tensor_pipe_active_pct": tp, "duration": f"{dur} {du}", "sources": SOURCES, "sources_sha256": h.hexdigest()}
            json.dump(rec, open(args.record, "w"), indent=1)
            print("record:", json.dumps(rec)[:300])
    text = "\n".join(lines) + "\n"
    if args.summary:
        open(args.summary, "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
