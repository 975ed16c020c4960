#!/usr/bin/env python
"""Condense ncu artefacts brought back in gpurun_out/ into the small text files committed under profiles/.

    python profiles/summarize.py launches <launches.csv> <out.txt> [title]
    python profiles/summarize.py full <prof.ncu-rep | prof.raw.csv> <out.txt> [title]

`launches`: per-kernel count / total device time / share of the step from the
`ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list.
`full`: the headline metrics of every kernel in one `ncu --set full` report (read with
`ncu -i ... --page raw --csv`)."""
import collections
import csv
import io
import re
import subprocess
import sys

FULL_KEYS = [
    r"^gpu__time_duration\.sum$", r"^launch__grid_size$", r"^launch__block_size$", r"^launch__registers_per_thread$",
    r"^launch__shared_mem_per_block_dynamic$", r"^sm__cycles_elapsed\.max$", r"^sm__cycles_active\.avg$",
    r"^gpc__cycles_elapsed\.avg\.per_second$",
    r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__inst_executed_pipe_tensor.*hmma\.avg\.pct_of_peak_sustained_active$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_bytes\.sum$", r"^lts__t_sectors_srcunit_tex\.sum$", r"^lts__t_sector_hit_rate\.pct$",
    r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_(ld|st)\.sum$",
    r"^smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$",
    r"^smsp__average_warp_latency_issue_stalled_.*\.ratio$",
]


def launches(path, out, title):
    rows = [l for l in open(path) if not l.startswith("==")]
    r = list(csv.DictReader(io.StringIO("".join(rows))))
    agg = collections.OrderedDict()
    for x in r:
        k = x["Kernel Name"]
        v = float(x["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(x["Metric Unit"], 1.0)
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = min(a[2], v)
        a[3] = max(a[3], v)
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch "
                f"times are cold-cache and serialised: compare SHARES)\n# launches={len(r)} total_device_time_ms={tot / 1e6:.3f}\n")
        f.write(f"{'share%':>7s} {'n':>5s} {'total_us':>12s} {'min_us':>10s} {'max_us':>10s}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / tot * 100:7.2f} {v[0]:5d} {v[1] / 1e3:12.1f} {v[2] / 1e3:10.1f} {v[3] / 1e3:10.1f}  {k[:110]}\n")


def full(path, out, title):
    if path.endswith(".ncu-rep"):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True,
                             check=True).stdout
    else:
        txt = open(path).read()
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(p) for p in FULL_KEYS]
    cols = [i for i, h in enumerate(hdr) if any(p.search(h) for p in pats)]
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: {path} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}  (launch id {r[hdr.index('ID')]})\n")
            for i in cols:
                if r[i] not in ("", "n/a"):
                    f.write(f"  {hdr[i]:88s} {r[i]:>18s} {units[i]}\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
