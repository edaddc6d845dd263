"""Summarise ncu outputs brought back in gpurun_out/ into tracked files under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.md> [title]
  python tools/summarize_ncu.py kernel <report.ncu-rep> <out.md> [title]
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are "
                "cold-cache and serialised: compare SHARES).\n\n| share | total ms | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| {100 * v[1] / tot:.1f}% | {v[1] / 1e6:.3f} | {v[0]} | `{k[:140]}` |\n")
        f.write(f"\nTotal: {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches.\n")


def kernel(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `{path}` (ncu --set full --clock-control none --import-source on).\n\n")
        for row in r[2:]:
            f.write(f"## {row[hdr.index('Kernel Name')][:120]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in RAW:
                if m in hdr:
                    f.write(f"| {m} | {row[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
            f.write("\n")


if __name__ == "__main__":
    mode, path, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else path
    {"launches": launches, "kernel": kernel}[mode](path, out, title)
