"""Turn raw captures in gpurun_out/ (ncu launch lists, .ncu-rep) into the tracked summaries under profiles/.

    python profiles/summarize.py launches <gpurun_out/x_launches.csv> <profiles/out.md> [exclude-regex]
    python profiles/summarize.py rep <gpurun_out/x.ncu-rep> <profiles/out.csv>
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def launches(src, dst, exclude=None):
    import re
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        k = row["Kernel Name"].split("(")[0]
        if exclude and re.search(exclude, k):
            continue
        tot[k] += v
        cnt[k] += 1
    s = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n")
        if exclude:
            f.write(f"launches matching /{exclude}/ left out\n")
        f.write("\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| `{k}` | {cnt[k]} | {v / 1e3:.3f} | {v / cnt[k]:.2f} | {v / s:.1%} |\n")
        f.write(f"\nsum of kernel time: {s / 1e3:.3f} ms over {sum(cnt.values())} launches\n")
    print(open(dst).read())


def rep(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](*sys.argv[2:])
