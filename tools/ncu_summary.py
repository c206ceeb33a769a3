"""Turns gpurun_out/*.ncu-rep + the launch-list csv into the tracked summaries under profiles/."""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            d[h.split(".TriageCompute.")[-1] if ".TriageCompute." in h else h] = (v, u)
        res.append(d)
    return res


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("hvn::", "")
        tot[name] += v
        cnt[name] += 1
    return tot, cnt


def main(tag):
    out = ["# ncu summaries, round %s (B200, `--clock-control none`)" % tag, ""]
    lp = os.path.join(ROOT, "gpurun_out", "%s_launches.csv" % tag)
    if os.path.exists(lp):
        tot, cnt = launches(lp)
        T = sum(tot.values())
        out += ["## Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 "
                "--warmup 1 --batch 16 --no-cpu-baseline`", "",
                "Per-launch times under ncu are serialised and cold-cache: read the SHARES. Total %.1f ms over %d launches "
                "(warm-up step + timed step + e2e step + profile passes)." % (T, sum(cnt.values())), "",
                "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
        for k, v in sorted(tot.items(), key=lambda x: -x[1])[:18]:
            out.append("| `%s` | %d | %.3f | %.1f %% |" % (k[:70], cnt[k], v, 100 * v / T))
        out.append("")
    for f in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
        if not (f.startswith(tag + "_") and f.endswith(".ncu-rep")):
            continue
        for d in raw(os.path.join(ROOT, "gpurun_out", f)):
            out += ["## `%s` — %s" % (f, d.get("Kernel Name", ("?", ""))[0][:90]), "", "| metric | value |", "|---|---|"]
            for k in KEYS:
                if k in d:
                    out.append("| %s | %s %s |" % (k, d[k][0], d[k][1]))
            out.append("")
    path = os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % tag)
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
