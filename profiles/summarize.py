"""Turns the ncu exports brought back in gpurun_out/ into the committed summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_rX.csv  > profiles/launches_rX.md
  python profiles/summarize.py raw      gpurun_out/raw_rX.csv       > profiles/ncu_full_rX.md
(raw_rX.csv = `ncu -i prof.ncu-rep --page raw --csv`)."""
import collections
import csv
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        k = row["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share | grid | block |\n|---|---:|---:|---:|---:|---|---|")
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% | {g} | {b} |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches "
          "(ncu serialises launches and runs them cold: compare shares, not absolutes)")


WANT = [
    ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes/instr"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed_op_global_red.sum", "global RED instr"),
]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"\n### `{r[idx['Kernel Name']].split('(')[0]}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        print("| metric | value |\n|---|---:|")
        for key, name in WANT:
            if key in idx:
                print(f"| {name} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
