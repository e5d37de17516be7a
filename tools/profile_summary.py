"""Turn an `ncu --set full` report into the markdown summary kept under profiles/.

    python tools/profile_summary.py gpurun_out/r15_full.ncu-rep [launches.csv] > profiles/r01_....md

Per captured kernel (first launch of each name): duration, DRAM traffic, pipe utilisation, registers,
occupancy and the ten SASS instructions with the most warp-stall samples (needs -lineinfo builds).
Runs here on the CPU box (`ncu -i` only reads the report)."""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.avg", "sm cycles elapsed"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_not_selected",
          "stall_selected", "stall_barrier", "stall_branch_resolving", "stall_no_inst", "stall_lg", "stall_dispatch"]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200; "
          "per-launch values are cold-cache and serialised (compare shares, not absolutes).\n")
    if len(sys.argv) > 2:
        lines = [l for l in open(sys.argv[2]) if not l.startswith("==")]
        rows = list(csv.DictReader(lines))
        agg = collections.OrderedDict()
        for r in rows:
            name = r["Kernel Name"].split("(")[0].replace("void ", "")[:60]
            agg.setdefault(name, []).append(float(r["Metric Value"]))
        skip = ("distribution_elementwise", "vectorized_elementwise", "unrolled_elementwise")
        agg = collections.OrderedDict((k, v) for k, v in agg.items() if not any(s in k for s in skip))
        steps = len(max(agg.values(), key=sum))      # launches of the heaviest kernel: one per step
        tot = sum(sum(v) / steps for v in agg.values())
        print(f"## launch list (`{sys.argv[2].split('/')[-1]}`, gpu__time_duration.sum, {steps} steps)\n")
        print("| kernel | launches | mean us | us/step | share |\n|---|---:|---:|---:|---:|")
        for k, v in agg.items():
            print(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/steps/1e3:.2f} | {sum(v)/steps/tot:.1%} |")
        print(f"| sum per step | | | {tot/1e3:.2f} | |\n")
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    seen = set()
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        print(f"## `{name}`\n")
        print("| metric | value |\n|---|---:|")
        for key, label in RAW_KEYS:
            if key in d:
                print(f"| {label} (`{key}`) | {d[key]} {u.get(key, '')} |")
        print()
        src = ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{name.split('<')[0].split('::')[-1]}"])
        blocks, cur = [], None
        for row in csv.reader(io.StringIO(src)):
            if row and row[0] == "Kernel Name":
                cur = []
                blocks.append(cur)
            elif cur is not None:
                cur.append(row)
        if not blocks or len(blocks[0]) < 2:
            continue
        h = blocks[0][0]
        ix = {k: i for i, k in enumerate(h)}
        data = blocks[0][1:]
        if "# Samples" not in ix:
            continue
        tot = sum(int(x[ix["# Samples"]]) for x in data) or 1
        top = sorted(data, key=lambda x: -int(x[ix["# Samples"]]))[:10]
        print(f"top stall sites ({tot} samples, {len(data)} SASS instructions):\n")
        print("| samples | share | SASS | stall reasons |\n|---:|---:|---|---|")
        for x in top:
            s = int(x[ix["# Samples"]])
            why = " ".join(f"{k[6:]}={x[ix[k]]}" for k in STALLS if k in ix and int(x[ix[k]]) > 0)
            print(f"| {s} | {s/tot:.1%} | `{x[ix['Source']].strip()[:70]}` | {why} |")
        print()


if __name__ == "__main__":
    main()
