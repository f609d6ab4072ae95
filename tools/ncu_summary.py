"""Summarise an .ncu-rep (read here, without a GPU) into profiles/<name>.md:
key raw metrics of each profiled launch plus the hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.md "title" """
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = ["# " + title, "", "source: `%s` (ncu --set full --clock-control none --import-source on)" % rep, ""]
    for ki, row in enumerate(raw[2:]):
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "kernel"
        lines += ["## launch %d: `%s`" % (ki, name[:120]), "", "| metric | value | unit |", "|---|---|---|"]
        for h, u, v in zip(hdr, units, row):
            if h in KEYS:
                lines.append("| %s | %s | %s |" % (h, v, u))
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    hi = next(i for i, r in enumerate(src[:10]) if "Source" in r)
    agg = collections.OrderedDict()
    stall_cols = [(j, h) for j, h in enumerate(src[hi]) if h.startswith("stall_") and "Not Issued" not in h]
    stalls = collections.Counter()
    for r in src[hi + 1:]:
        if r[0] != "" and len(r) > 8:
            try:
                inst, samp = int(r[7]), int(r[6])
            except ValueError:
                continue
            a = agg.setdefault((r[0], r[1].strip()[:90]), [0, 0])
            a[0] += inst
            a[1] += samp
            for j, h in stall_cols:
                try:
                    stalls[h] += int(r[j])
                except (ValueError, IndexError):
                    pass
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[1] for a in agg.values()) or 1
    lines += ["## hottest source lines (first profiled launch)", "",
              "warp instructions executed: %d, stall samples: %d" % (tot, tots), "",
              "| % inst | % samples | line | source |", "|---|---|---|---|"]
    for (ln, s), (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        lines.append("| %.1f | %.1f | %s | `%s` |" % (100 * inst / tot, 100 * samp / tots, ln, s.replace("|", "\\|")))
    st = sum(stalls.values()) or 1
    lines += ["", "## warp stall reasons (all samples)", "", "| reason | % |", "|---|---|"]
    for h, c in stalls.most_common(8):
        lines.append("| %s | %.1f |" % (h, 100 * c / st))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
