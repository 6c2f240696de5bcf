"""Turns ncu artefacts (gpurun_out/, scratch) into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches <launches.csv>           per-kernel share of the step (gpu__time_duration)
    python tools/ncu_summary.py kernels  <report.ncu-rep>         one line of key counters per profiled launch
    python tools/ncu_summary.py source   <report.ncu-rep> <kernel substring> [n]   hottest source lines
"""
import collections, csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_%"),
        ("sm__cycles_active.avg", "sm_active_cyc"), ("sm__cycles_elapsed.max", "elapsed_cyc")]
STALLS = ["long_scoreboard", "wait", "no_instruction", "short_scoreboard", "math_pipe_throttle", "branch_resolving", "not_selected",
          "barrier", "lg_throttle", "mio_throttle"]


def short_name(full):
    """kernel name without its parameter list (template arguments such as <(int)-1, 1> kept)"""
    depth, cut = 0, len(full)
    for i, ch in enumerate(full):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return full[:cut].replace("void ", "").replace("qz::", "").replace("(int)", "").replace("(bool)", "").replace("<unnamed>::", "")


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = short_name(r[ik])
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total / 1e3:.2f} ms of kernel time (ncu: serialised, cold caches)")
    print(f"{'kernel':44s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:44s} {n:8d} {us:10.1f} {100 * us / total:6.1f}% {us / n:8.2f}")


def kernels(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    cols = [(hdr.index(k), label) for k, label in KEYS if k in hdr]
    print(f"# {rep}")
    print("kernel | " + " | ".join(label for _, label in cols) + " | top stalls (warps per issue)")
    for r in rows[2:]:
        name = short_name(r[ik])
        vals = []
        for i, label in cols:
            v = r[i]
            try:
                f = float(v)
                if label in ("dram_rd", "dram_wr"):
                    f = f * {"Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3}.get(units[i], 1.0)
                    v = f"{f:.1f}MB"
                elif label == "us":
                    f = f * {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3}.get(units[i], 1.0)
                    v = f"{f:.1f}"
                else:
                    v = f"{f:.1f}" if f < 1e5 else f"{f:.3g}"
            except ValueError:
                pass
            vals.append(v)
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in hdr and r[hdr.index(k)]:
                st.append((float(r[hdr.index(k)]), s))
        st = ", ".join(f"{s} {v:.2f}" for v, s in sorted(st, reverse=True)[:4])
        print(f"{name} | " + " | ".join(vals) + f" | {st}")


def source(rep, want, n=25):
    rows = ncu_csv(rep, "source", ("--print-source", "cuda,sass"))
    cur_file = cur_func = hdr = None
    agg = collections.defaultdict(lambda: [0, 0, 0, ""])
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1]
        elif len(r) == 2 and r[0] == "Function Name":
            cur_func = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and cur_func and want in cur_func and r[0]:
            d = dict(zip(hdr, r))
            try:
                inst, thr, smp = int(d["Instructions Executed"] or 0), int(d["Thread Instructions Executed"] or 0), int(d["# Samples"] or 0)
            except ValueError:
                continue
            a = agg[(cur_file.split("/")[-1], int(d["Line No"]))]
            a[0] += inst; a[1] += thr; a[2] += smp; a[3] = r[1].strip()[:80]
    tot, tots = sum(a[0] for a in agg.values()), sum(a[2] for a in agg.values())
    print(f"# {rep}: {want}: {tot} warp instructions, {sum(a[1] for a in agg.values()) / max(tot, 1):.1f} active threads per instruction, {tots} stall samples")
    print("file:line | % of instructions | threads/inst | % of samples | source")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(n)]:
        print(f"{k[0]}:{k[1]} | {100 * a[0] / max(tot, 1):.1f} | {a[1] / max(a[0], 1):.1f} | {100 * a[2] / max(tots, 1):.1f} | {a[3]}")


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels, "source": source}[sys.argv[1]](*sys.argv[2:])
