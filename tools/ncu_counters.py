"""ncu counters per unit of work -> profiles/r02_counters.json (read by bench.py for `roofline.traffic`, `l2_gbs`, `issue`).

    python tools/ncu_counters.py <scene> <metrics.csv> <profile_step.log> [<scene> <csv> <log> ...] > profiles/r02_counters.json

<metrics.csv> is the log of
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file <metrics.csv> python tools/profile_step.py --workload <w> --spp <n>
over a WHOLE (small) render with QZ_GRAPH=0; <profile_step.log> is that command's stdout (the render's ray / bounce counts).
Stages: traversal = k_trace_lane* + k_step_flat (per ray), shading = k_shade* + k_albedo_conductor (per shaded bounce),
sampler = k_sample + k_memo_* (per shaded bounce).
"""
import ast, collections, csv, json, re, sys

STAGES = {"traversal": ("k_trace_lane", "k_step_flat"), "shading": ("k_shade", "k_albedo_conductor"), "sampler": ("k_sample", "k_memo_")}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1.0, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def one(scene, csv_path, log_path):
    stats = None
    for line in open(log_path):
        m = re.match(r"^(\{.*\}) rc (\d+)", line.strip())
        if m:
            stats = ast.literal_eval(m.group(1))
    if stats is None:
        raise SystemExit(f"{log_path}: no stats line")
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    acc = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in rows[1:]:
        stage = next((s for s, names in STAGES.items() if any(n in r[ik] for n in names)), None)
        if stage is None:
            continue
        acc[stage][r[im]] += float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
    units = {"traversal": stats["rays_closest"] + stats["rays_shadow"], "shading": stats["shade_calls"], "sampler": stats["shade_calls"]}
    out = {}
    for stage, m in acc.items():
        n = max(units[stage], 1)
        warp = m.get("smsp__inst_executed.sum", 0.0)
        out[stage] = {"dram_bytes_per_unit": (m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)) / n,
                      "l2_bytes_per_unit": m.get("lts__t_bytes.sum", 0.0) / n,
                      "warp_inst_per_unit": warp / n,
                      "threads_per_inst": m.get("smsp__thread_inst_executed.sum", 0.0) / max(warp, 1.0),
                      "kernel_us_under_ncu": m.get("gpu__time_duration.sum", 0.0), "units": n,
                      "unit": "ray" if stage == "traversal" else "shaded bounce"}
    out["_source"] = {"csv": csv_path, "paths": stats["paths"], "command": "tools/profile_step.py under ncu --metrics (see tools/ncu_counters.py)"}
    return out


if __name__ == "__main__":
    a = sys.argv[1:]
    print(json.dumps({a[i]: one(a[i], a[i + 1], a[i + 2]) for i in range(0, len(a), 3)}, indent=1))
