"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv            > profiles/X_launches.txt
    python tools/ncu_summary.py raw gpurun_out/X.ncu-rep                      > profiles/X_metrics.txt
    python tools/ncu_summary.py source gpurun_out/X.ncu-rep [top]             > profiles/X_hotspots.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def short(name):
    return re.sub(r"^void |<unnamed>::|\(.*$", "", name)


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ix["Metric Unit"]], 1.0)
        a = agg.setdefault(short(r[ix["Kernel Name"]]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: per-kernel totals of gpu__time_duration.sum (ncu, cold cache, serialised -- compare SHARES)")
    for k, (n, ms) in agg.items():
        print(f"{k:50s} launches={n:4d} total_ms={ms:9.3f} share={ms / tot * 100:5.1f}%")


def raw(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== {short(r[ix['Kernel Name']])}  grid={r[ix['Grid Size']]} block={r[ix['Block Size']]}")
        for m in METRICS:
            if m in ix:
                print(f"   {m} [{units[ix[m]]}] = {r[ix[m]]}")


def source(rep, top=40):
    rows = ncu_csv(rep, "source")
    # several kernels may follow one another: split on the "Kernel Name" marker rows
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": short(r[1]), "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        if len(b["rows"]) < 2:
            continue
        hdr, data = b["rows"][0], b["rows"][1:]
        key = (b["name"], len(data), tuple(r[0] for r in data[:3]))
        if key in seen:                       # ncu lists each result twice when --page source is exported as csv
            continue
        seen.add(key)
        ix = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        inst = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
        print(f"== {b['name']}: {len(data)} SASS instructions, {inst} warp instructions executed, {tot} stall samples")
        agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
        print("   stall samples: " + ", ".join(f"{k[6:]}={v} ({v / tot * 100:.0f}%)" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
        ops_i, ops_s = collections.Counter(), collections.Counter()
        for r in data:
            p = r[ix["Source"]].split()
            op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
            ops_i[op] += int(r[ix["Instructions Executed"]])
            ops_s[op] += int(r[ix["# Samples"]])
        print("   opcode mix (share of executed warp instructions / of stall samples):")
        for op, v in ops_i.most_common(14):
            print(f"      {op:8s} {v / inst * 100:5.1f}% / {ops_s[op] / tot * 100:5.1f}%")
        print(f"   top {top} instructions by samples (sass index, source, samples %, executions, avg active threads, top stalls):")
        order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
        for i in sorted(order):
            r = data[i]
            st = sorted(((s[6:], int(r[ix[s]])) for s in stalls if int(r[ix[s]]) > 0), key=lambda x: -x[1])[:2]
            print(f"      {i:5d} {r[ix['Source']].strip()[:64]:64s} {int(r[ix['# Samples']]) / tot * 100:5.1f}% x{r[ix['Instructions Executed']]:>9s} thr={r[ix['Avg. Threads Executed']]:>3s} {st}")


def frame_metrics(path, workload="dragon4k", out=None):
    """gpurun_out/X_frame_metrics.csv (one frame's kernels, `ncu --metrics ... --csv`) -> per-kernel table on stdout and, with
    `out`, the entry of profiles/issue.json that bench.py's roofline.issue reads."""
    import json
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[1:]:
        k = (r[ix["ID"]], short(r[ix["Kernel Name"]]))
        per.setdefault(k, {})[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    # the capture may hold several frames: keep the last one (from its k_primary launch on)
    keys = list(per.keys())
    starts = [i for i, k in enumerate(keys) if k[1].startswith("k_primary")]
    if starts:
        per = collections.OrderedDict((k, per[k]) for k in keys[starts[-1]:])
    agg = collections.OrderedDict()
    print(f"# {path}: one frame of {workload}, kernel by kernel (ncu, serialised)")
    for (kid, name), m in per.items():
        inst = m["smsp__inst_executed.sum"][0]
        lanes = m["smsp__thread_inst_executed_per_inst_executed.ratio"][0]
        t, unit = m["gpu__time_duration.sum"]
        ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1.0)
        def val(key, scale=1.0):
            return m[key][0] * scale if key in m else float("nan")
        dram = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
        du = m.get("dram__bytes_read.sum", (0, "byte"))[1]
        dram_mb = dram * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(du, 1e-6)
        print(f"{name:28s} {ms:8.3f} ms  warp-instr {inst / 1e6:9.1f} M  lanes/instr {lanes:5.2f}  issue {val('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} %"
              f"  L1 hit {val('l1tex__t_sector_hit_rate.pct'):5.1f} %  L2 hit {val('lts__t_sector_hit_rate.pct'):5.1f} %  DRAM {dram_mb:7.1f} MB")
        key = name.split("<")[0].replace("k_", "")
        if key.startswith("overflow"):
            key = "overflow_shadow" if "<1" in name or "(TraverseMode)1" in name else "overflow_bounce"
        a = agg.setdefault(key, {"inst": 0.0, "lane_inst": 0.0, "ms": 0.0})
        a["inst"] += inst; a["lane_inst"] += inst * lanes; a["ms"] += ms
    tot_i = sum(a["inst"] for a in agg.values()); tot_l = sum(a["lane_inst"] for a in agg.values())
    print("# per kernel type: " + "  ".join(f"{k}: {a['inst'] / 1e6:.0f} M @ {a['lane_inst'] / a['inst']:.1f} lanes" for k, a in agg.items()))
    print(f"# frame: {tot_i / 1e6:.0f} M warp instructions, {tot_l / tot_i:.2f} lanes per instruction")
    if out:
        try:
            doc = json.load(open(out))
        except Exception:
            doc = {}
        doc[workload] = {"warp_instructions_per_frame": {k: a["inst"] for k, a in agg.items()},
                         "lanes_per_instruction": {k: a["lane_inst"] / a["inst"] for k, a in agg.items()},
                         "source": os.path.basename(path) + " (ncu --metrics smsp__inst_executed.sum, one frame)"}
        json.dump(doc, open(out, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    import os
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "frame":
        frame_metrics(path, *(sys.argv[3:5]))
    elif mode == "launches":
        launches(path)
    elif mode == "raw":
        raw(path)
    else:
        source(path, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
