"""Summarise ncu artefacts from gpurun_out/ into small committed files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r01b_launches.csv profiles/r01b_launch_list_summary.txt "<cmd>"
    python tools/ncu_summary.py report   gpurun_out/r01b_step_full.ncu-rep profiles/r01b_step_full_v2.json
"""
import csv, io, json, subprocess, sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__warps_eligible.avg.per_cycle_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__icc_request_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][-60:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 (%s)\n" % cmd)
        f.write("first 400 launches (burn-in dominated); per-launch times are cold-cache and serialised: compare SHARES\n\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-62s launches %4d  total %10.1f us  mean %7.1f us  share %5.1f%%\n" % (name, n, us, us / n, 100 * us / tot))
    print(open(dst).read())


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = OrderedDict(kernel=r[hdr.index("Kernel Name")].split("(")[0], launch_id=r[hdr.index("ID")])
        m = OrderedDict()
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                m[k] = "%s %s" % (r[i], units[i])
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        d["metrics"] = m
        d["warp_stalls_per_issue"] = OrderedDict((n, round(v, 3)) for v, n in sorted(stalls, reverse=True)[:8])
        try:
            d["dram_bytes_per_launch"] = (float(r[hdr.index("dram__bytes_read.sum")]) * _scale(units[hdr.index("dram__bytes_read.sum")])
                                          + float(r[hdr.index("dram__bytes_write.sum")]) * _scale(units[hdr.index("dram__bytes_write.sum")]))
        except Exception:
            pass
        res.append(d)
    json.dump({"source": src, "how": "ncu --set full --clock-control none --import-source on (one launch per entry)", "launches": res},
              open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1)[:6000])


def _scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        report(sys.argv[2], sys.argv[3])
