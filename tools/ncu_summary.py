"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small JSON for profiles/.
Usage: python tools/ncu_summary.py REPORT.ncu-rep OUT.json [--schedule prof_sched.json] [--note "..."]"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
sched, note = None, ""
for i, a in enumerate(sys.argv):
    if a == "--schedule":
        sched = json.load(open(sys.argv[i + 1]))["schedule_n_alive_n_step_n_samples"]
    if a == "--note":
        note = sys.argv[i + 1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = {
    "duration_us": "gpu__time_duration.sum", "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
    "l2_bytes": "lts__t_bytes.sum", "warp_inst": "smsp__inst_executed.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct", "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size", "block": "launch__block_size", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "tensor_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "stall_lg_throttle": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "stall_mio_throttle": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "l2_red_sectors": "lts__t_sectors_srcunit_tex_op_red.sum",
    "l2_red_pct_of_peak_avg_slice": "lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed",
    "l2_throughput_pct_avg_slice": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_throughput_pct_max_slice": "lts__throughput.max.pct_of_peak_sustained_elapsed",
    "l2_throughput_pct_min_slice": "lts__throughput.min.pct_of_peak_sustained_elapsed",
    "l1_red_requests": "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
    "l2_sectors": "lts__t_sectors.sum",
}
scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
launches = []
for k, r in enumerate(data):
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for key, metric in want.items():
        if metric in hdr:
            i = hdr.index(metric)
            try:
                d[key] = float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
            except ValueError:
                pass
    if sched and k < len(sched):
        d["loop_iteration"], d["n_alive"], d["n_step"], d["samples"] = k, *sched[k]
        if d["samples"]:
            d["dram_bytes_per_sample"] = (d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0)) / d["samples"]
            d["l2_bytes_per_sample"] = d.get("l2_bytes", 0) / d["samples"]
            d["thread_inst_per_sample"] = 32 * d.get("warp_inst", 0) / d["samples"]
            d["gsamples_per_s"] = d["samples"] / d["duration_us"] / 1e3
    launches.append(d)
summary = {"report": rep.split("/")[-1], "note": note, "launches": launches}
act = [d for d in launches if d.get("samples")]
if act:
    ns = sum(d["samples"] for d in act)
    summary["per_sample"] = {"dram_bytes": sum(d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0) for d in act) / ns,
                             "l2_bytes": sum(d.get("l2_bytes", 0) for d in act) / ns,
                             "thread_instructions": 32 * sum(d.get("warp_inst", 0) for d in act) / ns,
                             "samples": ns, "launches": len(act)}
json.dump(summary, open(out, "w"), indent=1)
print(json.dumps(summary.get("per_sample", {})))
