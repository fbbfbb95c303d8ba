#!/usr/bin/env python
"""Key metrics of one kernel launch from an `ncu --set full` report, as JSON (the summaries under profiles/ are made with this).
usage: summarize_ncu.py report.ncu-rep chain_steps_in_the_profiled_launch [note]"""
import csv, json, subprocess, sys
rep, chain_steps = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit, val = rows[0], rows[1], rows[-1]
col = {h: i for i, h in enumerate(hdr)}
def g(name, scale=1.0):
    if name not in col: return None
    v = val[col[name]].replace(",", "")
    try: x = float(v)
    except ValueError: return v
    u = unit[col[name]]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return x * mult * scale
stall_names = ["long_scoreboard", "wait", "not_selected", "selected", "short_scoreboard", "math_pipe_throttle", "no_instructions", "dispatch_stall", "branch_resolving",
               "barrier", "lg_throttle", "mio_throttle", "tex_throttle", "drain", "imc_miss", "membar", "sleeping", "misc"]
stalls = {n: g("smsp__pcsamp_warps_issue_stalled_" + n) for n in stall_names}
tot = sum(v for v in stalls.values() if isinstance(v, float))
dur = g("gpu__time_duration.sum")
rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
inst = g("smsp__inst_executed.sum")
out = {
    "kernel": val[col["Kernel Name"]] if "Kernel Name" in col else None,
    "grid": val[col["Grid Size"]] if "Grid Size" in col else None, "block": val[col["Block Size"]] if "Block Size" in col else None,
    "duration_ms": dur * 1e3, "chain_steps_in_launch": chain_steps, "chain_steps_per_s_under_ncu": chain_steps / dur,
    "registers_per_thread": g("launch__registers_per_thread"), "local_memory_bytes_per_thread_note": "see profiles/r02_sass_*.txt (cuobjdump -res-usage STACK)",
    "warp_instructions": inst, "warp_instructions_per_chain_step": inst / chain_steps,
    "active_lanes_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "issue_slots_busy_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "ipc_per_sm": g("sm__inst_executed.avg.per_cycle_active"),
    "fp64_pipe_active_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "inst_executed_pipe_fp64": g("sm__inst_executed_pipe_fp64.sum"), "inst_executed_pipe_tensor_dmma_pct": g("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
    "achieved_occupancy_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_chain_step": (rd + wr) / chain_steps, "dram_throughput_pct": g("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    "dram_sectors_read": g("dram__sectors_read.sum"), "l2_read_miss_sectors_from_l1": g("lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum"),
    "local_load_sectors": g("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"), "local_store_sectors": g("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
    "stall_shares_pct": {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -(kv[1] or 0)) if isinstance(v, float) and v > 0},
    "note": sys.argv[3] if len(sys.argv) > 3 else "",
}
print(json.dumps(out, indent=1))
