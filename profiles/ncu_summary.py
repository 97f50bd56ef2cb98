"""Extract the judged metrics from .ncu-rep files into a JSON summary.
usage: python profiles/ncu_summary.py out.json rep1.ncu-rep [rep2 ...]"""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_cycles_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}

out = {}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        rec = {}
        for k, short in KEYS.items():
            if k in d and d[k] not in ("", "n/a"):
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if u.get(k) in UNIT and short in ("duration", "dram_read", "dram_write", "dyn_smem_per_block"):
                    v *= UNIT[u[k]]
                rec[short] = v
        if "dram_read" in rec:
            rec["dram_traffic_bytes"] = rec["dram_read"] + rec.get("dram_write", 0.0)
        out.setdefault(name, []).append(rec)
json.dump(out, open(sys.argv[1], "w"), indent=1)
for k, v in out.items():
    for rec in v:
        print("%-28s dur %.1f us  dram %.1f MB  dram%% %.1f  issue%% %.1f  warps%% %.1f  inst %.0fM  thr/inst %.1f  regs %d" % (
            k[:28], rec.get("duration", 0) * 1e6, rec.get("dram_traffic_bytes", 0) / 1e6, rec.get("dram_pct_of_peak", 0),
            rec.get("issue_active_pct", 0), rec.get("warps_active_pct", 0), rec.get("warp_instructions", 0) / 1e6,
            rec.get("threads_per_instruction", 0), rec.get("registers_per_thread", 0)))
