"""Summarise an .ncu-rep (one kernel launch) into a small JSON + text for profiles/."""
import csv, io, json, subprocess, sys, collections

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}

def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in data:
        for h in names:
            agg[h] += int(r[ix[h]])
    tot = sum(agg.values()) or 1
    return {h: round(v / tot * 100, 2) for h, v in agg.most_common() if v}

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

if __name__ == "__main__":
    rep, out = sys.argv[1], sys.argv[2]
    r = raw(rep)
    summary = {"report": rep, "kernel": r.get("Kernel Name", ("", ""))[0], "metrics": {}, "stall_pct_of_samples": stalls(rep)}
    for k in KEYS:
        if k in r:
            summary["metrics"][k] = {"value": r[k][0], "unit": r[k][1]}
    def num(k):
        v, u = r[k]
        v = float(v.replace(",", ""))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return v * mult
    if "dram__bytes_read.sum" in r:
        summary["dram_bytes_per_launch"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps(summary, indent=1)[:3000])
