"""Extracts the judged metrics of an .ncu-rep (ncu --set full) into a small csv: python tools/ncu_extract.py rep.ncu-rep out.csv "<header comment>" """
import csv, io, subprocess, sys
rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ("dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
ki = hdr.index("Kernel Name")
with open(out, "w") as f:
    f.write(f"# {note}\n")
    f.write("metric,unit," + ",".join(f"launch_{i}:{r[ki].split('(')[0][-40:]}" for i, r in enumerate(data)) + "\n")
    for j, name in enumerate(hdr):
        if name in keep or name.startswith("smsp__average_warps_issue_stalled") or "tc_wavefronts" in name or "mem_shared" in name and "pct" in name:
            f.write(f"{name},{units[j]}," + ",".join(r[j] for r in data) + "\n")
print(open(out).read()[:1500])
