"""Summarise an `ncu --set full` report as the metric table kept under profiles/.
    python profiles/ncu_summary.py report.ncu-rep [kernel-name-substring] > profiles/rN_<kernel>.txt
One column per captured launch (launches whose name does not contain the substring are skipped)."""
import csv
import subprocess
import sys

KEYS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed.avg.per_cycle_elapsed smsp__inst_executed.sum
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum sm__cycles_elapsed.avg""".split()


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    data = [d for d in data if want in d[name_i]]
    keys = KEYS + [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
    print("# source: %s (ncu --set full --clock-control none), one column per captured launch" % rep.split("/")[-1])
    print("kernel".ljust(78), " | ".join(d[name_i].split("(")[0][-40:] for d in data))
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(k.ljust(78), " | ".join(d[i] for d in data), " [%s]" % units[i])


if __name__ == "__main__":
    main()
