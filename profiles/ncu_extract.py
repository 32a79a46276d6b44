"""`ncu -i X.ncu-rep --page raw --csv` (one row per launch, ~1000 columns) -> compact `metric,unit,launch0,launch1,...` table with the metrics the roofline
discussion uses (duration, DRAM bytes, tensor-pipe / issue / L2 / L1 utilisation, instruction and request counts, registers).
usage: ncu_extract.py raw.csv out.csv [kernel-substring ...]"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    subs = sys.argv[3:]
    kn = hdr.index("Kernel Name")
    if subs:
        data = [r for r in data if any(s in r[kn] for s in subs)]
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        w.writerow(["Kernel Name", ""] + [r[kn] for r in data])
        for k in KEEP:
            if k in hdr:
                c = hdr.index(k)
                w.writerow([k, units[c]] + [r[c] for r in data])


if __name__ == "__main__":
    main()
