"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with the counters the roofline needs."""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("launch__occupancy_limit_registers", "occ_lim_regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("smsp__inst_executed.sum", "inst"), ("sm__cycles_elapsed.max", "cycles"),
        ("lts__t_bytes.sum", "l2_bytes"), ("l1tex__t_bytes.sum", "l1_bytes"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("sm__inst_executed_pipe_xu.sum", "xu_inst")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0][-60:]
        out = [name, "grid=" + r[idx["Grid Size"]] if "Grid Size" in idx else ""]
        for k, short in KEYS:
            if k in idx:
                out.append(f"{short}={r[idx[k]]}{units[idx[k]] if short in ('time','dram_rd','dram_wr','l2_bytes','l1_bytes') else ''}")
        print("  ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
