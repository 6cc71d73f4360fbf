"""Condense an `ncu --page raw --csv` dump into one row per launch with the metrics the roofline uses."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "dur_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("launch__occupancy_limit_registers", "occ_lim_regs"), ("lts__t_bytes.sum", "l2_MB"),
        # global loads that missed L1 (32-byte sectors): what the gather-fused kernels pull from L2
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum", "l2_gather_MB")]


def conv(val, unit, want_mb):
    v = float(val.replace(",", ""))
    if want_mb:
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        return v * scale
    if unit in ("ns", "nsecond"):
        return v / 1e3
    if unit in ("ms", "msecond"):
        return v * 1e3
    return v


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in WANT if m in idx]
    print("kernel,grid," + ",".join(n for _, n in cols))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("void ", "").replace("lsdm::<unnamed>::", "").replace("lsdm::", "")
        name = name.split("(")[0].replace("unnamed>::", "").replace(",", ";")
        out = [name, r[idx["Grid Size"]].replace(",", " ")]
        for m, n in cols:
            try:
                if n == "l2_gather_MB":
                    out.append(f"{float(r[idx[m]].replace(',', '')) * 32e-6:.3f}")
                    continue
                out.append(f"{conv(r[idx[m]], units[idx[m]], n.endswith('_MB')):.3f}")
            except Exception:
                out.append("")
        print(",".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
