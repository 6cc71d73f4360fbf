"""Group the per-launch ncu summary (tools/ncu_summary.py output) by kernel -> profiles/*_ncu_kernel_metrics.json.

usage: python tools/ncu_kernel_metrics.py profiles/r1_ncu_full_summary.csv "<source note>" > profiles/r1_ncu_kernel_metrics.json
The `_tensor_class` entry (all tcgen05 kernels) is what bench.py reports as roofline.traffic (DRAM bytes per launch)."""
import csv
import json
import sys

TENSOR = ("gemm_ws_kernel", "gemm_tc_kernel", "sa_fused", "fp_fused_kernel", "fp1_fused_kernel", "fp1_tail_kernel", "x0net_fused_kernel")


def main(path, source):
    rows = list(csv.DictReader(open(path)))
    out, order = {}, []
    for r in rows:
        k = r["kernel"]
        if k not in out:
            out[k] = {"launches_captured": 0, "dur": 0.0, "dram": 0.0, "tensor_w": 0.0, "issue_w": 0.0, "dram_w": 0.0}
            order.append(k)
        o = out[k]
        d = float(r["dur_us"])
        o["launches_captured"] += 1
        o["dur"] += d
        o["dram"] += float(r["dram_rd_MB"] or 0) + float(r["dram_wr_MB"] or 0)
        o["tensor_w"] += d * float(r.get("tensor_pct") or 0)
        o["issue_w"] += d * float(r.get("issue_pct") or 0)
        o["dram_w"] += d * float(r.get("dram_pct") or 0)
    res = {}
    for k in order:
        o = out[k]
        n, dur = o["launches_captured"], o["dur"]
        res[k] = {"launches_captured": n, "avg_dur_us": dur / n, "avg_dram_MB_per_launch": o["dram"] / n,
                  "tensor_pipe_pct_time_weighted": o["tensor_w"] / dur, "issue_active_pct_time_weighted": o["issue_w"] / dur,
                  "dram_throughput_pct_time_weighted": o["dram_w"] / dur}
    tk = [k for k in order if k.startswith(TENSOR)]
    n = sum(out[k]["launches_captured"] for k in tk)
    dram = sum(out[k]["dram"] for k in tk)
    res["_tensor_class"] = {"kernels": tk, "launches_captured": n, "dram_MB_total": dram,
                            "avg_dram_bytes_per_launch": dram * 1e6 / max(n, 1),
                            "dur_us_total": sum(out[k]["dur"] for k in tk), "source": source}
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
