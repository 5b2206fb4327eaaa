"""Summarise an ncu --set full report for profiles/: selected metrics per launch (transposed CSV) and
the DRAM traffic per launch of the SpMV kernels (profiles/spmv_ncu_traffic.json, read by bench.py).

  python scripts/ncu_summary.py gpurun_out/spmv_r02.ncu-rep profiles/r02_spmv_v3_ncu_full_summary.csv [--traffic]
"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    units = rows[1]
    data = rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for m in METRICS:
            if m in col:
                w.writerow([m, units[col[m]]] + [r[col[m]] for r in data])
    print("wrote", out, "launches:", len(data))
    if "--traffic" in sys.argv:
        def num(r, m):
            return float(r[col[m]].replace(",", ""))

        def to_bytes(v, unit):
            u = unit.lower()
            return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)

        tr = {}
        # launches are classified by the kernel's template argument: spmv_flag_kernel<4> = K1 (POST_MUL, tmp = R_y^-1 A p; <1> = POST_DIV before round 2 call E),
        # <2> = K2 (POST_FMA_DOT + alpha hook, Gp = R_x p + A' tmp) -- the in-loop kernels of scripts/prof_cg.py;
        # <0> launches alternate A x (even) / A'x (odd) in scripts/prof_spmv.py
        name = col["Kernel Name"]
        groups = {"K1": [r for r in data if "spmv_flag_kernel<4>" in r[name] or "spmv_flag_kernel<1>" in r[name]],
                  "K2": [r for r in data if "spmv_flag_kernel<2>" in r[name]]}
        plain = [r for r in data if "spmv_flag_kernel<0>" in r[name]]
        groups["A x"], groups["A'x"] = plain[0::2], plain[1::2]
        for key, sel in groups.items():
            if not sel:
                continue
            tot = [to_bytes(num(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]]) +
                   to_bytes(num(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]]) for r in sel]
            tr[key] = {"dram_bytes_per_launch": sum(tot) / len(tot), "launches": len(tot),
                       "kernel": sel[0][name][:60], "source": os.path.basename(out)}
        p = os.path.join(os.path.dirname(out), "spmv_ncu_traffic.json")
        try:
            prev = json.load(open(p))
        except Exception:
            prev = {}
        prev.update(tr)
        tr = prev
        json.dump(tr, open(p, "w"), indent=1)
        print("wrote", p, tr)


if __name__ == "__main__":
    main()
