"""Summarise an ncu raw CSV (ncu -i X.ncu-rep --page raw --csv) into one compact line per profiled launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [("grid", "Grid Size"), ("us", "gpu__time_duration.sum"),
        ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("tensor_rt_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        ("hmma_ops_pct", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2_hit", "lts__t_sector_hit_rate.pct"),
        ("regs", "launch__registers_per_thread"), ("waves", "launch__waves_per_multiprocessor"),
        ("occ_smem", "launch__occupancy_limit_shared_mem"), ("smem_KB", "launch__shared_mem_per_block_dynamic")]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


for r in rows[2:]:
    out = [r[idx["Kernel Name"]].split("(")[0][-28:]]
    for short, name in WANT:
        if name not in idx:
            continue
        v, u = num(r[idx[name]]), units[idx[name]]
        if v is None:
            out.append(f"{short}={r[idx[name]]}")
            continue
        if short == "us":
            v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        if short.endswith("_MB"):
            v = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0) * v
        if short == "smem_KB":
            v = {"byte": 1 / 1024, "Kbyte": 1.0}.get(u, 1.0) * v
        out.append(f"{short}={v:.1f}" if isinstance(v, float) else f"{short}={v}")
    print(" ".join(out))
