"""Summarise `ncu -i X.ncu-rep --page raw --csv` files: one line per captured launch with the counters the
roofline discussion uses (FP64 / DMMA pipe utilisation, DRAM bytes and throughput, registers, occupancy)."""
import csv, sys, re
KEYS = [("gpu__time_duration.sum", "t"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_inst%"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_inst%"),
        ("TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "fp64_pipe%"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__cycles_active.avg", "cyc")]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        parts = []
        for k, short in KEYS:
            if k in d and d[k] != "":
                u = units[hdr.index(k)]
                parts.append(f"{short}={d[k]}{u if u not in ('%','') else ''}")
        print(f"{name[:40]:40s} " + " ".join(parts))
