#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> (1) a markdown table of the captured launches, (2) the per-kernel DRAM traffic JSON that
bench.py's roofline.traffic reads.  One step's worth of launches: the first `--launches N` rows (default: all).
usage: ncu -i rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv profiles/rNN_ncu_full_summary.md profiles/rNN_traffic.json "title" [N]"""
import csv
import json
import sys

COLS = [("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"),
        ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        # the counters BASELINE.json's north_star names: shared-memory bank conflicts, sectors per global request (32 B sectors:
        # 4 = fully coalesced 128 B per warp request for 4-byte loads), achieved occupancy
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smem conflicts ld"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smem conflicts st"),
        ("l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio", "sectors/req ld"),
        ("l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio", "sectors/req st"),
        ("sm__maximum_warps_per_active_cycle_pct", "theoretical occ %"),
        ("smsp__cycles_active.avg", "SMSP active cycles")]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
PEAK_GBS = 6543.1      # MEASURED_PEAKS.json hbm_gbs (copy burst on this pool's B200s)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    n = int(sys.argv[5]) if len(sys.argv) > 5 else len(rows) - 2
    body = rows[2:2 + n]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [(m, t) for m, t in COLS if m in ix]
    out = ["# " + sys.argv[4], "", "`ncu --set full --import-source on --clock-control none` under gpurun (one GPU); one row per captured launch, in launch order.",
           "Per-launch times here are cold-cache and serialised (every kernel is replayed ~40 times); bench.py's CUDA-event times are the ones to quote.", "",
           "| kernel | " + " | ".join("%s [%s]" % (t, units[ix[m]]) if units[ix[m]] else t for m, t in cols) + " |", "|---|" + "---:|" * len(cols)]
    traffic = {}
    def num(r, m):
        try:
            return float(r[ix[m]].replace(",", "")) * SCALE.get(units[ix[m]], 1.0)
        except (KeyError, ValueError):
            return None

    # derived columns: achieved DRAM GB/s, sectors per global load / store request (4 = a fully coalesced warp of 4-byte accesses)
    def gbs(r):
        return (num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")) / (num(r, "gpu__time_duration.sum") or 1e30) / 1e9

    derived = [("DRAM GB/s", lambda r: gbs(r) if num(r, "dram__bytes_read.sum") is not None else None),
               ("DRAM % of measured peak", lambda r: 100.0 * gbs(r) / PEAK_GBS if num(r, "dram__bytes_read.sum") is not None else None),
               ("sectors/req ld", lambda r: num(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / max(num(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), 1.0)
                if "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum" in ix else None),
               ("sectors/req st", lambda r: num(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum") / max(num(r, "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"), 1.0)
                if "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum" in ix else None),
               ("global atomics+reds", lambda r: (num(r, "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum") or 0) + (num(r, "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum") or 0)
                if "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum" in ix else None)]
    out[-2] = out[-2][:-2] + " | " + " | ".join(t for t, _ in derived) + " |"
    out[-1] = out[-1] + "---:|" * len(derived)
    for r in body:
        name = r[ix["Kernel Name"]].split("(")[0]
        cells = []
        for m, _ in cols:
            v = r[ix[m]]
            try:
                cells.append("%.3f" % float(v.replace(",", "")) if "." in v else v)
            except ValueError:
                cells.append(v)
        for _, f in derived:
            try:
                v = f(r)
            except (TypeError, ZeroDivisionError):
                v = None
            cells.append("-" if v is None else ("%.2f" % v if v < 1e6 else "%.3g" % v))
        out.append("| %s | %s |" % (name, " | ".join(cells)))
        t = traffic.setdefault(name, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        t["launches"] += 1
        for key, m in (("dram_read_bytes", "dram__bytes_read.sum"), ("dram_write_bytes", "dram__bytes_write.sum")):
            t[key] += float(r[ix[m]].replace(",", "")) * SCALE.get(units[ix[m]], 1.0)
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
    json.dump({"source": sys.argv[4], "note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of the kernel in one step",
               "kernels": traffic}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
