"""Turns the ncu raw pages a measurement pass left in gpurun_out/ into the small tracked summaries
under profiles/ (run here, no GPU needed):   python tools/summarize_profiles.py r02"""
import csv
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
SRC, DST = os.path.join(HERE, "gpurun_out"), os.path.join(HERE, "profiles")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def raw_page(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                try:
                    d[h + (" [" + units[i] + "]" if units[i] else "")] = float(r[i].replace(",", ""))
                except ValueError:
                    d[h] = r[i]
        out.append(d)
    return out


def main():
    captures = []
    meta = {f"{R}_scan_f32_12m5_raw.csv": (12_500_000, "f32"), f"{R}_scan_f32_1m_raw.csv": (1_000_000, "f32"),
            f"{R}_scan_bf16_1m_raw.csv": (1_000_000, "bf16")}
    summary = {}
    for name in sorted(os.listdir(SRC)):
        if not (name.startswith(R) and name.endswith("_raw.csv")):
            continue
        page = raw_page(os.path.join(SRC, name))
        summary[name] = page
        shutil.copy(os.path.join(SRC, name), os.path.join(DST, name))
        if name in meta and page:
            rows, dtype = meta[name]
            last = page[-1]

            def get(prefix, scale):
                for k, v in last.items():
                    if k.startswith(prefix + " ["):
                        unit = k[k.index("[") + 1:-1]
                        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
                                    "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[unit] / scale
                return None
            captures.append({"rows": rows, "dtype": dtype, "dram_bytes_read": get("dram__bytes_read.sum", 1),
                             "dram_bytes_write": get("dram__bytes_write.sum", 1),
                             "duration_s": get("gpu__time_duration.sum", 1),
                             "source": f"ncu --set full --clock-control none, profiles/{name}, last captured launch"})
    json.dump({"captures": captures}, open(os.path.join(DST, f"{R}_scan_summary.json"), "w"), indent=1)
    json.dump(summary, open(os.path.join(DST, f"{R}_ncu_summary.json"), "w"), indent=1)
    for name in os.listdir(SRC):
        if name.startswith(R) and (name.endswith("launches.csv") or name.endswith(".json") or name.startswith(R + "_clocks")):
            shutil.copy(os.path.join(SRC, name), os.path.join(DST, name))
    print(json.dumps({"captures": captures}, indent=1))


if __name__ == "__main__":
    main()
