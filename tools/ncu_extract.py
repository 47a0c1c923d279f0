"""Turns `ncu -i REPORT --page raw --csv` into the small JSON summaries committed under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_extract.py /tmp/raw.csv profiles/rNN_name.json ["free-text source line"]
"""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active": "tc_inst_pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "tc_smem_wavefront_pct",
    "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active": "umma_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "sm__cycles_elapsed.max": "sm_cycles",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_bytes.sum": "l2_bytes",
}


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return s


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"id": d["ID"], "kernel": d["Kernel Name"][:160], "grid": d["Grid Size"], "block": d["Block Size"]}
        for k, v in d.items():
            for w, name in WANT.items():
                if k.endswith(w):
                    rec[name] = num(v)
                    rec[name + "_unit"] = units[hdr.index(k)]
        out.append(rec)
    json.dump({"source": sys.argv[3] if len(sys.argv) > 3 else src, "launches": out}, open(dst, "w"), indent=1)
    for r in out:
        print({k: v for k, v in r.items() if not k.endswith("_unit")})


if __name__ == "__main__":
    main()
