"""Summarise an `ncu --set full` report: one line per profiled launch with duration, DRAM traffic, pipe utilisation.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx_full.txt
"""
import csv
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st_sectors"),
    ("lts__t_sectors_op_atom.sum", "l2_atom_sectors"),
    ("lts__t_sectors_op_red.sum", "l2_red_sectors"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    grid_i = hdr.index("Grid Size") if "Grid Size" in hdr else None
    print(f"# {rep}: ncu --set full --clock-control none; per-launch values (cold-cache, serialised replays)")
    present = [(hdr.index(m), short, units[hdr.index(m)]) for m, short in COLS if m in hdr]
    print("# columns: kernel | " + " | ".join(f"{s} [{u}]" for _, s, u in present))
    for r in rows[2:]:
        name = r[name_i].split("(")[0][-48:]
        vals = []
        for i, s, u in present:
            try:
                vals.append(f"{s}={float(r[i]):.4g}")
            except ValueError:
                vals.append(f"{s}={r[i]}")
        g = f" grid={r[grid_i]}" if grid_i is not None else ""
        print(f"{name:<48}{g} " + " ".join(vals))


if __name__ == "__main__":
    main()
