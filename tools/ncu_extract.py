"""One line per kernel launch of an .ncu-rep (ncu --set full) with the metrics DESIGN.md argues from:  python tools/ncu_extract.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("smsp__inst_executed.sum", "inst"),
        ("lts__t_sectors_srcunit_tex_op_red.sum", "l2_red_sectors"), ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_read_sectors")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in out.splitlines() if not l.startswith("==")]))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    parts = []
    for key, short in WANT:
        if key in hdr:
            i = hdr.index(key)
            v = r[i].replace(",", "")
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            parts.append(f"{short}={v}{units[i] if units[i] not in ('', '%') else ''}")
    print(name[:70], "|", "  ".join(parts))
