"""Condense an `ncu --page raw --csv` export into one line per profiled launch.
usage: python tools/ncu_summary.py file.csv [more metrics...]"""
import csv, re, sys

KEYS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "MB_rd"),
    ("dram__bytes_write.sum", "MB_wr"),
    ("dram__bytes.sum.per_second", "dramGB/s"),          # achieved HBM bandwidth (read + write)
    ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram_act%"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor%"),
    ("lts__t_bytes.sum", "L2_MB"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__cycles_active.avg", "cyc"),
]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6,
        "byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1.0, "Tbyte/s": 1e3}
HBM_PEAK_GBS = 6544.0  # MEASURED_PEAKS.json hbm_gbs on this pool's B200s; the last column is dramGB/s over this

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hi], rows[hi + 1]
extra = sys.argv[2:]
cols = {h: i for i, h in enumerate(hdr)}
if "--list" in extra:
    for h, u in zip(hdr, units):
        print(h, u)
    sys.exit(0)
want = [(k, n) for k, n in KEYS if k in cols] + [(k, k[-24:]) for k in extra if k in cols]
print("%-44s %9s " % ("kernel", "grid") + " ".join("%9s" % n for _, n in want) + " %9s" % "hbm_frac")
for r in rows[hi + 2:]:
    if len(r) != len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[cols["Kernel Name"]]).replace("void ", "").replace("goten::", "")
    vals, frac = [], ""
    for k, n in want:
        v = r[cols[k]].replace(",", "")
        try:
            f = float(v) * UNIT.get(units[cols[k]], 1.0)
            vals.append("%9.1f" % f)
            if n == "dramGB/s":
                frac = "%9.3f" % (f / HBM_PEAK_GBS)
        except ValueError:
            vals.append("%9s" % v[:9])
    print("%-44s %9s " % (name[:44], r[cols["Grid Size"]].replace(" ", "")[:9]) + " ".join(vals) + " " + frac)
