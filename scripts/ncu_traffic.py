"""`ncu --page raw --csv` dump -> per-kernel JSON of the figures bench.py quotes: duration, DRAM bytes (roofline.traffic),
L2 reduction sectors (roofline.l2_red), instruction count and issue utilisation, one entry per launch."""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def num(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return None
units = rows[1]
def scaled(r, name):
    v = num(r, name)
    if v is None:
        return None
    u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
out = {}
for r in rows[2:]:
    name = re.sub(r"\(int\)", "", r[col["Kernel Name"]])
    m = re.search(r"(k_\w+(<\d+>)?)", name)
    key = m.group(1) if m else name[:40]
    dur = num(r, "gpu__time_duration.sum")
    du = units[col["gpu__time_duration.sum"]].lower()
    ms = dur * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1, "ms": 1, "second": 1e3, "nsecond": 1e-6}.get(du, 1)
    out.setdefault(key, []).append({
        "dram_bytes": (scaled(r, "dram__bytes_read.sum") or 0) + (scaled(r, "dram__bytes_write.sum") or 0),
        "l2_red_sectors": num(r, "lts__t_sectors_srcunit_tex_op_red.sum"),
        "ms": ms, "issue": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), "inst": num(r, "smsp__inst_executed.sum")})
json.dump(out, sys.stdout, indent=1)
