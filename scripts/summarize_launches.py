"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list into per-kernel totals (profiles/r1_launches_summary.csv) and the per-launch DRAM traffic
of the level-scan kernel (profiles/traffic_k_scan_quad_regs.json).
usage: summarize_launches.py gpurun_out/launches_r2.csv profiles "<command that was profiled>" [round tag, default r2] """
import collections, csv, json, re, sys

src, outdir, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
tag = sys.argv[4] if len(sys.argv) > 4 else "r2"
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(r[ii], {"k": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for d in per.values():
    name = re.sub(r"\(.*$", "", d["k"]).replace(",", ";").strip()
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    a[2] += d.get("dram__bytes_read.sum", 0.0) / 1e6; a[3] += d.get("dram__bytes_write.sum", 0.0) / 1e6
tot = sum(a[1] for a in agg.values())
with open(outdir + "/%s_launches_summary.csv" % tag, "w") as f:
    f.write("# ncu launch list summary, %s (final kernels): `%s`\n" % (tag, cmd))
    f.write("# (cold-cache, serialised launches: compare SHARES, not absolutes).  full list: profiles/%s_launches.csv\n" % tag)
    f.write("kernel,launches,total_us,avg_us,share_of_all,avg_dram_read_MB,avg_dram_write_MB\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.1f,%.2f,%.3f,%.2f,%.2f\n" % (k, a[0], a[1], a[1] / a[0], a[1] / tot, a[2] / a[0], a[3] / a[0]))
    step = {k: a for k, a in agg.items() if "k_trigger" in k or "k_scan_quad" in k}
    st = sum(a[1] / a[0] for a in step.values())
    f.write("# per-step kernels (2 launches per step): " + "; ".join(
        "%s %.2f us (%.0f%% of the step kernels)" % (re.sub(r"<.*", "", k.replace("void sbdk::", "")), a[1] / a[0], 100 * a[1] / a[0] / st)
        for k, a in step.items()) + "\n")
for k, a in agg.items():
    if "k_scan_quad_regs" in k:
        json.dump({"kernel": "sbdk::k_scan_quad_regs<5,128>", "dram_bytes_per_launch": 1e6 * (a[2] + a[3]) / a[0],
                   "dram_read_bytes": 1e6 * a[2] / a[0], "dram_write_bytes": 1e6 * a[3] / a[0],
                   "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, mean over the launches in profiles/%s_launches.csv" % tag,
                   "round": tag}, open(outdir + "/traffic_k_scan_quad_regs.json", "w"), indent=1)
