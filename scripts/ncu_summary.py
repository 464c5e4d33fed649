"""Turns gpurun_out/*.ncu-rep (+ the launch-list csv) into the committed text summaries under profiles/.
usage: python scripts/ncu_summary.py <report.ncu-rep> <out.txt> [launches.csv]"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_read.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_op_global_red.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
        "lts__t_requests.sum", "lts__t_requests_srcunit_tex.sum", "lts__t_requests_srcunit_ltcfabric.sum",
        "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_requests_srcunit_tex_op_red.sum", "lts__cycles_active.sum",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    traffic = {}
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        name = rec.get("Kernel Name", "?")
        lines.append("== %s  (launch id %s)" % (name, rec.get("ID", "?")))
        for i, h in enumerate(hdr):
            if h in KEEP or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                lines.append("  %-78s %18s %s" % (h, vals[i], units[i]))
        try:
            def num(k):
                v, u = float(rec[k].replace(",", "")), units[hdr.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic[name.split("<")[0].split("(")[0].replace("void ", "").strip()] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        except Exception:
            pass
    if len(sys.argv) > 3 and os.path.isfile(sys.argv[3]):
        lines.append("")
        lines.append("== launch list (gpu__time_duration.sum, cold-cache & serialised: compare shares)")
        with open(sys.argv[3]) as f:
            body = [l for l in f if not l.startswith("==")]
        tot = {}
        for r in csv.DictReader(body):
            k = r["Kernel Name"].split("(")[0]
            tot.setdefault(k, []).append(float(r["Metric Value"].replace(",", "")))
        allsum = sum(sum(v) for v in tot.values())
        for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
            lines.append("  %-70s n=%3d  mean %10.1f ns  share %5.1f %%" % (k[:70], len(v), sum(v) / len(v), 100 * sum(v) / allsum))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if traffic:
        tj = os.path.join(os.path.dirname(out), "roofline_traffic.json")
        old = json.load(open(tj)) if os.path.isfile(tj) else {}
        old.update(traffic)
        json.dump(old, open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
