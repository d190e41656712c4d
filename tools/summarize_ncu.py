"""Turn the raw ncu outputs of tools/gpu_final.sh into the committed text summaries under profiles/.
usage: python tools/summarize_ncu.py [gpurun_out] [profiles] [rNN]"""
import csv, io, os, re, subprocess, sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles"
tag = sys.argv[3] if len(sys.argv) > 3 else "r01"

# ---- launch list -> per-kernel totals
rows = [r for r in csv.reader(l for l in open(os.path.join(src, "launches.csv")) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows:
    name = re.sub(r"\(.*", "", r[ik]).strip()
    ns = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[iu], 1.0)
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += ns
total = sum(v[1] for v in tot.values())
with open(os.path.join(dst, "%s_ncu_launch_summary_batch2.txt" % tag), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off : one step of tools/profile_step.py 2 "
            "(batch 2, 5 levels, 512x512); cold-cache serialised launches\n")
    f.write("total %.1f us over %d launches\n\n" % (total / 1e3, len(rows)))
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write("%-50s n=%4d %10.1f us  %5.1f%%  avg %8.1f us\n" % (name[:50], n, ns / 1e3, 100 * ns / total, ns / n / 1e3))
with open(os.path.join(dst, "%s_ncu_launch_list_batch2.csv" % tag), "w") as f:
    f.write("".join(l for l in open(os.path.join(src, "launches.csv")) if l.startswith('"')))

# ---- full capture of the dominant kernel -> key metrics
rep = os.path.join(src, "prof_conv.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u = rr[0], rr[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
with open(os.path.join(dst, "%s_ncu_full_conv_tc2.txt" % tag), "w") as f:
    f.write("ncu --set full --clock-control none -k regex:conv_tc2 -s 20 -c 3 python tools/profile_step.py 2   (batch 2; launches 21-23 of one step)\n")
    for r in rr[2:]:
        d = dict(zip(h, r))
        f.write("---- %s grid %s block %s\n" % (d["Kernel Name"][:60], d["Grid Size"], d["Block Size"]))
        for k in KEYS:
            if k in d:
                f.write("   %-75s %s %s\n" % (k, d[k], u[h.index(k)]))
print("wrote summaries to", dst)
