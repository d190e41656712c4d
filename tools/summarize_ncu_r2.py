"""Turn the raw ncu outputs of tools/gpu_r2_ncu.sh (gpurun_out/r2_*) into the committed summaries under profiles/.
usage: python tools/summarize_ncu_r2.py"""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    return rr[0], rr[1], rr[2:]


def stalls(h, r):
    st = []
    for i, name in enumerate(h):
        if "pcsamp_warps_issue_stalled" in name and "not_issued" not in name:
            try:
                st.append((float(r[i].replace(",", "")), name.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    return ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in sorted(st, reverse=True)[:7])


def full(rep, out, header, note=None):
    h, u, rows = raw_rows(rep)
    with open(out, "w") as f:
        f.write(header + "\n")
        for r in rows:
            d = dict(zip(h, r))
            f.write("---- %s   grid %s block %s\n" % (d["Kernel Name"][:70], d.get("Grid Size", "?"), d.get("Block Size", "?")))
            for k in KEYS:
                if k in d:
                    f.write("   %-72s %s %s\n" % (k, d[k], u[h.index(k)]))
            f.write("   warp stall sampling (all samples): %s\n" % stalls(h, r))
        if note:
            f.write("\n" + note + "\n")
    return h, u, rows


# ---- launch list of one step (batch 2)
lp = os.path.join(src, "r2_launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r[ik]).strip()
        ns = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[iu], 1.0)
        t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += ns
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(dst, "r02_ncu_launch_summary_batch2.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python tools/profile_step.py 2\n"
                "(one step, batch 2, 5 levels, 512x512; cold-cache serialised launches: compare SHARES, not absolutes)\n")
        f.write("total %.1f us over %d launches\n\n" % (total / 1e3, len(rows)))
        for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-50s n=%4d %10.1f us  %5.1f%%  avg %8.1f us\n" % (name[:50], n, ns / 1e3, 100 * ns / total, ns / n / 1e3))
    with open(os.path.join(dst, "r02_ncu_launch_list_batch2.csv"), "w") as f:
        f.write("".join(l for l in open(lp) if l.startswith('"')))

# ---- conv kernels at the bench batch (30 frames)
cp = os.path.join(src, "r2_conv_b30.ncu-rep")
if os.path.exists(cp):
    h, u, rows = full(cp, os.path.join(dst, "r02_ncu_full_conv.txt"),
                      "ncu --set full --clock-control none --import-source on -k regex:conv_tc2 python tools/ncu_conv.py 30   (one launch each, 30 frames:\n"
                      "3x3 64->64 @512, UP2 64->64 @512 out, 3x3 512->512 @64, 3x3 256->256 @128)")
    d = dict(zip(h, rows[-1]))          # 256->256 @128: the layer shape with the largest share of a step
    def num(k):
        v = float(d[k].replace(",", ""))
        unit = u[h.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    B, hw, cin, cout = 30, 128, 256, 256
    act_in = 2 * 2 * B * (hw + 2) * (hw + 2) * cin          # two fp16 planes
    act_out = 2 * 2 * B * (hw + 2) * (hw + 2) * cout
    wts = 2 * 2 * 9 * cin * cout
    json.dump({"dram_bytes_per_launch": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
               "dram_read": int(num("dram__bytes_read.sum")), "dram_write": int(num("dram__bytes_write.sum")),
               "algorithmic_operand_bytes": {"activations_in": act_in, "activations_out": act_out, "weights": wts,
                                             "total": act_in + act_out + wts},
               "of": "conv_tc2_kernel<128,fused> 3x3 256->256 @128x128, 30 frames (the layer shape with the largest share of the "
                     "step), one launch under `ncu --set full` (profiles/r02_ncu_full_conv.txt)"},
              open(os.path.join(dst, "r02_conv_traffic.json"), "w"), indent=1)

jp = os.path.join(src, "r2_jacobi.ncu-rep")
if os.path.exists(jp):
    full(jp, os.path.join(dst, "r02_ncu_full_jacobi.txt"),
         "ncu --set full --clock-control none --import-source on -k regex:k_jacobi python tools/jacobi_once.py   (k_jacobi<512>, 4 matrices = 4 clusters of 8 CTAs)")
vp = os.path.join(src, "r2_cov.ncu-rep")
if os.path.exists(vp):
    full(vp, os.path.join(dst, "r02_ncu_full_cov.txt"),
         "ncu --set full --clock-control none --import-source on -k regex:cov_tc_kernel python tools/cov_once.py 16   (16 frames: C64@512, C128@256, C256@128, C512@64)")
tp = os.path.join(src, "r2_tail.ncu-rep")
if os.path.exists(tp):
    full(tp, os.path.join(dst, "r02_ncu_full_tail_head.txt"),
         "ncu --set full --clock-control none -k regex:conv_(tail|head)_tc_kernel -s 2 -c 2 python tools/tail_head_once.py 16   (16 frames 512x512: decoder tail 64->3, encoder head 3->64)",
         note="tail: compulsory traffic 16*514*514*256 B read + 16*512*512*12 B written = 1.13 GB; head: 16*512*512*12 B read + 16*514*514*256 B written = 1.13 GB.")
np_ = os.path.join(src, "r2_nsgemm.ncu-rep")
if os.path.exists(np_):
    full(np_, os.path.join(dst, "r02_ncu_full_ns_gemm.txt"),
         "ncu --set full --clock-control none -k regex:ns_gemm -s 55 -c 1 python tools/matfun_once.py   (one batched 512x512x512 product of the Newton-Schulz iteration, 15 matrices = 240 tiles of 128x128)")
print("wrote summaries to", dst)
