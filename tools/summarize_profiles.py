"""Turn gpurun_out/launches.csv (ncu gpu__time_duration per launch of tools/prof_step.py) and
gpurun_out/prof_kernels.ncu-rep (ncu --set full of tools/prof_kernels.py) into the committed summaries
under profiles/.  Usage: python tools/summarize_profiles.py r1"""
import collections
import csv
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

lc = os.path.join(root, "gpurun_out", "launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    half = len(data) // 2                       # prof_step.py runs two identical iterations; keep the warm one
    agg, tot = collections.OrderedDict(), 0.0
    for r in data[half:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    with open(os.path.join(out_dir, "%s_launch_list_summary.txt" % tag), "w") as f:
        f.write("# one IWGAN iteration (5 critic + 1 generator runs, B=512, L=200, 32x32x3), eager, under\n"
                "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n"
                "# command: ncu ... python tools/prof_step.py   (second of two iterations)\n")
        f.write("%-52s %6s %12s %7s\n" % ("kernel", "n", "total_us", "share"))
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-52s %6d %12.1f %6.1f%%\n" % (k[:52], n, v / 1e3, 100 * v / tot))
        f.write("%-52s %6d %12.1f\n" % ("TOTAL", sum(n for n, _ in agg.values()), tot / 1e3))
    print("wrote launch summary")

rep = os.path.join(root, "gpurun_out", "prof_kernels.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(out_dir, "%s_ncu_kernels.csv" % tag), "w") as f:
        w = csv.writer(f)
        w.writerow(["# ncu --set full --clock-control none, tools/prof_kernels.py: c2 (16x16x208->8x8x400, 200 logical channels), c3 (8x8x400->4x4x800), "
                    "c1 (32x32x3->16x16x208) fprop/dgrad/wgrad at B=512; pix2pix e2/d7 (128x128x64->64x64x128 k4 s2, B=16: persistent kernel)"])
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            if "at::" in r[hdr.index("Kernel Name")]:
                continue                        # torch fill / copy kernels of the test set-up, not ours
            w.writerow([r[i][:60] for i in idx])
    print("wrote ncu kernel summary")
