"""Turn raw ncu outputs under gpurun_out/ into the tracked summaries under profiles/.
  python tools/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep> [calls.jsonl]"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
calls = sys.argv[4] if len(sys.argv) > 4 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list (gpu__time_duration per launch of ONE full-size step, cold-cache & serialised under ncu)
lines = [l for l in open(launches) if not l.startswith("==")]
rd = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rd:
    name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e3 if u.startswith("us") else (v / 1e6 if u.startswith("ns") else v)
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
tot = sum(a[0] for a in agg.values())
with open(os.path.join(out_dir, "%s_launch_list_summary.txt" % tag), "w") as f:
    f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none  python tools/one_step.py\n")
    f.write("# one FuseTrack step (1024x2048 pair, bf16 mode); %d launches, %.3f ms summed kernel time\n" % (len(rd), tot))
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's breakdown, not absolutes\n")
    f.write("%-66s %6s %10s %7s\n" % ("kernel", "calls", "ms", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write("%-66s %6d %10.3f %6.1f%%\n" % (k[:66], a[1], a[0], 100 * a[0] / tot))
with open(os.path.join(out_dir, "%s_launch_list.csv" % tag), "w") as f:
    f.write("id,kernel,grid,block,duration_us\n")
    for r in rd:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v if u.startswith("us") else (v / 1e3 if u.startswith("ns") else v * 1e3)
        f.write("%s,%s,%s,%s,%.3f\n" % (r["ID"], r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace(",", ";"),
                                      r["Grid Size"].replace(",", "x").replace(" ", ""), r["Block Size"].replace(",", "x").replace(" ", ""), v))

# ---- full capture of the dominant kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
idx = {h: i for i, h in enumerate(hdr)}
with open(os.path.join(out_dir, "%s_conv_tc_full.txt" % tag), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:conv_igemm  python tools/prof_conv.py\n")
    f.write("# kernel: conv_igemm_tc_kernel on the 3x3 256->256 conv at 256x512 (FPN P2 / TCEA shape, 154.6 GFLOP, bf16)\n")
    for r in rows[2:]:
        for w in want:
            if w in idx:
                f.write("%-66s %s %s\n" % (w, r[idx[w]], units[idx[w]]))
        f.write("\n")
if calls:
    rows = [json.loads(l) for l in open(calls)]
    agg = collections.OrderedDict()
    for r in rows:
        key = r["fn"] + (" " + r["tag"] if r["tag"] else "")
        a = agg.setdefault(key, [0.0, 0.0, 0])
        a[0] += r["ms"]; a[1] += r["flops"]; a[2] += 1
    with open(os.path.join(out_dir, "%s_per_call_device_times.txt" % tag), "w") as f:
        f.write("# bench.py --profile-out: CUDA-event time of every C-ABI call of one step (not under ncu)\n")
        f.write("%-60s %5s %9s %9s %8s\n" % ("call", "n", "ms", "GFLOP", "TFLOP/s"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write("%-60s %5d %9.3f %9.1f %8.1f\n" % (k[:60], a[2], a[0], a[1] / 1e9, a[1] / a[0] / 1e9 if a[0] > 0 else 0))
print("profiles written to", out_dir)
