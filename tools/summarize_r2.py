"""Round-2 profile summaries: raw ncu / bench output under gpurun_out/ -> tracked text / json under profiles/.
  python tools/summarize_r2.py
Inputs (produced by tools/gpu_session_r2.sh on the GPU box):
  r2_launch_list_tc32_raw.csv  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
                               --clock-control none --csv  python tools/one_step.py tc32        (ONE 1024x2048 step, tc32)
  r2_conv_tc32_full.ncu-rep    ncu --set full --clock-control none --import-source on -k regex:conv_igemm_tc32 -s 3 -c 1
                               python tools/prof_conv.py --tc32                                 (3x3 256->256 @256x512)
  r2_per_call_tc32.jsonl       bench.py --profile-out (CUDA-event time of every C-ABI call of one step, NOT under ncu)
"""
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
HBM_PEAK = 6572.5


def short(name):
    name = name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    return name


def launch_list():
    lines = [l for l in open(os.path.join(G, "r2_launch_list_tc32_raw.csv")) if not l.startswith("==")]
    rd = list(csv.DictReader(lines))
    per = collections.OrderedDict()          # launch id -> dict
    for r in rd:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if u.startswith("ns") else (v if u.startswith("us") else v * 1e3)
        else:
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            d["rd" if "read" in r["Metric Name"] else "wr"] = v * mul
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["kernel"], [0.0, 0, 0.0, 0.0])
        a[0] += d["us"]; a[1] += 1; a[2] += d.get("rd", 0.0); a[3] += d.get("wr", 0.0)
    tot = sum(a[0] for a in agg.values())
    with open(os.path.join(P, "r2_launch_list_tc32_summary.txt"), "w") as f:
        f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n")
        f.write("#     python tools/one_step.py tc32      -- ONE FuseTrack step, 1024x2048 pair, parity precision tc32\n")
        f.write("# %d launches, %.3f ms summed kernel time (cold-cache, serialised under ncu: compare SHARES with bench.py, not absolutes)\n" % (len(per), tot / 1e3))
        f.write("# DRAM GB/s = (dram read + write) / kernel time; frac = GB/s / %.1f (MEASURED_PEAKS.json hbm_gbs)\n" % HBM_PEAK)
        f.write("%-58s %5s %9s %6s %9s %9s %8s %6s\n" % ("kernel", "calls", "ms", "share", "rd MB", "wr MB", "GB/s", "frac"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            gbs = (a[2] + a[3]) / (a[0] * 1e-6) / 1e9 if a[0] > 0 else 0
            f.write("%-58s %5d %9.3f %5.1f%% %9.1f %9.1f %8.0f %6.3f\n" % (k[:58], a[1], a[0] / 1e3, 100 * a[0] / tot, a[2] / 1e6, a[3] / 1e6, gbs, gbs / HBM_PEAK))
    with open(os.path.join(P, "r2_launch_list_tc32.csv"), "w") as f:
        f.write("id,kernel,grid,block,duration_us,dram_read_bytes,dram_write_bytes\n")
        for i, d in per.items():
            f.write("%s,%s,%s,%s,%.3f,%d,%d\n" % (i, d["kernel"].replace(",", ";"), d["grid"].replace(",", "x").replace(" ", ""),
                                                  d["block"].replace(",", "x").replace(" ", ""), d["us"], d.get("rd", 0), d.get("wr", 0)))
    conv = [a for k, a in agg.items() if "igemm_tc32" in k]
    rd_b, wr_b, us = sum(a[2] for a in conv), sum(a[3] for a in conv), sum(a[0] for a in conv)
    return {"kernels": "conv_igemm_tc32_kernel + dcn_igemm_tc32_kernel", "launches": int(sum(a[1] for a in conv)),
            "dram_read_bytes_per_step": rd_b, "dram_write_bytes_per_step": wr_b, "dram_bytes_per_step": rd_b + wr_b,
            "kernel_ms_under_ncu": us / 1e3, "share_of_step_under_ncu": us / tot,
            "source": "profiles/r2_launch_list_tc32.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, one step)"}


def algorithmic_bytes():
    """fp32 in + fp32 out of every tc32 conv / dcn call of one step (weights excluded: < 1 % and L2 resident)"""
    tot = 0.0
    for l in open(os.path.join(G, "r2_per_call_tc32.jsonl")):
        r = json.loads(l)
        if not r["fn"].startswith(("vps_conv2d_tc32", "vps_deform_conv_tc32")):
            continue
        tag = r["tag"]           # "3x3 s1 256->256 @256x512" / "dcn3x3 256->256 @256x512" / "2x2 x4 phases 162->16 @512x1024"
        try:
            ch = tag.split("->")
            cin = int(ch[0].split()[-1]); cout = int(ch[1].split()[0])
            oh, ow = (int(v) for v in tag.split("@")[1].split("x"))
            s = 2 if " s2 " in tag else 1
            nimg = 2 if r["scope"] == "r50fpn" else 1
            mult = 4 if "phases" in tag else 1               # transposed conv: 4 phases write 4x the pixels
            tot += 4.0 * nimg * (cin * oh * s * ow * s + cout * oh * ow * mult)
        except Exception:
            pass
    return tot


def full_capture():
    rep = os.path.join(G, "r2_conv_tc32_full.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
            "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed",
            "smsp__mem_tensor_writes_op_utcmma.sum.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
    with open(os.path.join(P, "r2_conv_tc32_full.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:conv_igemm_tc32 -s 3 -c 1  python tools/prof_conv.py --tc32\n")
        f.write("# kernel: conv_igemm_tc32_kernel, 3x3 256->256 @256x512 (FPN P2 / TCEA shape, 154.6 GFLOP algorithmic, fp32 in / fp32 out,\n")
        f.write("#         3 f16 tensor-core products per MAC); algorithmic DRAM traffic 134 MB in + 134 MB out\n")
        for r in rows[2:]:
            for w in want:
                if w in idx:
                    f.write("%-88s %s %s\n" % (w, r[idx[w]], units[idx[w]]))


def per_call():
    rows = [json.loads(l) for l in open(os.path.join(G, "r2_per_call_tc32.jsonl"))]
    agg = collections.OrderedDict()
    for r in rows:
        key = r["fn"] + (" " + r["tag"] if r["tag"] else "")
        a = agg.setdefault(key, [0.0, 0.0, 0])
        a[0] += r["ms"]; a[1] += r["flops"]; a[2] += 1
    with open(os.path.join(P, "r2_per_call_device_times_tc32.txt"), "w") as f:
        f.write("# bench.py --profile-out: CUDA-event time of every C-ABI call of one step, tc32 precision (not under ncu)\n")
        f.write("%-60s %5s %9s %9s %8s\n" % ("call", "n", "ms", "GFLOP", "TFLOP/s"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write("%-60s %5d %9.3f %9.1f %8.1f\n" % (k[:60], a[2], a[0], a[1] / 1e9, a[1] / a[0] / 1e9 if a[0] > 0 else 0))


def main():
    os.makedirs(P, exist_ok=True)
    t = launch_list()
    t["algorithmic_bytes_per_step"] = algorithmic_bytes()
    t["traffic_over_algorithmic"] = t["dram_bytes_per_step"] / t["algorithmic_bytes_per_step"] if t["algorithmic_bytes_per_step"] else None
    old = {}
    tp = os.path.join(P, "r2_dram_traffic.json")
    if os.path.exists(tp):
        old = json.load(open(tp))
    old["tc32"] = t
    json.dump(old, open(tp, "w"), indent=1)
    full_capture()
    per_call()
    for name in ("r2_bench_line.json", "r2_bench_viper.json", "r2_microbench_flow.json", "r2_sanitizer_memcheck.log"):
        src = os.path.join(G, name)
        if os.path.exists(src):
            open(os.path.join(P, name), "w").write(open(src).read())
    print(json.dumps(t, indent=1))


if __name__ == "__main__":
    main()
