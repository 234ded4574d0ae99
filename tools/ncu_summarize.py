"""Dev helper: turn the ncu CSVs brought back from the GPU box into the text summaries kept under profiles/.
   python tools/ncu_summarize.py launches gpurun_out/final_launches.csv "<command line>" [last_n]
   python tools/ncu_summarize.py full gpurun_out/final_full.ncu-rep "<command line>"   (needs ncu on PATH)"""
import collections, csv, io, subprocess, sys

mode, path, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
if mode == "launches":
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    last_n = int(sys.argv[4]) if len(sys.argv) > 4 else 600
    tail = rows[-last_n:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in tail:
        name = r[4].split("(")[0].strip() if "<" not in r[4] else r[4].split("(RbParams")[0].strip()
        agg[name][0] += 1; agg[name][1] += float(r[-1]) / 1000.0
    tot = sum(v[1] for v in agg.values())
    print(cmd)
    print(f"{len(rows)} launches captured; per-launch device time of the LAST {len(tail)} launches (settled regime + host round trips; cold-cache, serialised under ncu: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:56s} launches={v[0]:4d} avg={v[1] / v[0]:9.2f} us share={100 * v[1] / tot:5.1f}%")
else:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out))); h = r[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    units = r[1]
    print(cmd)
    for v in r[2:]:
        print(f"\n== {v[h.index('Kernel Name')]}   grid {v[h.index('Grid Size')]}")
        for n in want:
            if n in h: print(f"   {n:76s} {v[h.index(n)]:>16s} {units[h.index(n)]}")
        try:
            us = float(v[h.index('gpu__time_duration.sum')]); rd = float(v[h.index('dram__bytes_read.sum')]); wr = float(v[h.index('dram__bytes_write.sum')])
            ur = units[h.index('dram__bytes_read.sum')]; uw = units[h.index('dram__bytes_write.sum')]
            f = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
            tb = rd * f[ur] + wr * f[uw]
            print(f"   -> DRAM traffic {tb / 1e6:.1f} MB in {us:.1f} us = {tb / us / 1e3:.0f} GB/s ({100 * tb / us / 1e3 / 6451.5:.1f} % of the measured 6451.5 GB/s copy peak)")
        except Exception as e:
            pass
