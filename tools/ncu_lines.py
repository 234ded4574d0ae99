"""Dev helper: per-source-line totals (stall samples, warp instructions, lanes) of one kernel from an .ncu-rep.
   python tools/ncu_lines.py rep.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name": continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        def f(k):
            try: return float(d.get(k, "0"))
            except ValueError: return 0.0
        lines.append((fpath, int(r[0]), r[1].strip()[:110], f("# Samples"), f("Instructions Executed"), f("Thread Instructions Executed"), d))
ts = sum(l[3] for l in lines); ti = sum(l[4] for l in lines); tt = sum(l[5] for l in lines)
print(f"total samples {ts:.0f}  warp-inst {ti:.0f}  lanes/inst {tt / max(ti, 1):.2f}")
print("--- by samples")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    st = {k: v for k, v in l[6].items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("0", "-", "")}
    big = sorted(st.items(), key=lambda kv: -float(kv[1]))[:3]
    print(f"{100 * l[3] / ts:5.1f}% smp {100 * l[4] / ti:5.1f}% inst lanes {l[5] / max(l[4], 1):5.1f} {l[0]}:{l[1]:5d} {l[2]}  {big}")
print("--- by instructions")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{100 * l[4] / ti:5.1f}% inst {100 * l[3] / ts:5.1f}% smp lanes {l[5] / max(l[4], 1):5.1f} {l[0]}:{l[1]:5d} {l[2]}")
if len(sys.argv) > 4:
    # buckets: "file:lo-hi=name,..." 
    print("--- buckets")
    for spec in sys.argv[4].split(","):
        rng, name = spec.split("=")
        fn, lh = rng.split(":"); lo, hi = map(int, lh.split("-"))
        sel = [l for l in lines if l[0] == fn and lo <= l[1] <= hi]
        i = sum(l[4] for l in sel); s = sum(l[3] for l in sel); t = sum(l[5] for l in sel)
        print(f"{name:28s} inst {100 * i / ti:5.1f}%  samples {100 * s / ts:5.1f}%  lanes {t / max(i, 1):5.1f}")
