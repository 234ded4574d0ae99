"""Dev helper (GPU box): A/B the variants built by tools/build_variant.py on the C3 bench.
   python tools/ab.py pf0 pf1 ...   -> one line per variant"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = os.environ.get("AB_STEPS", "400")
for spec in sys.argv[1:]:
    name, *kv = spec.split(":")              # name[:ENV=value[:ENV=value...]]
    env = dict(os.environ)
    for item in kv:
        k, v = item.split("=", 1); env[k] = v
    if name != "default":
        env["REBOOT_B200_LIB"] = os.path.join(ROOT, "variants", f"lib_{name}.so")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "10", "--settle", "600", "--no-cpu", "--no-c5"],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(spec, "ms/step %.4f" % j["ms_per_step"], "free-fall %.4f" % j["free_fall"]["ms_per_step"], "e2e %.1f" % j["e2e"]["value"],
              {k: round(v, 4) for k, v in j["roofline"]["kernel_ms_all"].items()},
              {k: j["audit"][k] for k in ("list_ticks", "list_rebuilds", "timed_list_rebuilds", "list_hot_last")}, "cks", j["state_checksum"], "clk", j["clocks"]["sm_mhz"], "launches", j["gpu_launches"], flush=True)
    except Exception as e:
        print(spec, "FAILED", e, out.stderr[-800:], flush=True)
