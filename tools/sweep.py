"""Run bench.py over a few kernel configurations and print one line each (scratch helper).
usage: sweep.py WORKLOAD "slots,threads,persistent" ..."""
import json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1]
for cfg in sys.argv[2:]:
    slots, threads, pers = cfg.split(",")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", "10", "--no-cpu-baseline",
                          "--slots", slots, "--threads", threads, "--persistent", pers], capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not lines:
        print(cfg, "FAILED", out.stderr[-400:]); continue
    d = json.loads(lines[-1])
    print(f"slots={slots} threads={threads} persistent={pers}: {d['value']/1e9:.3f} Gelem/s  frac {d['roofline']['frac']:.4f}  "
          f"setups/elem {d['setup']['element_setups_per_element']:.2f} pad {d['setup']['item_padding_efficiency']:.3f}", flush=True)
