"""Shared-memory wavefronts per SASS instruction from `ncu --page source --csv --print-source sass`:
lists the instructions that account for the wavefronts (total, ideal, excessive), per element."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
nelem = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ix = {k: hdr.index(k) for k in ("Source", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                                 "L1 Wavefronts Shared Excessive", "# Samples", "stall_mio", "stall_short_sb", "stall_long_sb",
                                 "stall_barrier", "stall_math", "stall_wait", "stall_lg", "stall_not_selected")}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
data = [r for r in rows[2:] if len(r) > ix["# Samples"]]
tw = sum(f(r, "L1 Wavefronts Shared") for r in data)
ts = sum(f(r, "# Samples") for r in data)
print(f"shared wavefronts {tw:.0f} = {tw/nelem:.2f}/elem, ideal {sum(f(r,'L1 Wavefronts Shared Ideal') for r in data)/nelem:.2f}, samples {ts:.0f}")
for k in ("stall_mio", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_math", "stall_wait", "stall_lg", "stall_not_selected"):
    print(f"  {k:20s} {sum(f(r,k) for r in data)/ts*100:5.1f}% of samples")
print("idx   instr/elem  wf/elem  ideal  excess  samples%  source")
for i, r in enumerate(data):
    w = f(r, "L1 Wavefronts Shared")
    if w / max(tw, 1) > 0.004 or f(r, "# Samples") / ts > 0.01:
        print(f"{i:4d} {f(r,'Instructions Executed')/nelem:9.3f} {w/nelem:8.3f} {f(r,'L1 Wavefronts Shared Ideal')/nelem:6.3f} "
              f"{f(r,'L1 Wavefronts Shared Excessive')/nelem:6.3f} {f(r,'# Samples')/ts*100:7.2f}  {r[ix['Source']].strip()[:70]}")
