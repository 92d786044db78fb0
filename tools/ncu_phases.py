"""Summarise an `ncu --page source --csv --print-source sass` dump: instruction and stall-sample
share per run of SASS instructions with the same execution count (= per phase / loop)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iI, iS, iSamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = [(r[iS].strip(), float(r[iI]), float(r[iSamp])) for r in rows[2:] if len(r) > iI]
tot, totS = sum(d[1] for d in data), sum(d[2] for d in data)
nelem = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"kernel: {rows[0][1][:100]}")
print(f"warp instructions {tot:.0f} ({tot / nelem:.1f} per element), samples {totS:.0f}, static {len(data)}")
def opcode(s):
    t = s.split()
    o = t[1] if t[0].startswith('@') else t[0]
    return o.split('.')[0]
start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][1] - data[start][1]) > 0.25 * max(data[start][1], 1):
        ins = sum(d[1] for d in data[start:i]); sm = sum(d[2] for d in data[start:i])
        if ins / tot > 0.01 or sm / totS > 0.01:
            fp = sum(d[1] for d in data[start:i] if opcode(d[0]) in ("DFMA", "DMUL", "DADD"))
            ls = sum(d[1] for d in data[start:i] if opcode(d[0]) in ("LDS", "STS", "LDG", "STG", "LDSM"))
            hot = max(data[start:i], key=lambda d: d[2])
            print(f"[{start:4d},{i:4d}) x{data[start][1]:9.0f}  instr {ins / tot * 100:5.1f}%  samples {sm / totS * 100:5.1f}%  fp64 {fp / tot * 100:4.1f}%  ld/st {ls / tot * 100:4.1f}%  hottest: {hot[0][:44]} ({hot[2] / totS * 100:.1f}%)")
        start = i
