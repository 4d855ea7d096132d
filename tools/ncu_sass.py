"""Aggregate the SASS-level source page of an .ncu-rep: python tools/ncu_sass.py <rep> [--list]
Prints executed warp-instructions per opcode and per region (regions split at BAR.SYNC), shared-memory
wavefronts per instruction, and the top stall samples."""
import csv, io, subprocess, sys, collections, re

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first line is the kernel name
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def num(x):
    try: return float(x.replace(",", ""))
    except: return 0.0
tot = collections.Counter(); region = 0; reg_tot = collections.Counter(); reg_wave = collections.Counter(); reg_samp = collections.Counter()
reg_thr = collections.Counter()
insts = []
for r in rows[1:]:
    if len(r) < len(hdr): continue
    sass = r[ix["Source"]].strip()
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op0 = op.split(".")[0]
    ex = num(r[ix["Instructions Executed"]]); thr = num(r[ix["Thread Instructions Executed"]])
    wv = num(r[ix["L1 Wavefronts Shared"]]); wi = num(r[ix["L1 Wavefronts Shared Ideal"]])
    smp = num(r[ix["# Samples"]])
    tot[op0] += ex
    reg_tot[region] += ex; reg_wave[region] += wv; reg_samp[region] += smp; reg_thr[region] += thr
    insts.append((region, sass, ex, thr, wv, wi, smp))
    if op0 == "BAR": region += 1
T = sum(tot.values())
print(f"total warp-instructions {T:.4g}")
for k, v in tot.most_common(25): print(f"  {k:10s} {v:12.4g} {100*v/T:5.1f}%")
print("regions (split at BAR.SYNC): warp-inst, thread-inst/32, shared wavefronts, samples")
for k in sorted(reg_tot): print(f"  region {k}: {reg_tot[k]:12.4g} {reg_thr[k]/32:12.4g} {reg_wave[k]:12.4g} {reg_samp[k]:8.0f}")
print("shared-memory instructions (region, wavefronts, ideal, executed):")
for reg, sass, ex, thr, wv, wi, smp in insts:
    if wv > 0: print(f"  r{reg} {wv:12.4g} {wi:12.4g} {ex:12.4g}  {sass[:90]}")
if "--list" in sys.argv:
    for reg, sass, ex, thr, wv, wi, smp in insts:
        print(f"r{reg} {ex:10.4g} {thr/ max(ex,1):5.1f} {smp:6.0f}  {sass[:110]}")
