"""Stall-reason samples per region of taxim_kernel.cu from `ncu --page source --print-source cuda,sass --csv` (CUDA view)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
regions = [(78, 162, "hpass"), (163, 190, "hpass-wrap"), (196, 244, "vpass_static"), (245, 301, "vpass"), (302, 340, "reimpose"), (341, 388, "flatcopy"), (389, 430, "blur_level"),
           (431, 454, "flat_rgb"), (455, 592, "prologue/min"), (593, 723, "box/split/restage"), (724, 827, "mask pass"), (828, 883, "between"), (884, 1028, "colour")]
hdr = None; fname = ''
agg = collections.defaultdict(lambda: collections.Counter()); inst = collections.Counter(); tot = 0
lines = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        ln = int(r[0]); d = dict(zip(hdr, r))
        reg = fname
        if fname == 'taxim_kernel.cu':
            reg = next((n for a, b, n in regions if a <= ln <= b), 'other')
        try: smp = int(float(d['# Samples'])); ins = int(float(d['Instructions Executed']))
        except ValueError: continue
        inst[reg] += ins; tot += smp
        for k, v in d.items():
            if k.startswith('stall_') and 'Not Issued' not in k:
                try: agg[reg][k] += int(float(v))
                except ValueError: pass
        lines[(fname, ln, r[1].strip()[:80])] += smp
ti = sum(inst.values())
print("total samples", tot, "inst", ti)
for reg in sorted(agg, key=lambda k: -sum(agg[k].values())):
    s = sum(agg[reg].values())
    top = ", ".join(f"{k[6:]} {100*v/s:.0f}%" for k, v in agg[reg].most_common(5) if s)
    print(f"{reg:28s} smp {100*s/tot:5.1f}%  inst {100*inst[reg]/ti:5.1f}%   {top}")
if len(sys.argv) > 2:
    for (f, ln, src), s in lines.most_common(int(sys.argv[2])):
        print(f"{100*s/tot:5.1f}%  {f}:{ln}  {src}")
