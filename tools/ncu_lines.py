"""Per-CUDA-line sample / instruction totals from `ncu --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
out = []; fname = ''
tot_s = tot_i = 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        try:
            s = int(float(r[6])); n = int(float(r[7]))
        except ValueError:
            continue
        out.append((s, n, fname, int(r[0]), r[1].strip()[:90])); tot_s += s; tot_i += n
out.sort(reverse=True)
print('samples', tot_s, 'inst', tot_i)
for s, n, f, ln, src in out[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f'{100*s/tot_s:5.1f}% smp {100*n/tot_i:5.1f}% inst  {f}:{ln}  {src}')
