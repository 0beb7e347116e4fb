"""Per-source-line summary of an `ncu --page source --print-source cuda,sass --csv` export (gzip or plain).
    python tools/ncu_lines.py file.csv.gz [top]"""
import csv, gzip, io, sys
fn = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = (gzip.open(fn, 'rt') if fn.endswith('.gz') else open(fn)).read()
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; out = []
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if r and r[0] and r[0].isdigit() and hdr:
        ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); it = hdr.index('Thread Instructions Executed')
        try: out.append((cur, int(r[0]), r[1][:100], int(r[ie]), int(r[isamp]), int(r[it])))
        except ValueError: pass
tot = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print('total warp-instructions', tot, 'samples', ts, 'avg lanes %.1f' % (sum(o[5] for o in out) / tot))
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print('%-16s %4d inst %5.2f%% samp %5.2f%% lanes %4.1f | %s' % (o[0], o[1], 100 * o[3] / tot, 100 * o[4] / ts, o[5] / max(o[3], 1), o[2]))
