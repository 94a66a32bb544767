"""Aggregate an ncu source page (--print-source cuda,sass --csv) into per-source-line instruction / sample shares.
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:X > f.csv; python scripts/ncu_lines.py f.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, out = None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0] != "" and len(r) > 7:
        try:
            out.append((int(r[7]), int(r[6]), cur_file, r[0], r[1]))
        except ValueError:
            pass
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print("total warp-instructions", tot, "samples", tots)
for o in sorted(out, reverse=True)[:N]:
    print(f"{o[0]:>10} {100*o[0]/tot:5.1f}% | samples {100*o[1]/max(tots,1):5.1f}% | {o[2]}:{o[3]}: {o[4][:100]}")
