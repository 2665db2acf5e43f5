"""Per-source-line view of an .ncu-rep captured with --import-source on (no GPU needed): share of the warp-stall samples
and of the executed warp instructions per CUDA source line, top lines first.
usage: ncu_lines.py rep [csrc-dir] [top-n]"""
import csv, io, os, subprocess, sys

rep = sys.argv[1]
csrc = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "phybin_b200", "csrc")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg, cur, cur_file, tot_i, tot_s = {}, None, None, 0, 0
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] in ("Line No", "Function Name", "Kernel Name"):
        if r[0] != "Line No":
            print(r[0] + ":", r[1])
        continue
    if r[0] != "":
        if r[0].isdigit():
            cur = (cur_file, int(r[0]))
        continue
    if len(r) < 8 or r[2] == "...":
        continue
    try:
        samples, inst = int(r[4]), int(r[7])
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0]); a[0] += samples; a[1] += inst
    tot_i += inst; tot_s += samples
print(f"warp-stall samples {tot_s}, executed warp instructions {tot_i} (instructions inlined into several lines are counted at each)")
cache = {}
def line_of(f, l):
    if f not in cache:
        try:
            cache[f] = open(os.path.join(csrc, f)).read().split("\n")
        except OSError:
            cache[f] = []
    return cache[f][l - 1].strip()[:100] if 0 < l <= len(cache[f]) else ""
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l:<4d} samples {100 * v[0] / max(tot_s, 1):5.1f}%  inst {100 * v[1] / max(tot_i, 1):5.1f}%  | {line_of(f, l)}")
