#!/usr/bin/env python3
"""Per-REGION (source line ranges) executed instructions and stall reasons of one kernel of an .ncu-rep.
Usage: python tools/ncu_regions.py rep.ncu-rep kernel_regex mangled_substring name:lo-hi [name:lo-hi ...]
Lines outside every region are reported as 'other'.  Same SASS/line join as tools/ncu_lines.py."""
import csv, io, re, subprocess, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from ncu_lines import sass_lines

STALLS = ["stall_barrier", "stall_short_sb", "stall_long_sb", "stall_wait", "stall_no_inst", "stall_not_selected",
          "stall_selected", "stall_math", "stall_mio", "stall_branch_resolving", "stall_lg", "stall_membar", "stall_dispatch"]


def main():
    rep, kern, mangled = sys.argv[1:4]
    regions = []
    for a in sys.argv[4:]:
        name, rng = a.split(":"); lo, hi = rng.split("-"); regions.append((name, int(lo), int(hi)))
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], text=True, stderr=subprocess.DEVNULL)
    body = '"Kernel Name"' + raw.split('"Kernel Name"')[1]
    rows = list(csv.reader(io.StringIO(body)))
    hdr = rows[1]
    ii = hdr.index("Instructions Executed")
    cols = {s: hdr.index(s) for s in STALLS if s in hdr}
    data = [r for r in rows[2:] if len(r) > ii and re.match(r"^[0-9a-fx]+$", r[0].strip())]
    lines = sass_lines(mangled)
    if len(lines) != len(data):
        print(f"warning: {len(lines)} SASS instructions in the library vs {len(data)} in the report", file=sys.stderr)
    acc = {}
    for k in range(min(len(lines), len(data))):
        ln = lines[k] or 0
        name = next((n for n, lo, hi in regions if lo <= ln <= hi), "other")
        a = acc.setdefault(name, {"inst": 0, "sass": 0, **{s: 0 for s in cols}})
        a["inst"] += int(data[k][ii] or 0); a["sass"] += 1
        for s, c in cols.items():
            a[s] += int(data[k][c] or 0)
    tot_i = sum(a["inst"] for a in acc.values())
    tot_s = sum(sum(a[s] for s in cols) for a in acc.values())
    print(f"{'region':14s} {'winst':>11s} {'%':>5s} {'sass':>5s} {'samples':>8s} {'%':>5s}  top stall reasons (share of the region's samples)")
    for name, a in sorted(acc.items(), key=lambda kv: -kv[1]["inst"]):
        smp = sum(a[s] for s in cols)
        top = sorted(((a[s], s) for s in cols), reverse=True)[:5]
        desc = "  ".join(f"{s.replace('stall_', '')} {100 * v / max(smp, 1):.0f}%" for v, s in top if v)
        print(f"{name:14s} {a['inst']:11d} {100 * a['inst'] / max(tot_i, 1):5.1f} {a['sass']:5d} {smp:8d} {100 * smp / max(tot_s, 1):5.1f}  {desc}")


if __name__ == "__main__":
    main()
