#!/usr/bin/env python3
"""Per-source-line executed-instruction / stall-sample totals for one kernel of an .ncu-rep.

Joins `ncu --page source --csv` (SASS view: per-instruction counters, in program order) with
`nvdisasm --print-line-info` of the in-tree library (same program order) so that counters can be
summed per CUDA source line.  Usage:
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_noise_small [mangled-substring] [--top 40]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("UWCUDA_LIB") or os.path.join(ROOT, "underwaterworld_b200", "lib", "libuwcuda.so")


def sass_lines(mangled_sub):
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=td, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.check_output(["nvdisasm", "--print-line-info", os.path.join(td, cub)], text=True)
    out, cur, infn, inl = [], None, False, None
    for ln in txt.splitlines():
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            infn = mangled_sub in m.group(1)
            cur = None
            continue
        if ln.startswith("//-----") or ln.startswith("\t.section"):
            if infn and out and ln.startswith("//-----"):
                infn = False
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = int(m.group(2))
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append(cur)
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    mangled = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else kern
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                                  text=True, stderr=subprocess.DEVNULL)
    # first kernel instance only
    blocks = raw.split('"Kernel Name"')
    body = '"Kernel Name"' + blocks[1]
    rows = list(csv.reader(io.StringIO(body)))
    hdr = rows[1]
    ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    inst = [(r[ti], int(r[ii] or 0), int(r[si] or 0)) for r in rows[2:] if len(r) > ii and r[0].strip().isdigit() or (len(r) > ii and re.match(r"^[0-9a-fx]+$", r[0].strip()))]
    lines = sass_lines(mangled)
    if len(lines) != len(inst):
        print(f"warning: {len(lines)} SASS instructions in the library vs {len(inst)} in the report", file=sys.stderr)
    n = min(len(lines), len(inst))
    per = {}
    for k in range(n):
        a = per.setdefault(lines[k], [0, 0, 0])
        a[0] += inst[k][1]; a[1] += inst[k][2]; a[2] += 1
    src = open(os.path.join(ROOT, "underwaterworld_b200", "csrc", "uw_kernels.cuh")).read().splitlines()
    tot_i = sum(v[0] for v in per.values()); tot_s = sum(v[1] for v in per.values())
    print(f"kernel {kern}: {n} SASS instrs, {tot_i} warp-instructions executed, {tot_s} stall samples")
    print(f"{'line':>5} {'winst':>10} {'%':>5} {'samples':>8} {'%':>5} {'sass':>5}  source")
    for ln, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        s = src[ln - 1].strip()[:100] if ln and ln <= len(src) else "?"
        print(f"{ln or 0:5d} {v[0]:10d} {100*v[0]/max(tot_i,1):5.1f} {v[1]:8d} {100*v[1]/max(tot_s,1):5.1f} {v[2]:5d}  {s}")


if __name__ == "__main__":
    main()
