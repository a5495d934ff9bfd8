#!/usr/bin/env python3
"""SASS mnemonic counts per kernel of libuwcuda.so (the .so is git-ignored; the counts that back DESIGN.md's claims --
FFMA.SAT clamp, bulk async copies + mbarriers in the classify kernel, 16-byte stores in the PEER kernels -- are kept
under profiles/).    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "underwaterworld_b200", "lib", "libuwcuda.so")


def main():
    txt = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
    funcs = re.split(r"\n\s*Function : ", txt)
    print("# SASS evidence: cuobjdump -sass underwaterworld_b200/lib/libuwcuda.so (sm_100a cubin), mnemonic counts per kernel.")
    print("# PEER kernels (…Lb1…) = outputs in another GPU's memory: staged 16-byte vector stores (more STG.128, more STS/LDS).")
    print("#")
    print(f"# {'kernel':64s} {'instr':>6s} {'FFMA':>5s} {'F.SAT':>5s} {'LDS':>4s} {'STS':>4s} {'STG128':>6s} {'STG64':>5s} {'BAR':>4s} {'UBLKCP':>6s} {'SYNCS':>5s} {'ATOM':>5s} {'FP64':>5s}")
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        ins = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", f, flags=re.M)
        try:
            short = subprocess.check_output(["c++filt", name], text=True).strip()
        except Exception:
            short = name
        short = re.sub(r"\(.*", "", short).replace("void ", "").replace("unsigned short", "u16").replace("unsigned int", "u32")
        if not short.startswith("k_"):
            continue
        cnt = lambda p: sum(1 for i in ins if re.search(p, i))
        print(f"  {short:64s} {len(ins):6d} {cnt(r'FFMA'):5d} {cnt(r'FFMA\.SAT'):5d} {cnt(r'LDS'):4d} {cnt(r'STS'):4d} {cnt(r'STG\.E\.128'):6d} "
              f"{cnt(r'STG\.E\.64'):5d} {cnt(r'BAR\.'):4d} {cnt(r'UBLKCP'):6d} {cnt(r'SYNCS'):5d} {cnt(r'ATOM|RED\.'):5d} {cnt(r'D(FMA|MUL|ADD)'):5d}")
    lines = txt.splitlines()
    print("#\n# examples")
    for pat in (r"FFMA\.SAT", r"UBLKCP", r"SYNCS\.", r"STG\.E\.128", r"MEMBAR\.SC\.SYS|MEMBAR.*SYS"):
        for l in [x.strip() for x in lines if re.search(pat, x)][:2]:
            print("#   " + l[:160])


if __name__ == "__main__":
    main()
