"""python tools/scratch/e2e_ab.py libA.so libB.so : host-API timings per library, interleaved, same box"""
import os, subprocess, sys
for rnd in range(3):
    for l in sys.argv[1:]:
        env = dict(os.environ); env["UWCUDA_LIB"] = os.path.abspath(l)
        out = subprocess.run([sys.executable, "tools/scratch/e2e_pipe.py"], env=env, capture_output=True, text=True)
        lines = out.stdout.strip().splitlines()
        print(os.path.basename(l), " | ".join(x.split("us")[0] for x in lines[-3:]))
