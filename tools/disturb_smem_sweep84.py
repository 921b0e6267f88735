"""Stripes per image (CLIPPPO_DISTURB_SMEM_KB) x blur row splits x CTA size on the 84 x 84 shapes, after the cluster barrier became cheap."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
CHILD = open(os.path.join(HERE, 'disturb_nthreads_sweep.py')).read().split("CHILD = r'''")[1].split("'''")[0]
S84 = "[(16384,3,84,84,'SEVERE'),(16384,3,84,84,'HARD'),(16384,3,84,84,'MODERATE'),(16384,3,84,84,'MILD'),(16384,1,84,84,'HARD'),(16384,1,84,84,'SEVERE'),(64,3,84,84,'SEVERE'),(1024,3,84,84,'SEVERE')]"
for kb in (113, 56, 40, 24, 16):
    for ns in (0, 1, 2, 3, 4):
        for nt in ((0,) if ns == 0 else (0, 128, 192)):
            print(f"SMEM_KB={kb} NSPLIT={ns} NTHREADS={nt}", flush=True)
            env = dict(os.environ, CLIPPPO_DISTURB_SMEM_KB=str(kb), CLIPPPO_DISTURB_NSPLIT=str(ns), CLIPPPO_DISTURB_NTHREADS=str(nt), SHAPES=S84)
            subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
