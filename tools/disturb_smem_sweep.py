import os, subprocess, sys
CHILD = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'disturb_nthreads_sweep.py')).read().split("CHILD = r'''")[1].split("'''")[0]
S84 = "[(16384,3,84,84,'SEVERE'),(16384,3,84,84,'MODERATE'),(16384,3,84,84,'MILD'),(16384,1,84,84,'HARD')]"
S224 = "[(4096,3,224,224,'SEVERE'),(4096,3,224,224,'MODERATE'),(4096,3,224,224,'MILD')]"
for shapes, kbs, nss in ((S84, (113, 75, 56, 40), (0, 2, 3)), (S224, (113, 75, 56), (0, 1, 2))):
    for kb in kbs:
        for ns in nss:
            print(f"SMEM_KB={kb} NSPLIT={ns}", flush=True)
            env = dict(os.environ, CLIPPPO_DISTURB_SMEM_KB=str(kb), CLIPPPO_DISTURB_NSPLIT=str(ns), SHAPES=shapes)
            subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
