# Sweep of the fast disturbance kernel's launch knobs (stripes per image via the smem budget, blur-task splits).
for cfg in "113 0" "113 1" "113 2" "113 3" "113 4" "75 0" "75 2" "75 3" "227 1" "227 2"; do
  set -- $cfg
  echo "SMEM_KB=$1 NSPLIT=$2"
  CLIPPPO_DISTURB_SMEM_KB=$1 CLIPPPO_DISTURB_NSPLIT=$2 python tools/bench_kernels.py disturb 2>&1 | grep -E "B=4096 C=3 224x224|B=16384 C=3" | cut -c1-125
done
