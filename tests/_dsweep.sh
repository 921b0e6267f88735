mkdir -p gpurun_out
for cfg in "8 0 56 0" "8 2 56 0" "8 3 56 0" "16 0 56 0" "8 2 56 1" "8 4 56 0"; do
  set -- $cfg
  echo "MAXCL=$1 NSPLIT=$2 SMEM_KB=$3 P1=$4"
  CLIPPPO_DISTURB_MAXCL=$1 CLIPPPO_DISTURB_NSPLIT=$2 CLIPPPO_DISTURB_SMEM_KB=$3 CLIPPPO_DISTURB_P1=$4 python tests/bench_kernels.py disturb 2>&1 | grep -E "B=4096 C=3 224x224|B=16384" | cut -c1-125
done
