mkdir -p gpurun_out
for cl in 8 16; do for p1 in 0 1; do for kb in 56 113; do
  echo "MAXCL=$cl P1=$p1 SMEM_KB=$kb"
  CLIPPPO_DISTURB_MAXCL=$cl CLIPPPO_DISTURB_P1=$p1 CLIPPPO_DISTURB_SMEM_KB=$kb python tests/bench_kernels.py disturb 2>&1 | grep -E "B=4096 C=3 224x224 MODERATE|B=16384 C=3 84x84" | cut -c1-120
done; done; done
