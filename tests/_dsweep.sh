mkdir -p gpurun_out
for cfg in "0 0 16" "1 0 16" "2 0 16" "4 0 16" "16 0 16" "8 2 8" "8 1 8"; do
  set -- $cfg
  echo "IPC=$1 NSPLIT=$2 MAXCL=$3"
  CLIPPPO_DISTURB_IPC=$1 CLIPPPO_DISTURB_NSPLIT=$2 CLIPPPO_DISTURB_MAXCL=$3 python tests/bench_kernels.py disturb 2>&1 | grep -E "B=4096 C=3 224x224|B=501" | cut -c1-125
done
