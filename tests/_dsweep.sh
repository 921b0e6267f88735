mkdir -p gpurun_out
for cfg in "0 0" "0 2" "1 0" "4 0" "16 0" "8 2" "8 3"; do
  set -- $cfg
  echo "IPC=$1 NSPLIT=$2"
  CLIPPPO_DISTURB_IPC=$1 CLIPPPO_DISTURB_NSPLIT=$2 python tests/bench_kernels.py disturb 2>&1 | grep -E "B=4096 C=3 224x224|B=501|B=16384 C=3" | cut -c1-125
done
