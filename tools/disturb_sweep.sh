for cfg in "8 2 6" "8 2 8" "8 2 4" "8 3 6" "8 3 8" "16 1 6" "16 1 8" "16 2 6" "4 2 6"; do
  set -- $cfg
  echo "MAXCL=$1 NSPLIT=$2 UNR=$3"
  CLIPPPO_DISTURB_MAXCL=$1 CLIPPPO_DISTURB_NSPLIT=$2 CLIPPPO_DISTURB_UNR=$3 python tools/bench_kernels.py disturb 2>&1 | grep -E "B=4096 C=3 224x224|B=16384" | cut -c1-125
done
