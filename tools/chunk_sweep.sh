mkdir -p gpurun_out
for c in 1024 1366 2048 4096; do
  echo "CHUNK=$c"
  CLIPPPO_VIT_CHUNK=$c python bench.py --no-cpu-baseline --no-e2e --steps 8 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(' fps %.0f  ms %.2f  gemm TF %.0f frac %.3f share %.3f launches %d' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], r['share_of_step'], d['gpu_launches']))
        print('  ', [(b['N'],b['K'],b['tflops']) for b in r['by_shape']])
"
done
