# A/B of traversal-loop experiment builds (make -C real-time-opencl-raytracer_b200/csrc variant V=...)
B=real-time-opencl-raytracer_b200/csrc
export RTB_PROBE_PARTS=1,8
for V in "" PREFETCH_FAR LEAF_PIPE PREFETCH_FAR+LEAF_PIPE; do
  if [ -z "$V" ]; then L=$B/librtb200.so; else L=$B/build/librtb200_$V.so; fi
  echo "== variant '${V:-base}'"
  RTB200_LIB=$L timeout 100 python tools/timeline_probe.py /tmp/x.json 3840 2160 2>&1 | grep "^primary\|^shaded" | python -c "
import sys,json
for l in sys.stdin:
    k,_,j=l.partition(' '); d=json.loads(j)
    print('  4K',k.ljust(15),'n1 %.4f  n8 max %.4f mean %.4f  noflush %.4f  x%.2f'%(d['n1']['max_ms'],d['n8']['max_ms'],d['n8']['mean_ms'],d['n8_no_flush_ms'],d['speedup_n8']))
"
done
python tools/ab_time.py $B/librtb200.so $B/build/librtb200_PREFETCH_FAR.so 3
python tools/ab_time.py $B/librtb200.so $B/build/librtb200_LEAF_PIPE.so 3
