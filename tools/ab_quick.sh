# quick A/B of build/librtb200_PREV.so against the current library
B=real-time-opencl-raytracer_b200/csrc
python tools/ab_time.py $B/build/librtb200_PREV.so $B/librtb200.so 3
export RTB_PROBE_PARTS=1,8
for L in $B/build/librtb200_PREV.so $B/librtb200.so; do
  echo "== $L 4K"
  RTB200_LIB=$L timeout 100 python tools/timeline_probe.py /tmp/x.json 3840 2160 2>&1 | grep "^primary\|^shaded" | python -c "
import sys,json
for l in sys.stdin:
    k,_,j=l.partition(' '); d=json.loads(j)
    print('   ',k.ljust(15),'n1 %.4f  n8 max %.4f mean %.4f  noflush %.4f  x%.2f'%(d['n1']['max_ms'],d['n8']['max_ms'],d['n8']['mean_ms'],d['n8_no_flush_ms'],d['speedup_n8']))
"
done
