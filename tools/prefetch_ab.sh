# A/B: queue prefetch on/off, occupancy variants; sched tests first
B=real-time-opencl-raytracer_b200/csrc
(timeout 300 python -m pytest tests/test_gpu_sched.py tests/test_gpu_fused.py -x -q 2>&1 | tail -3)
export RTB_PROBE_PARTS=1,8
for V in "" NO_QUEUE_PREFETCH; do
  if [ -z "$V" ]; then L=$B/librtb200.so; else L=$B/build/librtb200_$V.so; fi
  for R in "1920 1080" "3840 2160"; do set -- $R
  echo "== variant '${V:-base}' $1x$2"
  RTB200_LIB=$L timeout 100 python tools/timeline_probe.py /tmp/x.json $1 $2 2>&1 | grep "^primary\|^shaded" | python -c "
import sys,json
for l in sys.stdin:
    k,_,j=l.partition(' '); d=json.loads(j)
    print('   ',k.ljust(15),'n1 %.4f  n8 max %.4f mean %.4f  noflush %.4f  x%.2f'%(d['n1']['max_ms'],d['n8']['max_ms'],d['n8']['mean_ms'],d['n8_no_flush_ms'],d['speedup_n8']))
"
  done
done
python tools/ab_time.py $B/build/librtb200_NO_QUEUE_PREFETCH.so $B/librtb200.so 3
for V in MINB10 MINB12; do
RTB200_LIB=$B/build/librtb200_$V.so python tools/prof_configs.py c4 primary -1 8 | tail -1
RTB200_LIB=$B/build/librtb200_$V.so python tools/prof_configs.py c2 primary -1 8 | tail -1
done
python tools/prof_configs.py c4 primary -1 8 | tail -1
python tools/prof_configs.py c2 primary -1 8 | tail -1
# timeline with hints on: shape of the 1080p launch now
mkdir -p gpurun_out/tl1080h
RTB_PROBE_PARTS=1 RTB200_LIB=$B/build/librtb200_timeline.so timeout 100 python tools/timeline_probe.py gpurun_out/r2_tl_hints_1080.json 1920 1080 gpurun_out/tl1080h 2>&1 | grep timeline_n1 | cut -c1-1500
