B=real-time-opencl-raytracer_b200/csrc
for V in IE4 IE8 IE12; do
  timeout 200 python tools/ab_time.py $B/librtb200.so $B/build/librtb200_$V.so 2
done
for V in "" IE8; do
  if [ -z "$V" ]; then L=$B/librtb200.so; else L=$B/build/librtb200_$V.so; fi
  echo "== ${V:-base}"
  RTB200_LIB=$L timeout 120 python tools/prof_configs.py c4 primary -1 8 | tail -1
  RTB200_LIB=$L timeout 120 python tools/prof_configs.py c4 fused -1 8 | tail -1
done
