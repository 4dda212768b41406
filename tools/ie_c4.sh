B=real-time-opencl-raytracer_b200/csrc
for V in IE4 IE12 IE8; do
  L=$B/build/librtb200_$V.so
  echo "== $V"
  for P in primary fused shadow frame; do RTB200_LIB=$L timeout 120 python tools/prof_configs.py c4 $P -1 8 | tail -1; done
done
echo "== base"; for P in shadow frame; do timeout 120 python tools/prof_configs.py c4 $P -1 8 | tail -1; done
