B=real-time-opencl-raytracer_b200/csrc
for V in "" L2_PREFETCH_128 L2_PREFETCH_256; do
  if [ -z "$V" ]; then L=$B/librtb200.so; else L=$B/build/librtb200_$V.so; fi
  echo "== ${V:-base}"
  for P in primary shadow fused diffuse; do RTB200_LIB=$L python tools/prof_configs.py c4 $P -1 8 | tail -1; done
  RTB200_LIB=$L python tools/prof_configs.py c2 primary -1 8 | tail -1
  RTB200_LIB=$L python tools/prof_configs.py c2 diffuse -1 8 | tail -1
done
