set -x
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_configs.py c2 frame -1 6 > gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:render_kernel -s 4 -c 1 -o gpurun_out/r2_c2_frame python tools/prof_configs.py c2 frame -1 6 > gpurun_out/r2_prof_ncu.log 2>&1
grep -h median gpurun_out/r2_prof_plain.log
