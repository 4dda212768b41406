set -x
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_configs.py c4 primary -1 6 > gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:trace_kernel -s 5 -c 1 -o gpurun_out/r2_c4_primary python tools/prof_configs.py c4 primary -1 6 > gpurun_out/r2_prof_ncu.log 2>&1
python tools/prof_configs.py c4 shadow -1 6 >> gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:trace_kernel -s 5 -c 1 -o gpurun_out/r2_c4_shadow python tools/prof_configs.py c4 shadow -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
python tools/prof_configs.py c4 fused -1 6 >> gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:primary_shadow_kernel -s 4 -c 1 -o gpurun_out/r2_c4_fused python tools/prof_configs.py c4 fused -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
grep -h median gpurun_out/r2_prof_plain.log
