# round-2 ncu captures (one gpurun call): launch list of the bench command + --set full of the kernels VERDICT r1 named
set -x
NCU="ncu --set full --clock-control none --import-source on"
python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps20.csv python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/r2_ncu_bench.log 2>&1
python tools/prof_configs.py c2 primary -1 6 > gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:trace_kernel -s 5 -c 1 -o gpurun_out/r2_c2_primary python tools/prof_configs.py c2 primary -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
python tools/prof_configs.py c4 primary -1 6 >> gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:trace_kernel -s 5 -c 1 -o gpurun_out/r2_c4_primary python tools/prof_configs.py c4 primary -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
python tools/prof_configs.py c4 shadow -1 6 >> gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:trace_kernel -s 5 -c 1 -o gpurun_out/r2_c4_shadow python tools/prof_configs.py c4 shadow -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
python tools/prof_configs.py c4 fused -1 6 >> gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:primary_shadow_kernel -s 4 -c 1 -o gpurun_out/r2_c4_fused python tools/prof_configs.py c4 fused -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
python tools/prof_configs.py c4 diffuse -1 6 >> gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:trace_lanes_kernel -s 4 -c 1 -o gpurun_out/r2_c4_diffuse python tools/prof_configs.py c4 diffuse -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
python tools/prof_configs.py c1 primary -1 6 >> gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:trace_kernel -s 5 -c 1 -o gpurun_out/r2_c1_primary python tools/prof_configs.py c1 primary -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
python tools/prof_configs.py c2 fused -1 6 >> gpurun_out/r2_prof_plain.log 2>&1 &&
$NCU -k regex:primary_shadow_kernel -s 4 -c 1 -o gpurun_out/r2_c2_fused python tools/prof_configs.py c2 fused -1 6 >> gpurun_out/r2_prof_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep; grep -h "median" gpurun_out/r2_prof_plain.log gpurun_out/r2_prof_ncu.log
