mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --secondary "
$B '' > gpurun_out/s16_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv $B '' > gpurun_out/s16_ncu1.log 2>&1
python tools/profile_target.py S64_disc 4 > gpurun_out/s16_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_sym -s 1 -c 2 -o gpurun_out/r02_pair_sym_final python tools/profile_target.py S64_disc 4 > gpurun_out/s16_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:'k_vertex|k_kick_drift|k_pair_reduce|k_mesh' -s 8 -c 5 -o gpurun_out/r02_small_kernels_final python tools/profile_target.py S64_disc 4 > gpurun_out/s16_ncu3.log 2>&1
