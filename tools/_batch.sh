mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-refine > gpurun_out/r02e_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'encoder_layer|gather_tile|vote_kernel|window_stream|traverse_kernel|roll_from_pairs|box_rows|box_cols' -s 24 -c 12 -o gpurun_out/r02e_full python tools/profile_frame.py --frames 4 > gpurun_out/r02e_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:'rf_' -c 24 -o gpurun_out/r02e_refine python tools/profile_refine.py --frames 1 > gpurun_out/r02e_ncu_refine.log 2>&1
tail -2 gpurun_out/r02e_ncu_full.log gpurun_out/r02e_ncu_refine.log
