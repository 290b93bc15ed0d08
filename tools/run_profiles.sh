set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_ts_tree.py 2>&1 | grep -v -i warn > gpurun_out/r2y_time_loop_tree.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err
python tools/bench_configs.py c3 c4 c5 ts_tree > gpurun_out/r2y_other_configs.jsonl 2> gpurun_out/r2y_other.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2y_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2y_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2y_ncu1.log 2>&1
python tools/profile_step.py > gpurun_out/r2y_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"step_resident" -s 2 -c 1 -f -o gpurun_out/r2y_full_c2 python tools/profile_step.py > gpurun_out/r2y_ncu2.log 2>&1
python tools/profile_step.py 17 7 2 > gpurun_out/r2y_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|schur_solve" -s 4 -c 4 -f -o gpurun_out/r2y_full_scale python tools/profile_step.py 17 7 2 > gpurun_out/r2y_ncu3.log 2>&1
tail -2 gpurun_out/r2y_time_loop_tree.txt; tail -c 300 gpurun_out/r2y_bench.err
