# The measurement batch kept under profiles/: GPU tests, time-loop profile, bench, reference arm, other configurations,
# ncu launch list of the bench command and two `--set full` captures (each ncu run after the same command exited 0 without
# ncu).   bash tools/run_profiles.sh [tag]   -> gpurun_out/<tag>_*
TAG=${1:-r3a}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_ts_tree.py 2>&1 | grep -v -i warn > gpurun_out/${TAG}_time_loop_tree.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python tools/bench_configs.py c3 c4 c5 ts_tree > gpurun_out/${TAG}_other_configs.jsonl 2> gpurun_out/${TAG}_other.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu1.log 2>&1
python tools/profile_step.py > gpurun_out/${TAG}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"step_resident" -s 2 -c 1 -f -o gpurun_out/${TAG}_full_c2 python tools/profile_step.py > gpurun_out/${TAG}_ncu2.log 2>&1
python tools/profile_step.py 17 7 2 > gpurun_out/${TAG}_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|schur_solve" -s 4 -c 4 -f -o gpurun_out/${TAG}_full_scale python tools/profile_step.py 17 7 2 > gpurun_out/${TAG}_ncu3.log 2>&1
tail -2 gpurun_out/${TAG}_time_loop_tree.txt; tail -c 300 gpurun_out/${TAG}_bench.err
