set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python scripts/gpu_quick.py > gpurun_out/quick.log 2>&1
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'grid_rows|compact_rows|comp_|cn_count' -o gpurun_out/prof_r1d -f python scripts/ncu_c2.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -12 gpurun_out/quick.log; tail -3 gpurun_out/ncu_full.log
