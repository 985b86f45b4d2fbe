# usage: bash scripts/run_gpu_check.sh [tests] [quick] [bench] [ncu]
set -x
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log;;
quick) python tests/checks/gpu_quick.py > gpurun_out/quick.log 2>&1; tail -14 gpurun_out/quick.log;;
bench) python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cut -c1-400 gpurun_out/bench_c2.json;;
ncu) ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'grid_|comp_|cn_count|pack_rows|fused_|bbox|cell_id|gather_kernel' -o gpurun_out/prof -f python scripts/ncu_c2.py > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log;;
ncu_grid) COUNT=0 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:grid_rows -o gpurun_out/prof_grid -f python scripts/ncu_c2.py > gpurun_out/ncu_grid.log 2>&1; tail -2 gpurun_out/ncu_grid.log;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/ncu_c2.py > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log;;
distmat) python tests/checks/distmat_bench.py > gpurun_out/distmat.json 2> gpurun_out/distmat.err; tail -3 gpurun_out/distmat.err; cut -c1-1500 gpurun_out/distmat.json;;
periodic) python scripts/periodic_bench.py > gpurun_out/periodic.json 2> gpurun_out/periodic.err; tail -3 gpurun_out/periodic.err; cut -c1-2500 gpurun_out/periodic.json;;
esac
done
