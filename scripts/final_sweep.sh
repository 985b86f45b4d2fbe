# usage: bash scripts/final_sweep.sh <tag>     (one GPU) -- the GPU suite, one bench line per workload, the reference arm, launch lists
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
bash scripts/bench_matrix.sh 1 $TAG c2 c1 c3 c4 c5 2>&1 | cut -c1-200
for w in c2 c1; do python bench.py --impl reference --workload $w > gpurun_out/${TAG}_bench_${w}_reference.json 2> gpurun_out/${TAG}_bench_${w}_reference.err; cut -c1-200 gpurun_out/${TAG}_bench_${w}_reference.json; done
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_c2.csv python scripts/ncu_c2.py > gpurun_out/${TAG}_l2.log 2>&1
N=200000 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_c3.csv python scripts/ncu_c3.py > gpurun_out/${TAG}_l3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches_bench_py.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_lb.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
