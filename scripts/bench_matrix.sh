# usage: bash scripts/bench_matrix.sh <n_gpus> <tag> [workloads...]
# Runs bench.py for the given workloads on <n_gpus> GPUs of this box and leaves one JSON line per run in
# gpurun_out/<tag>_bench_<workload>[_g<n>].json (copy to profiles/ to commit).
set -u
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
for w in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --workload $w > gpurun_out/${TAG}_bench_${w}.json 2> gpurun_out/${TAG}_bench_${w}.err
    tail -1 gpurun_out/${TAG}_bench_${w}.err | cut -c1-200; cut -c1-160 gpurun_out/${TAG}_bench_${w}.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --workload $w \
      > gpurun_out/${TAG}_bench_${w}_g${N}.json 2> gpurun_out/${TAG}_bench_${w}_g${N}.err
    tail -1 gpurun_out/${TAG}_bench_${w}_g${N}.err | cut -c1-200; cut -c1-160 gpurun_out/${TAG}_bench_${w}_g${N}.json
  fi
done
