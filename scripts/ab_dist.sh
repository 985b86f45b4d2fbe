for combo in ${COMBOS:-"0 0" "0 1" "250 1"}; do set -- $combo; 
CNNB200_BALANCE_ROW=$1 CNNB200_DIST_WINDOWS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $G --workload c2 2>/dev/null > gpurun_out/ab.json; python -c "
import json
d=json.loads([l for l in open('gpurun_out/ab.json') if l.startswith('{')][-1])
print('K=$1 win=$2', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'unstaged', round(d.get('unstaged_ms_per_step',0),3), 'flatten', round(d['stage_ms_per_step']['comp_flatten'],3))
"; done
