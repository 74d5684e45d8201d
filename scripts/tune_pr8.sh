#!/bin/bash
# PageRank section only on N GPUs (run on the GPU box): bash scripts/tune_pr8.sh N "weights"
N=$1; port=29600
for w in $2; do port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --only-pagerank --no-cpu-baseline --row-weight $w > gpurun_out/pr_n${N}_w$w.log 2>&1
  grep "^{" gpurun_out/pr_n${N}_w$w.log | python -c "
import json,sys; d=json.loads(sys.stdin.read())
for k,v in d['extra'].items():
    if 'pagerank' in k: print('N=$N w=$w', k, round(v['ms_per_iter'],3), 'ms/iter', v['partition'], 'max rows/rank', v['rows_per_rank_max'])
    elif 'rmat' in k: print('N=$N w=$w', k, round(v['ms'],3))
"
done
