mkdir -p gpurun_out/r02g
for opt in "--reserve-sms 0" "--reserve-sms 8" "--reserve-sms 16" "--no-shard" "--no-shard --reserve-sms 16"; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-c3 --no-e2e --no-parity $opt 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$opt', round(d['value'],1), round(d['ms_per_step'],3))
"
done
