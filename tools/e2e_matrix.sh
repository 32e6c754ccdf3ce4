for N in 1 2; do for bind in "" "--no-bind"; do for r in 1 3; do
 if [ $N = 1 ]; then out=$(python bench.py --steps 200 --warmup 5 --no-cpu-baseline --host-route $r $bind 2>/dev/null); else out=$(python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 200 --warmup 5 --host-route $r $bind 2>/dev/null | grep "^{"); fi
 echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N bind=[$bind] route=$r whole', round(d['e2e']['ms_per_step'],3) if 'soil_step' in d['e2e']['api'] else round(d['e2e']['other_route']['ms_per_step'],3), 'stage', round(d['e2e']['other_route']['ms_per_step'],3) if 'soil_step' in d['e2e']['api'] else round(d['e2e']['ms_per_step'],3))"
done; done; done
