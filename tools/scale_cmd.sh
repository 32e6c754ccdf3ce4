N=$1
mkdir -p gpurun_out/r2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/pcie_ceiling.py > gpurun_out/r2/pcie_ceiling_${N}gpu.jsonl 2> gpurun_out/r2/pcie_ceiling_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 1000 --warmup 10 > gpurun_out/r2/scale_${N}gpu.json 2> gpurun_out/r2/scale_${N}gpu.err
tail -2 gpurun_out/r2/pcie_ceiling_${N}gpu.jsonl; tail -c 1500 gpurun_out/r2/scale_${N}gpu.json
