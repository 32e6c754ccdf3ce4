#!/bin/bash
# tools/scale_cmd.sh N [nopcie]: the host's transfer ceiling and the bench at N GPUs of one box (gpurun --gpus N)
N=$1
mkdir -p gpurun_out/r2
if [ "$2" != "nopcie" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/pcie_ceiling.py > gpurun_out/r2/pcie_ceiling_${N}gpu.jsonl 2> gpurun_out/r2/pcie_ceiling_${N}gpu.err
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 1000 --warmup 10 > gpurun_out/r2/scale_${N}gpu.json 2> gpurun_out/r2/scale_${N}gpu.err
tail -2 gpurun_out/r2/pcie_ceiling_${N}gpu.jsonl; tail -c 1500 gpurun_out/r2/scale_${N}gpu.json
