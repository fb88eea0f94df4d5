#!/bin/bash
# Strong scaling of the BASELINE still (book1 1080p x 100 spp) on one 8-GPU box, final code of round 2: N = 1 and N = 8
# (the GPU-minute budget of the round did not leave room for N = 2 and 4: an N-GPU call is charged 8x).
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_scale_n1.log 2>&1; tail -1 gpurun_out/r02x_scale_n1.log | cut -c1-200
for n in 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_scale_n$n.log 2>&1
  echo "N=$n rc=$?"; grep '^{"metric"' gpurun_out/r02x_scale_n$n.log | tail -1 | cut -c1-200
done
