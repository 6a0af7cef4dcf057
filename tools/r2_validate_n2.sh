#!/bin/bash
# Two-GPU validation: the whole GPU suite (two-device layouts, one process per GPU over CUDA IPC, the C++ class on two
# devices) and the 2-rank bench line with its 1-GPU parity replay.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_n2_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2_n2_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
    > $O/r2_n2_bench.json 2> $O/r2_n2_bench.err; echo "bench rc=$?" >> $O/r2_n2_bench.err
tail -5 $O/r2_n2_tests.log; tail -2 $O/r2_n2_bench.err; head -c 600 $O/r2_n2_bench.json
