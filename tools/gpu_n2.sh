#!/bin/bash
# 2-GPU visit: data-parallel bench (SyncBN exchange inside the BN kernels, overlapped arena all-reduce) + peer all-reduce check.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "peer or exchange or sync_batchnorm" > gpurun_out/pytest_peer.log 2>&1; echo "pytest peer rc=$?"; tail -3 gpurun_out/pytest_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-300
tail -5 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    tools/test_peer_allreduce.py > gpurun_out/peer_n$N.log 2>&1; echo "peer rc=$?"; tail -$N gpurun_out/peer_n$N.log
