for g in copy p2p; do for d in 1 2 3; do
echo "== gather $g inflight $d"
ISB_BENCH_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cfg3 --no-e2e --no-cpu-baseline --gather $g --inflight $d 2>&1 | grep -E "debug|ms_per_step" | cut -c1-200
done; done
