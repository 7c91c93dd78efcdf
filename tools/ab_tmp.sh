set -x
timeout 600 python -m pytest tests/test_multigpu.py tests/test_gpu_parity_large.py tests/test_gpu_parity.py -m gpu -x -q -k "sharded or two_ranks or multigpu or mgpu" 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for w in standin16x16_brick standin8x8_brick; do
KMCB200_COMM_LL=0 timeout 300 $TR tools/pcg_micro.py $w
KMCB200_COMM_LL=1 timeout 300 $TR tools/pcg_micro.py $w
done
KMCB200_COMM_LL=1 KMCB200_PCG_PROFILE=1 timeout 300 $TR tools/pcg_micro.py standin8x8_brick 2>&1 | grep "pcg profile" | tail -2
