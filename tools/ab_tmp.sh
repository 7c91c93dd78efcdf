set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for w in standin4x4_brick standin2x2_brick; do
KMCB200_PCG_PERSISTENT=0 timeout 300 $TR tools/pcg_micro.py $w
KMCB200_PCG_PERSISTENT=1 timeout 300 $TR tools/pcg_micro.py $w
done
export CUDA_VISIBLE_DEVICES=0
for w in standin4x4_brick standin2x2_brick 5nm; do
KMCB200_PCG_PERSISTENT=0 timeout 300 python tools/pcg_micro.py $w
KMCB200_PCG_PERSISTENT=1 timeout 300 python tools/pcg_micro.py $w
done
