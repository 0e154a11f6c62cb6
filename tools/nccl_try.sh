run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 200 --warmup 5 --no-extra --no-cpu 2>gpurun_out/nccl_$2.err | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['ms_per_step'], d['value'])"; }
run 29601 base
NCCL_MAX_CTAS=1 NCCL_MIN_CTAS=1 run 29602 ctas1
TORCH_NCCL_HIGH_PRIORITY=1 run 29603 hiprio
NCCL_MAX_CTAS=1 NCCL_MIN_CTAS=1 TORCH_NCCL_HIGH_PRIORITY=1 run 29604 both
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING run 29605 dbg
grep -i "channels\|nvls\|AllReduce" gpurun_out/nccl_dbg.err | head -30
