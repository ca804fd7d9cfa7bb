cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/h2d_bw.py; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --images 100 --steps 2 --c4-rows 4000000 > gpurun_out/c13_bench2.log 2>&1; echo rc=$?; tail -2 gpurun_out/c13_bench2.log | grep -o '"c5_train_dp.*' | cut -c1-1500
