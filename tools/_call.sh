cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_trainer.py -x -q -m gpu 2>&1 | tail -4
for g in 1 0; do MC_MLP_GRAPH=$g timeout 120 python tools/bench_train.py --rows 400000 --epochs 2 --cpu-rows 2000 2>&1 | tail -1 | cut -c1-400; done
