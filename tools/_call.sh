cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
SECONDS=0; python bench.py --steps 5 > gpurun_out/c27_bench_default.log 2> gpurun_out/c27_bench_default.err; echo rc=$? wall=$SECONDS s; tail -1 gpurun_out/c27_bench_default.log | cut -c1-300
SECONDS=0; python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c27_bench_ref.log 2> gpurun_out/c27_bench_ref.err; echo rc=$? wall=$SECONDS s; tail -1 gpurun_out/c27_bench_ref.log | cut -c1-300
