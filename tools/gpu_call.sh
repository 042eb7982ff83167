# What a gpurun call of this repo runs (edit per call):  gpurun --timeout 2400 -- 'bash tools/gpu_call.sh'
# This version is the round's final single-GPU validation: GPU tests, the driver's two bench arms, the smoke test.
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_final_ref.log 2> gpurun_out/bench_final_ref.err; cut -c1-400 gpurun_out/bench_final_ref.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
