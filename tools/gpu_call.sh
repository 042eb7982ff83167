set -x
cd $GRAFT_REPO_ROOT
cap() {   # cap <name> <kernel regex> <skip> <count> <updates per launch> <cmd...>
  name=$1; rx=$2; skip=$3; cnt=$4; upd=$5; shift 5
  timeout 900 "$@" > gpurun_out/plain_${name}.log 2>&1 && cat gpurun_out/plain_${name}.log | tail -3 &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/prof_${name} "$@" > gpurun_out/ncu_${name}.log 2>&1
  tail -1 gpurun_out/ncu_${name}.log
  python tools/ncu_summary.py gpurun_out/prof_${name}.ncu-rep --updates $upd --top 25 > gpurun_out/summary_${name}.md 2> gpurun_out/summary_${name}.err
  rm -f gpurun_out/prof_${name}.ncu-rep
}
# 1. K1: how much of its time is the look-back chain?  (offsets are wrong with the flag: timing only)
timeout 200 python tools/index_probe.py --reps 2 --modes pinned32 > gpurun_out/k1_lookback_on_r2i.log 2>&1; grep -o "K1 [0-9.]* ms" gpurun_out/k1_lookback_on_r2i.log
GFASORT_K1_NO_LOOKBACK=1 timeout 200 python tools/index_probe.py --reps 2 --modes pinned32 > gpurun_out/k1_lookback_off_r2i.log 2>&1; grep -o "K1 [0-9.]* ms" gpurun_out/k1_lookback_off_r2i.log
# 2. the whole GPU suite
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_r2i.log 2>&1; tail -3 gpurun_out/pytest_r2i.log; grep -E "stress" gpurun_out/pytest_r2i.log | cut -c1-330
# 3. the default bench line, every leg on
( time timeout 1200 python bench.py > gpurun_out/bench_r2i.log 2> gpurun_out/bench_r2i.err ) 2>&1 | grep real; tail -4 gpurun_out/bench_r2i.err; cut -c1-400 gpurun_out/bench_r2i.log
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_r2i.log 2> gpurun_out/bench_ref_r2i.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_ref_r2i.log
# 4. whole-epoch captures (the launches the bench times): traffic.json
cap r2i_y10m_epoch sgd_kernel 1 2 833491505 python tools/ncu_target.py --workload y10m --slices 1 --launches 2
cap r2i_l10m_slice10 sgd_kernel 1 2 833491505 python tools/ncu_target.py --workload l10m --slices 10 --launches 2
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --e2e-epochs 0 --also 0 > gpurun_out/plain_launches_r2i.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2i.csv python bench.py --steps 6 --warmup 3 --no-cpu --e2e-epochs 0 --also 0 > gpurun_out/ncu_launches_r2i.log 2>&1
du -sh gpurun_out
