set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python tools/one_call_multi.py --dims 2 --reps 5 --sweep "GPUS=2;GPUS=1" > gpurun_out/one_call_r2u_l10m.log 2>&1
grep -E "one call|summary|Error|error" gpurun_out/one_call_r2u_l10m.log | cut -c1-420
