set -x
cd $GRAFT_REPO_ROOT
SW="GPUS=1;GPUS=2;GPUS=2,SYNCS=1;GPUS=2,SYNCS=40;GPUS=2,OVERLAP=2;GPUS=2,LAYOUT_F64=1;GPUS=1,LAYOUT_F64=1;GPUS=2,WINDOW=0;GPUS=1,WINDOW=0;GPUS=2,COHERENT=0;GPUS=1,COHERENT=0"
timeout 600 python tools/one_call_multi.py --nodes 200000 --paths 16 --dims 2 --reps 9 --sweep "$SW" > gpurun_out/one_call_r2r_l200k.log 2>&1
grep -E "summary|Error|error" gpurun_out/one_call_r2r_l200k.log | cut -c1-330
timeout 900 python tools/one_call_multi.py --nodes 1000000 --paths 32 --dims 2 --reps 5 --sweep "$SW" > gpurun_out/one_call_r2r_l1m.log 2>&1
grep -E "summary|Error|error" gpurun_out/one_call_r2r_l1m.log | cut -c1-330
