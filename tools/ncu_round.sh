#!/bin/bash
# One gpurun call: the round's ncu evidence.  Every profiled command first runs plain and must exit 0.
#   launches.csv          every launch of a short bench run with its device time
#   conv_full.ncu-rep     ncu --set full (with source) of the three conv kernels of one forward, 64 clips
#   all_kernels.ncu-rep   ncu --set full of one launch of every kernel of the path, 16 clips
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --clips-per-gpu 256 --cpu-clips 0 --no-e2e --no-configs"
$B > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
export PROF_CLIPS=64
python tools/prof_conv.py > gpurun_out/plain_conv.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:conv_umma|conv_l2_fused' -s 3 -c 3 -f -o gpurun_out/conv_full python tools/prof_conv.py > gpurun_out/ncu_conv2.log 2>&1
echo "conv1/conv2/conv3 full rc=$?"
export PROF_CLIPS=16
python tools/prof_sweep.py > gpurun_out/plain_sweep.log 2>&1
rc=$?
N=$(grep -o "profiled pass: [0-9]*" gpurun_out/plain_sweep.log | grep -o "[0-9]*")
if [ $rc -eq 0 ] && [ -n "$N" ]; then
  ncu --set full --clock-control none -k 'regex:conv_umma|conv_l2_fused|mfcc_|pack_|vstats|sgemm|splitk|sweep_score|gemm_umma|gru_|log_softmax|ctc_|transpose_whh' -s $N -c 40 -f -o gpurun_out/all_kernels python tools/prof_sweep.py > gpurun_out/ncu_all.log 2>&1
  echo "all kernels rc=$? (skipped $N)"
fi
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
