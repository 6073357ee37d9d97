#!/bin/bash
# ncu evidence of session 4 (run under gpurun): launch lists of the eager training step and of the eval pass, then a full
# capture of the inference convolution (osconv_tc_kernel<true>) at B=1024.
tag=${1:-r1s4}
python tools/eval_once.py --passes 1 > /dev/null 2>&1      # page the image in
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_step.csv \
    python tools/step_once.py --steps 3 > gpurun_out/${tag}_ncu_step.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_eval.csv \
    python tools/eval_once.py --B 1024 --passes 2 > gpurun_out/${tag}_ncu_eval.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"osconv_tc_kernel" -s 8 -c 7 -o gpurun_out/${tag}_infer_conv_B1024 -f \
    python tools/eval_once.py --B 1024 --passes 2 > gpurun_out/${tag}_ncu_infer.log 2>&1
tail -2 gpurun_out/${tag}_ncu_step.log gpurun_out/${tag}_ncu_eval.log gpurun_out/${tag}_ncu_infer.log
ls -la gpurun_out/${tag}_*
