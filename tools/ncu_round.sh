#!/bin/bash
# ncu evidence of one round (run under gpurun): launch list of the eager step, then full captures of the hot kernels.
tag=${1:-r1s3}
python tools/step_once.py --steps 1 > /dev/null 2>&1      # page the image in, JIT nothing
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_step.csv \
    python tools/step_once.py --steps 3 > gpurun_out/${tag}_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"osconv_tc_kernel|oswgrad_tc_kernel" -c 3 -o gpurun_out/${tag}_conv_B1024 -f \
    python tools/prof_kernels.py --layer 1 --B 1024 --iters 1 > gpurun_out/${tag}_ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rowstats|adain_fwd|adain_bwd|gram_fwd|gram_bwd" -c 5 -o gpurun_out/${tag}_style -f \
    python tools/prof_style.py --iters 1 > gpurun_out/${tag}_ncu_style.log 2>&1
tail -2 gpurun_out/${tag}_ncu_step.log gpurun_out/${tag}_ncu_conv.log gpurun_out/${tag}_ncu_style.log
python tools/prof_style.py
ls -la gpurun_out/${tag}_*
