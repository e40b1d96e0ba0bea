python tools/profile_case.py grad 5000 > gpurun_out/plain_grad.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grad_ -c 5 -o gpurun_out/prof_grad3 python tools/profile_case.py grad 5000 > gpurun_out/ncu_grad3.log 2>&1
tail -n 2 gpurun_out/plain_grad.log
