# Round-2 evidence run (1 GPU): parity tests, smoke, gradient timings (ND vs band), bench (both arms),
# ncu launch lists + full captures of the top PDPS kernel and of the nested-dissection factorisation.
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python tools/time_grad_nd.py 2>&1 | tee gpurun_out/time_grad_nd.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2.err
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2.json 2>> gpurun_out/bench_r2.err; tail -c 800 gpurun_out/bench_r2.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/profile_case.py tblock 12 > gpurun_out/plain_tblock.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdps_tblock -s 2 -c 2 -f -o gpurun_out/prof_tblock python tools/profile_case.py tblock 12 > gpurun_out/ncu_tblock.log 2>&1
timeout 300 python tools/prof_grad_nd.py 128 1 5000 > gpurun_out/plain_nd.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd --csv --log-file gpurun_out/nd_launches_1x128.csv python tools/prof_grad_nd.py 128 1 5000 > gpurun_out/ncu_nd.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd --csv --log-file gpurun_out/nd_launches_148x128.csv python tools/prof_grad_nd.py 128 148 1000 >> gpurun_out/ncu_nd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nd_factor -c 40 -f -o gpurun_out/prof_nd_factor_148x128 python tools/prof_grad_nd.py 128 148 1000 >> gpurun_out/ncu_nd.log 2>&1
tail -3 gpurun_out/plain_nd.log gpurun_out/ncu_nd.log
ls -la gpurun_out | head -40
