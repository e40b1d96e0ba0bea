# Round-2 run 2 (1 GPU): BallScale chain — selftest, parity tests with durations, PDPS timings; bench; ncu captures
# exported to CSV on the box (the .ncu-rep files of the multifrontal kernels exceed the 64 MiB that travel back).
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=40 > gpurun_out/gputests.txt 2>&1 ) 2>&1 | tail -4
tail -60 gpurun_out/gputests.txt
timeout 600 python tools/dbg_fp32_grad.py 2>&1 | tee gpurun_out/dbg_fp32_grad.txt
timeout 900 python tools/time_pdps.py 240 2>&1 | tee gpurun_out/time_pdps_f64.txt
BPLTV_PREC=32 timeout 900 python tools/time_pdps.py 240 2>&1 | tee gpurun_out/time_pdps_f32.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2.err
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2.json 2>> gpurun_out/bench_r2.err; tail -c 800 gpurun_out/bench_r2.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/profile_case.py tblock 12 > gpurun_out/plain_tblock.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdps_tblock -s 2 -c 2 -f -o gpurun_out/prof_tblock python tools/profile_case.py tblock 12 > gpurun_out/ncu_tblock.log 2>&1
timeout 300 python tools/prof_grad_nd.py 128 1 5000 > gpurun_out/plain_nd.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd --csv --log-file gpurun_out/nd_launches_1x128.csv python tools/prof_grad_nd.py 128 1 5000 > gpurun_out/ncu_nd.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd --csv --log-file gpurun_out/nd_launches_148x128.csv python tools/prof_grad_nd.py 128 148 1000 >> gpurun_out/ncu_nd.log 2>&1
# full capture of the factor kernels of ONE gradient (the last, largest levels dominate): raw page as CSV, report dropped
timeout 900 ncu --set full --clock-control none -k regex:nd_factor -c 24 -f -o /tmp/prof_nd_factor python tools/prof_grad_nd.py 128 148 1000 >> gpurun_out/ncu_nd.log 2>&1
ncu -i /tmp/prof_nd_factor.ncu-rep --page raw --csv > gpurun_out/prof_nd_factor_148x128_raw.csv 2>> gpurun_out/ncu_nd.log
timeout 900 ncu --set full --clock-control none -k regex:nd_factor -c 24 -f -o /tmp/prof_nd_factor1 python tools/prof_grad_nd.py 128 1 5000 >> gpurun_out/ncu_nd.log 2>&1
ncu -i /tmp/prof_nd_factor1.ncu-rep --page raw --csv > gpurun_out/prof_nd_factor_1x128_raw.csv 2>> gpurun_out/ncu_nd.log
tail -n 3 gpurun_out/plain_nd.log gpurun_out/ncu_nd.log
du -sh gpurun_out; ls -la gpurun_out | head -40
