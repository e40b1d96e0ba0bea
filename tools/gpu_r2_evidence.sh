# Round-2 evidence run (1 GPU): parity tests, smoke, bench (both arms), kernel timings, ncu launch lists and full captures.
# Everything profiles/r2_* was made from (tools/summarize_ncu.py turns the captures into the .md summaries).  Reports of the
# multifrontal kernels exceed the 64 MiB that travel back from the box: their raw page is exported to CSV there.
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/gputests.txt 2>&1 ) 2>&1 | tail -4
tail -15 gpurun_out/gputests.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2.err
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2.json 2>> gpurun_out/bench_r2.err; tail -c 800 gpurun_out/bench_r2.err
timeout 900 python tools/time_pdps.py 240 2>&1 | tee gpurun_out/time_pdps_f64.txt
BPLTV_PREC=32 timeout 900 python tools/time_pdps.py 240 2>&1 | tee gpurun_out/time_pdps_f32.txt
timeout 600 python tools/time_resident.py 2>&1 | tee gpurun_out/time_resident.txt
timeout 900 python tools/time_grad_nd.py 2>&1 | tee gpurun_out/time_grad_nd.txt
# launch list of the bench command (after it ran clean without ncu)
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
# full captures: the headline kernel, kernel B, the front factorisation (1 and 148 images)
timeout 300 python tools/profile_case.py tblock 16 > gpurun_out/plain_tblock.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdps_tblock -s 1 -c 2 -f -o gpurun_out/prof_tblock_t4 python tools/profile_case.py tblock 16 > gpurun_out/ncu_tblock.log 2>&1
timeout 300 python tools/profile_case.py resident 300 > gpurun_out/plain_resident.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdps_resident -c 1 -f -o gpurun_out/prof_resident python tools/profile_case.py resident 300 > gpurun_out/ncu_resident.log 2>&1
timeout 300 python tools/prof_grad_nd.py 128 1 5000 > gpurun_out/plain_nd.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd --csv --log-file gpurun_out/nd_launches_1x128.csv python tools/prof_grad_nd.py 128 1 5000 > gpurun_out/ncu_nd.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd --csv --log-file gpurun_out/nd_launches_148x128.csv python tools/prof_grad_nd.py 128 148 1000 >> gpurun_out/ncu_nd.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:nd_factor -c 24 -f -o /tmp/prof_nd_factor python tools/prof_grad_nd.py 128 148 1000 >> gpurun_out/ncu_nd.log 2>&1
ncu -i /tmp/prof_nd_factor.ncu-rep --page raw --csv > gpurun_out/prof_nd_factor_148x128_raw.csv 2>> gpurun_out/ncu_nd.log
timeout 900 ncu --set full --clock-control none -k regex:nd_factor -c 24 -f -o /tmp/prof_nd_factor1 python tools/prof_grad_nd.py 128 1 5000 >> gpurun_out/ncu_nd.log 2>&1
ncu -i /tmp/prof_nd_factor1.ncu-rep --page raw --csv > gpurun_out/prof_nd_factor_1x128_raw.csv 2>> gpurun_out/ncu_nd.log
du -sh gpurun_out
# sum-of-regularisers gradients on the nested-dissection solver (late round 2): timings, launch list, full capture of the
# cluster-shared front factorisation, solve share, learn runs, the 256x256 case
timeout 300 python tools/time_sumregs_grad.py 2>&1 | tee gpurun_out/time_sumregs_grad_nd.txt
timeout 300 python tools/time_nd_solve_share.py 2>&1 | tee gpurun_out/time_nd_solve_share.txt
timeout 600 python tools/time_sumregs_learn.py 2>&1 | tee gpurun_out/time_sumregs_learn.txt
timeout 300 python tools/time_sumregs_grad_256.py 2>&1 | tee gpurun_out/time_sumregs_grad_256.txt
timeout 200 python tools/prof_sumregs_grad_nd.py > gpurun_out/plain_nd3m.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd --csv --log-file gpurun_out/nd3m_launches_1x128.csv python tools/prof_sumregs_grad_nd.py > gpurun_out/ncu_nd3m.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nd_factor_cluster -s 12 -c 2 -f -o gpurun_out/prof_nd3m_cluster python tools/prof_sumregs_grad_nd.py > gpurun_out/ncu_nd3m_full.log 2>&1
du -sh gpurun_out
