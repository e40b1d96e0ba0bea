# Round-1 evidence run (1 GPU): parity tests, bench, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2>> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.err
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/profile_case.py tblock 12 > gpurun_out/plain_tblock.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pdps_tblock -s 2 -c 3 -f -o gpurun_out/prof_tblock python tools/profile_case.py tblock 12 > gpurun_out/ncu_tblock.log 2>&1
ls -la gpurun_out | head -40
