set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1.err
