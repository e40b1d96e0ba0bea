set -x
mkdir -p gpurun_out
python tools/prof_lu.py > gpurun_out/prof_lu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lu -c 4 -f -o gpurun_out/prof_lu python tools/prof_lu.py > gpurun_out/prof_lu_ncu.log 2>&1
tail -3 gpurun_out/prof_lu_plain.log gpurun_out/prof_lu_ncu.log
