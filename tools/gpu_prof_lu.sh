set -x
mkdir -p gpurun_out
python tools/prof_lu.py > gpurun_out/prof_lu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'lu_factor|lu3_solve' -c 2 -f -o gpurun_out/prof_lu2 python tools/prof_lu.py > gpurun_out/prof_lu_ncu.log 2>&1
tail -n 3 gpurun_out/prof_lu_plain.log gpurun_out/prof_lu_ncu.log
python tools/time_lu_cluster.py 2>&1 | grep -v "cluster [128]\b\|cluster 16" | tail -12
