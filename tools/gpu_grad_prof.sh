for d in 0 8 0 8; do
  export BPLTV_GRAD_DBG=$d
  python tools/profile_case.py grad 5000 > gpurun_out/plain_grad.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:grad_factor --csv --log-file gpurun_out/grad_times.csv python tools/profile_case.py grad 5000 > gpurun_out/ncu_grad4.log 2>&1
  echo "DBG=$d"; grep -E "grad_" gpurun_out/grad_times.csv | awk -F'","' '{print substr($5,1,40), $(NF)}' | head -3
done
