python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import time, numpy as np, bpldenoising_b200 as bp
data=bp.synthetic_dataset(128,128,10,seed=7)
with bp.Context([0],64) as c:
    for arith in (bp.STRICT, bp.FAST):
        for k in range(2):
            u=c.denoise(data[1],0.1,bp.pdps_opts(arith=arith)); st=c.stats()
        print('resident 10x128x128 5000 its arith',arith,'ms',st['ms_pdps'],'kernel',st['pdps_kernel_used'])
    # many small images: resident (waves of clusters) vs march
    big=bp.synthetic_dataset(128,128,512,seed=9)[1]
    for kid in (bp.KERNEL_RESIDENT, bp.KERNEL_MARCH):
        for k in range(2):
            u=c.denoise(big,0.1,bp.pdps_opts(maxiter=1000,kernel=kid)); st=c.stats()
        print('512x128x128 1000 its kernel',kid,'ms',st['ms_pdps'],'Gpix-it/s',512*16384*1000/st['ms_pdps']/1e6)
PY
