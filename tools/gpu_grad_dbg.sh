for d in 0 1 2 4 3 7; do echo "DBG=$d"; BPLTV_GRAD_DBG=$d python - <<'PY'
import time, numpy as np, bpldenoising_b200 as bp
data=bp.synthetic_dataset(128,128,1,seed=7)
with bp.Context([0],64) as c:
    c.set_dataset(data)
    u=c.denoise(None,0.1)
    for rep in range(2):
        try:
            g=c.gradient(0.1,u,regularised=False); st=c.stats(); print('  grad ms',st['ms_gradient'])
        except Exception as e: print('  err',str(e)[:60], c.stats()['ms_gradient'])
PY
done
