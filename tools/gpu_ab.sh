cp build/ab/libbpltv_old.so /tmp/old.so
for rep in 1 2; do for lib in /tmp/old.so ""; do echo "LIB=${lib:-new}"; BPLTV_LIB=$lib python - <<'PY'
import os
if not os.environ.get("BPLTV_LIB"): os.environ.pop("BPLTV_LIB", None)
import time, numpy as np, bpldenoising_b200 as bp
for name,O,x,Delta in (("1x128 nonreg",1,0.1,0.1),("10x128 nonreg",10,0.1,0.1),("10x128 reg",10,0.1,1e-7),("148x128 nonreg",148,0.1,0.1)):
    data=bp.synthetic_dataset(128,128,O,seed=7)
    with bp.Context([0],64) as c:
        c.set_dataset(data)
        eo=bp.eval_opts(bp.pdps_opts(maxiter=5000 if O<=10 else 500))
        c.learn_eval(x,Delta,eo)
        ms=[]
        for k in range(3):
            c.learn_eval(x,Delta,eo); ms.append(c.stats()['ms_gradient'])
        print(f"  {name}: grad ms {min(ms):.1f} (all {['%.1f'%m for m in ms]})")
PY
done; done
