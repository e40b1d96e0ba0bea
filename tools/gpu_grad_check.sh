python -m pytest tests/test_gpu_gradient.py -x -q 2>&1 | tail -8
python - <<'PY'
import time, numpy as np, bpldenoising_b200 as bp
for name,O,x,Delta,fb in (("1x128 nonreg",1,0.1,0.1,0),("10x128 nonreg",10,0.1,0.1,0),("10x128 reg",10,0.1,1e-7,0),("1x128 patch",1,0.01*np.ones((2,2)),1e-4,0),("148x128 nonreg",148,0.1,0.1,0)):
    data=bp.synthetic_dataset(128,128,O,seed=7)
    with bp.Context([0],64) as c:
        c.set_dataset(data)
        eo=bp.eval_opts(bp.pdps_opts(maxiter=5000 if O<=10 else 500))
        c.learn_eval(x,Delta,eo)
        t0=time.perf_counter(); u,cost,g=c.learn_eval(x,Delta,eo); dt=(time.perf_counter()-t0)*1e3
        st=c.stats()
        print(f"{name}: total {dt:.1f} ms pdps {st['ms_pdps']:.1f} grad {st['ms_gradient']:.1f} kernel {st['pdps_kernel_used']} cost {cost:.6f} grad {np.ravel(g)[:2]}")
PY
