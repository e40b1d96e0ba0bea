mkdir -p gpurun_out
for C in 8 2 1; do echo "== BPLTV_GRAD_CLUSTER=$C"; BPLTV_GRAD_CLUSTER=$C timeout 900 python -m pytest tests/test_gpu_gradient.py tests/test_gpu_sumregs.py -m gpu -x -q 2>&1 | tail -3; done
echo "== auto"; timeout 900 python -m pytest tests/test_gpu_gradient.py tests/test_gpu_sumregs.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.getcwd())
import bpldenoising_b200 as bp
z = np.load("tests/golden/datasets.npz")
def ds(name):
    return (np.asfortranarray(z[name+"/true"].astype(float)/z[name+"/true_div"]), np.asfortranarray(z[name+"/data"].astype(float)/z[name+"/data_div"]))
for C in ("1", "2", "4", "8", ""):
    if C: os.environ["BPLTV_GRAD_CLUSTER"] = C
    else: os.environ.pop("BPLTV_GRAD_CLUSTER", None)
    row = {}
    with bp.Context([0], 64) as c:
        for name, x, D in (("cameraman_128_5", 0.1, 0.1), ("faces_train_128_10", 0.1, 0.1), ("faces_train_128_10", 0.1, 1e-7)):
            c.set_dataset(ds(name)); c.learn_eval(x, D)
            _, cost, g = c.learn_eval(x, D); row[name + ("_reg" if D < 1e-6 else "")] = (round(c.stats()["ms_gradient"], 2), g)
        c.set_dataset(ds("cameraman_128_5")); x0 = np.array([0.001] * 3); c.sumregs_learn_eval(x0, 0.01)
        _, cost, g = c.sumregs_learn_eval(x0, 0.01); row["sumregs"] = (round(c.stats()["ms_gradient"], 1), g.tolist())
    print("cluster", C or "auto", row, flush=True)
PY
