"""Host restatement of the trust-region driver (bpldenoising_b200/trbox.py ←
/root/reference/src/TRBox.jl): the scalar control flow and its quirks, on cheap synthetic
learning functions (no GPU), plus one end-to-end learn run on the GPU."""
import numpy as np
import pytest

from bpldenoising_b200 import trbox


def quad(x0):
    def lf(x, ds, Delta):
        x = np.asarray(x, dtype=np.float64)
        f = float(np.sum((x - x0) ** 2))
        g = 2 * (x - x0)
        return np.zeros((2, 2, 1)), f, (float(g) if g.ndim == 0 else g)
    return lf


def test_bounds_and_steps():
    lb, ub = trbox.get_bounds(0.05, 0.1)                      # TRBox.jl:160-164
    assert lb == max(-0.1, trbox.EPS - 0.05) and ub == 0.1
    assert trbox.in_bounds(lb, 0.1, 0.1) and not trbox.in_bounds(lb, 0.1, -0.06)
    # QUIRK (:149-152): element-wise max of both ratios
    assert trbox.step_to_bound(-1.0, -0.05, 0.1) == 0.05 and trbox.step_to_bound(1.0, -0.05, 0.1) == 0.1
    # scalar dogleg with B = 0.1: gx > 0 → Cauchy step leaves the box → scaled to the bound,
    # limited by positivity x + p ≥ eps
    assert trbox.dogleg_box(0.5, 300.0, 0.1, 0.1) == pytest.approx(-0.1)
    assert trbox.dogleg_box(0.05, 300.0, 0.1, 0.1) == pytest.approx(-(0.05 - trbox.EPS))
    assert trbox.dogleg_box(0.5, -300.0, 0.1, 0.1) == pytest.approx(0.1)
    # QUIRK (:63): tiny gradient → the "Newton" step +gx/B is taken as is (uphill)
    assert trbox.dogleg_box(0.5, 1e-3, 0.1, 0.1) == pytest.approx(1e-2)
    assert trbox.pred(0.1, -0.1, 300.0) == pytest.approx(30.0 - 0.5 * 0.1 * 0.01)
    # QUIRK (:181-186): the scalar BFGS update is lost
    assert trbox.update_bfgs(0.1, 5.0, 0.2) == 0.1


def test_scalar_learn_on_a_quadratic():
    res = trbox.bilevel_learn(None, quad(0.3), 0.1, dict(maxiter=40, tol=1e-9))
    assert res.evaluations == len(res.log) + 1
    assert abs(res.x - 0.3) < 0.02
    fs = [e.function_value for e in res.log]
    assert all(b <= a + 1e-15 for a, b in zip(fs, fs[1:]))    # only ρ > 0 steps are accepted
    # radius: grows ×1.9 on very successful full steps, shrinks ×0.25 on failures
    r = [0.1] + [e.radius for e in res.log]
    ratios = {round(b / a, 6) for a, b in zip(r, r[1:])}
    assert ratios <= {1.0, 1.9, 0.25, round(0.25 * 0.25, 6), round(1.9 * 0.25, 6)}
    # stop rule: a logged iteration saw Δ < tol, or maxiter was reached
    res2 = trbox.bilevel_learn(None, quad(0.3), 0.1, dict(maxiter=200, tol=1e-2))
    assert res2.log[-1].radius < 1e-2 and len(res2.log) < 200


def test_lbfgs_operator_is_spd_and_secant():
    rng = np.random.default_rng(0)
    B = trbox.LBFGSOperator(4)
    A = np.diag([1.0, 2.0, 5.0, 10.0])
    for _ in range(7):
        s = rng.standard_normal(4)
        B.push(s, A @ s)
    D = B.dense()
    assert np.allclose(D, D.T, atol=1e-10) and np.all(np.linalg.eigvalsh(D) > 0)
    assert np.allclose(B.mul(B.s[-1]), B.y[-1], rtol=1e-10)   # secant equation for the newest pair
    assert len(B.s) == 5


def test_patch_learn_on_a_quadratic():
    x0 = np.array([[2e-4, 5e-5], [1.5e-4, 3e-4]])
    res = trbox.bilevel_learn(None, quad(x0), 1e-4 * np.ones((2, 2)), dict(maxiter=60, tol=1e-12))
    assert res.x.shape == (2, 2) and np.all(res.x > 0)
    f0 = float(np.sum((1e-4 - x0) ** 2))
    assert res.log[-1].function_value < 0.5 * f0


@pytest.mark.gpu
def test_learn_run_end_to_end_on_gpu(bp, ctx, datasets):
    # BASELINE config 1 end to end: 1 + ≤20 evaluations of the CUDA learning function
    data = datasets["cameraman_128_5"]
    res = trbox.scalar_bilevel_tv_learn(data, ctx=ctx)
    assert 2 <= res.evaluations <= 21 and res.u.shape == (128, 128, 1)
    f0 = bp.tv_op_learning_function(0.1, data, 0.1, ctx=ctx)[1]
    assert res.log[-1].function_value < f0            # the learned λ beats λ₀ = 0.1
    assert 0 < res.x < 0.1
    resp = trbox.patch_bilevel_tv_learn(datasets["circle_128_10"], ctx=ctx, maxiter=4)
    assert resp.x.shape == (2, 2) and resp.evaluations == 5


@pytest.mark.gpu
def test_run_experiment_cli_reproduces_the_artefacts(tmp_path, datasets):
    """tools/run_experiment.py: dataset directory in the reference's format → learn run → log, quality
    table and PNGs under <output>/<dataset_name>/ (BPLDenoising.jl:325-344, :185-217)."""
    import os
    import subprocess
    import sys
    from PIL import Image
    from conftest import ROOT
    t, d = datasets["cameraman_128_5"]
    root = tmp_path / "BPLDenoising" / "datasets" / "cameraman_128_5"
    root.mkdir(parents=True)
    Image.fromarray(np.round(t[:, :, 0] * 255).astype(np.uint8)).save(root / "cameraman_128_5_true_1.png")
    Image.fromarray(np.round(d[:, :, 0] * 255).astype(np.uint8)).save(root / "cameraman_128_5_data_1.png")
    (root / "filelist.txt").write_text("cameraman_128_5_true_1.png,cameraman_128_5_data_1.png\n")
    out = tmp_path / "output"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_experiment.py"), "scalar_bilevel_tv_learn",
                        "--dataset_name", "cameraman", "--datasets_dir", str(tmp_path / "BPLDenoising" / "datasets"),
                        "--maxiter", "4", "--output", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    files = sorted(os.listdir(out / "cameraman_128_5"))
    stem = "tv_optimal_parameter_scalar_cameraman_128_5"
    assert f"{stem}.txt" in files and f"{stem}_quality.txt" in files and f"{stem}_reco_1.png" in files
    q = open(out / "cameraman_128_5" / f"{stem}_quality.txt").read().splitlines()
    assert len(q) == 3 and float(q[1].split()[4]) > float(q[1].split()[2])    # out_psnr > orig_psnr


def test_cg_solve_is_the_inexact_newton_step_of_cg_lanczos():
    """newton_step (TRBox.jl:135-141) solves B pn = −g with Krylov.cg_lanczos: CG iterates, stopped at
    ‖r‖ ≤ √eps + √eps·‖b‖.  The restatement meets that bound, agrees with the exact solve to the same level and stops
    within 2n products of an L-BFGS operator holding several pairs."""
    from bpldenoising_b200 import trbox
    rng = np.random.default_rng(4)
    n = 12
    B = trbox.LBFGSOperator(n)
    H = rng.standard_normal((n, n)); H = H @ H.T + n * np.eye(n)
    for _ in range(7):                                   # more pairs than the memory holds
        s = rng.standard_normal(n)
        B.push(s, H @ s)
    D = B.dense()
    assert np.allclose(D, D.T, rtol=1e-12, atol=1e-12) and np.all(np.linalg.eigvalsh(0.5 * (D + D.T)) > 0)
    calls = []
    g = rng.standard_normal(n)
    x = trbox.cg_solve(lambda v: (calls.append(1), B.mul(v))[1], -g)
    tol = np.sqrt(np.finfo(float).eps)
    assert np.linalg.norm(D @ x + g) <= 1.01 * (tol + tol * np.linalg.norm(g)) and len(calls) <= 2 * n
    exact = np.linalg.solve(D, -g)
    assert np.linalg.norm(x - exact) <= 1e-6 * np.linalg.norm(exact)
    assert np.array_equal(trbox.cg_solve(B.mul, np.zeros(n)), np.zeros(n))
