#!/usr/bin/env python
"""Julia-free reproduction of the reference's experiment entry points:

    python tools/run_experiment.py scalar_bilevel_tv_learn   --dataset_name cameraman_128_5 --datasets_dir BPLDenoising/datasets
    python tools/run_experiment.py patch_bilevel_tv_learn    --dataset_name circle_128_10
    python tools/run_experiment.py scalar_bilevel_sumregs_learn --dataset_name cameraman_128_5
    python tools/run_experiment.py patch_bilevel_sumregs_learn  --dataset_name cameraman_128_5
    python tools/run_experiment.py validate_tv_parameter --parameter 0.07 --dataset_name faces_val_128_10
    python tools/run_experiment.py generate_scalar_tv_cost --dataset_name cameraman_128_5 --range 1e-3 1.0 64
    python tools/run_experiment.py generate_2d_tv_cost     --dataset_name circle_128_10   --range 1e-3 1.0 16

Each follows /root/reference/src/BPLDenoising.jl (:325-344, :359-376, :432-451, :464-481, :381-415): load the
dataset (filelist.txt + PNG pairs, Datasets.jl:54-65), take `num_samples` images, run the trust-region
driver with the library-backed learning function, stretch and save the artefacts (`save_results`,
:185-258) under <output>/<dataset_name>/.  Everything numerical happens in libbpltv on the GPU.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
from bpldenoising_b200 import results, trbox  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("experiment", choices=["scalar_bilevel_tv_learn", "patch_bilevel_tv_learn",
                                           "scalar_bilevel_sumregs_learn", "patch_bilevel_sumregs_learn",
                                           "validate_tv_parameter", "generate_scalar_tv_cost", "generate_2d_tv_cost"])
    ap.add_argument("--range", nargs=3, type=float, default=[1e-3, 1.0, 64], metavar=("FIRST", "LAST", "COUNT"),
                    help="cost curves: geometric parameter range (the reference's callers pass their own range)")
    ap.add_argument("--dataset_name", default="cameraman_128_5")          # default_params (:306-314)
    ap.add_argument("--datasets_dir", default="BPLDenoising/datasets/")   # Datasets.jl:9
    ap.add_argument("--num_samples", type=int, default=1)
    ap.add_argument("--maxiter", type=int, default=20)
    ap.add_argument("--tol", type=float, default=1e-5)
    ap.add_argument("--parameter", type=float, default=0.1, help="validate_tv_parameter only")
    ap.add_argument("--output", default=results.default_save_prefix)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)

    b, b_noisy = bp.testdataset(a.dataset_name, dataset_dir=a.datasets_dir)
    name = bp.datasets.full_datasetname(a.dataset_name)
    with bp.Context([a.device], 64) as ctx:
        if a.experiment == "validate_tv_parameter":
            res = bp.validate_tv_parameter(a.parameter, (b, b_noisy), ctx=ctx)
            prm = dict(dataset_name=name, save_prefix=f"val_tv_optimal_parameter_scalar_()_{name}", save_results=True)
            w = results.save_results(prm, b, b_noisy, a.parameter, res["u"], [], out_root=a.output)
            print(f"cost = {res['cost']:.6f}, mean SSIM {w['mean_ssim']:.4f}, mean PSNR {w['mean_psnr']:.3f} dB → {w['quality']}")
            return 0
        k = a.num_samples
        data = (np.asfortranarray(b[:, :, :k]), np.asfortranarray(b_noisy[:, :, :k]))
        if a.experiment in ("generate_scalar_tv_cost", "generate_2d_tv_cost"):
            # generate_cost / generate_2d_cost (:92-111, :136-158): the whole range as one batched launch, saved under the
            # reference's variable names (results.save_cost_curve)
            pr = np.geomspace(a.range[0], a.range[1], int(a.range[2]))
            if a.experiment == "generate_scalar_tv_cost":
                costs = bp.generate_scalar_tv_cost(data, pr, num_samples=k, ctx=ctx)
                w = results.save_cost_curve(name, pr, costs, out_root=a.output)
            else:
                costs = bp.generate_2d_tv_cost(data, pr, pr, num_samples=k, ctx=ctx)
                w = results.save_cost_curve(name, pr, costs, parameter_range_2=pr, out_root=a.output)
            print(f"{costs.size} costs, minimum {float(np.min(costs)):.6f} → {w['npz']}")
            return 0
        run = {"scalar_bilevel_tv_learn": (trbox.scalar_bilevel_tv_learn, "tv_optimal_parameter_scalar_"),
               "patch_bilevel_tv_learn": (trbox.patch_bilevel_tv_learn, "tv_optimal_parameter_(2, 2)_"),
               "scalar_bilevel_sumregs_learn": (trbox.scalar_bilevel_sumregs_learn, "sumregs_optimal_parameter_scalar_"),
               # "sumregs_optimal_parameter_patch_$(size(params.α₀))" * dataset_name (:467): no separator
               "patch_bilevel_sumregs_learn": (trbox.patch_bilevel_sumregs_learn, "sumregs_optimal_parameter_patch_(2, 2, 3)")}
        fn, prefix = run[a.experiment]
        res = fn(data, ctx=ctx, maxiter=a.maxiter, tol=a.tol)
        prm = dict(dataset_name=name, save_prefix=prefix + name, save_results=True, maxiter=a.maxiter, tol=a.tol,
                   num_samples=k)
        # adjust_histogram!(…, LinearStretching()) on u (and, for the TV experiments, on b and b_noisy) (:337-339)
        u = results.linear_stretch(res.u)
        if "sumregs" not in a.experiment:
            data = (results.linear_stretch(data[0]), results.linear_stretch(data[1]))
        w = results.save_results(prm, data[0], data[1], res.x, u, res.log, out_root=a.output)
        print(f"x = {np.asarray(res.x).tolist()}, {res.evaluations} evaluations in {res.seconds:.2f} s, "
              f"final cost {res.log[-1].function_value:.6f} → {w['log']}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
