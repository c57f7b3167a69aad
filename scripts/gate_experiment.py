"""Where does the 'tf32' route's error on the headline model come from?  One oracle run (CPU, ~1 min), then the CUDA model of
tests/test_headline_gate_gpu.py under every combination of the route's split-TF32 switches (functional.TF32_SPLIT_WHERE and
the library's x-stage option).  Prints relative L2 of the output and of each parameter gradient against the oracle.

    python scripts/gate_experiment.py > gpurun_out/gate_experiment.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import test_headline_gate_gpu as T  # noqa: E402
from pde_fluid_phi_b200 import _cabi  # noqa: E402
from pde_fluid_phi_b200 import functional as pf  # noqa: E402


def main():
    _, _, _, out_o, grads_o, share = T._oracle()
    lib = _cabi.load()
    variants = [
        ("r1: all single-pass", 0, dict(lift_analysis=False, adjoint_synthesis=False, backward_analysis=False, forward_analysis=False)),
        ("+ x stage split", 1, dict(lift_analysis=False, adjoint_synthesis=False, backward_analysis=False, forward_analysis=False)),
        ("+ lift analysis split", 1, dict(lift_analysis=True, adjoint_synthesis=False, backward_analysis=False, forward_analysis=False)),
        ("+ adjoint synthesis split (default)", 1, dict(lift_analysis=True, adjoint_synthesis=True, backward_analysis=False, forward_analysis=False)),
        ("+ backward analysis split", 1, dict(lift_analysis=True, adjoint_synthesis=True, backward_analysis=True, forward_analysis=False)),
        ("+ forward analysis split", 1, dict(lift_analysis=True, adjoint_synthesis=True, backward_analysis=True, forward_analysis=True)),
        ("forward analysis split only (no backward)", 1, dict(lift_analysis=True, adjoint_synthesis=True, backward_analysis=False, forward_analysis=True)),
    ]
    res = {"spectral_share": share, "variants": {}}
    for name, xsplit, where in variants:
        lib.pdephi_set_option(_cabi.OPT_TF32_XSTAGE_SPLIT, xsplit)
        pf.TF32_SPLIT_WHERE.update(where)
        model, out, rep = T._run_cuda("tf32", True)
        errs = {"out": T.rel(out, out_o)}
        for n, p in model.named_parameters():
            errs[n] = T.rel(p.grad, grads_o[n])
        res["variants"][name] = {"max": max(errs.values()), "errs": {k: float(f"{v:.3e}") for k, v in errs.items()},
                                 "kernels": {k: v[0] for k, v in rep.items()}}
        print(name, "max %.2e" % max(errs.values()), {k: "%.2e" % v for k, v in errs.items() if "coeffs" in k or k == "out"},
              file=sys.stderr, flush=True)
        del model, out
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
