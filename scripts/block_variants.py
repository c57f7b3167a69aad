"""GPU helper: CUDA-event time and achieved bandwidth of every variant of the block kernels at the cfg-2 shape
(middle / first (lift) / last (projection) block, forward fused with the z synthesis and backward), one launch each.

    [PDEPHI_LIB=...] python scripts/block_variants.py [iters]
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_fluid_phi_b200 import functional as pf, _cabi

torch.manual_seed(0)
dev = "cuda"
B, C, n = 8, 64, 128
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
mh = (32, 32, 17)
act = 4.0 * B * C * n ** 3
small = 4.0 * B * 3 * n ** 3
V = 4.0 * B * n * n * C * 36
# the `u` buffer: packed binary16 derivative (half an activation) under PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE = 2 (the default)
ub = act * (0.5 if _cabi.load().pdephi_get_option(_cabi.OPT_BLOCK_SAVES_DERIVATIVE) == 2 else 1.0)

h = torch.randn(B, C, n, n, n, device=dev)
x = torch.randn(B, 3, n, n, n, device=dev)
gout = torch.randn(B, 3, n, n, n, device=dev)
g = torch.randn_like(h)
Y = torch.randn(B, C, *mh, device=dev, dtype=torch.complex64) * 10
ms = torch.rand(mh, device=dev).reshape(-1) + 0.5
rs = torch.rand(B * C, 4, device=dev) + 0.5
W = (torch.randn(C, C, device=dev) * 0.1).contiguous()
b = torch.randn(C, device=dev)
Win = torch.randn(C, 3, device=dev).contiguous()
bin_ = torch.randn(C, device=dev)
Wp = (torch.randn(3, C, device=dev) * 0.1).contiguous()
bp = torch.randn(3, device=dev)
u, _ = pf.block_spectral_fwd(Y, ms, rs[:, 2], h, W, b)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

cases = {
    # name: (callable, algorithmic bytes of the block kernel itself)
    "fwd middle <fused>": (lambda: pf.block_spectral_fwd(Y, ms, rs[:, 2], h, W, b), 2 * act + ub + V),
    "fwd first  <fused,lift>": (lambda: pf.block_spectral_fwd(Y, ms, rs[:, 2], x, W, b, lift=(Win, bin_)), act + ub + V + small),
    "fwd last   <fused,proj>": (lambda: pf.block_spectral_fwd(Y, ms, rs[:, 2], h, W, b, head=(Wp, bp)), 2 * act + ub + V + small),
    # fused forward that also runs the z stage of the next layer's analysis on its output (vz out)
    "fwdz middle <fused>": (lambda: pf.block_spectral_fwd(Y, ms, rs[:, 2], h, W, b, want_vz=True), 2 * act + ub + 2 * V),
    "fwdz first  <fused,lift>": (lambda: pf.block_spectral_fwd(Y, ms, rs[:, 2], x, W, b, lift=(Win, bin_), want_vz=True), act + ub + 2 * V + small),
    "bwd middle": (lambda: pf.block_tail_bwd(g, u, h, W, True), 4 * act + ub),
    "bwd first  <lift, no t>": (lambda: pf.block_tail_bwd(g, u, x, W, True, write_t=False, lift=(Win, bin_)), 2 * act + ub + small),
    "bwd last   <proj>": (lambda: pf.block_tail_bwd(gout, u, h, W, True, head_w=Wp), 3 * act + ub + small),
    # fused backward (z stage of the adjoint analysis inside the kernel): gu is not stored, the z-analysed gu (Vg) is
    "bwdz middle": (lambda: pf.block_spectral_bwd(g, u, h, W, mh, True), 3 * act + ub + V),
    "bwdz first  <lift, no t>": (lambda: pf.block_spectral_bwd(g, u, x, W, mh, True, write_t=False, lift=(Win, bin_)), act + ub + small + V),
    "bwdz last   <proj>": (lambda: pf.block_spectral_bwd(gout, u, h, W, mh, True, head_w=Wp), 2 * act + ub + small + V),
}
res = {}
for name, (fn, nbytes) in cases.items():
    kname = ("block_spectral_fwd_kernel" if name.startswith("fwd") else
             "block_spectral_bwd_kernel" if name.startswith("bwdz") else "block_tail_bwd_kernel")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    _cabi.profile(enable=True, reset=True)
    for _ in range(iters):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    rep = _cabi.profile_report()
    _cabi.profile(enable=False, reset=True)
    ms_launch = rep[kname][1] / iters        # the bwd entry also counts the tiny reduce kernel's time
    res[name] = {"ms": round(ms_launch, 3), "GB": round(nbytes / 1e9, 2), "GB/s": round(nbytes / ms_launch / 1e6, 0)}
    print(f"{name:28s} {ms_launch:7.3f} ms  {nbytes / 1e9:6.2f} GB  {nbytes / ms_launch / 1e6:7.0f} GB/s", flush=True)
print(json.dumps({"lib": os.environ.get("PDEPHI_LIB", "default"), "variants": res}))
