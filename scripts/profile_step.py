"""GPU helper: torch.profiler breakdown of one cfg-2 training step (which kernels own the step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_fluid_phi_b200 as pb
from bench import turbulence_batch, CFG
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
dev = torch.device("cuda")
torch.manual_seed(1234)
model = pb.RationalFourierOperator3D(modes=CFG["modes"], width=CFG["width"], n_layers=CFG["n_layers"]).to(dev)
model.set_precision(prec)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
x = turbulence_batch(8, 128, 42, dev); y = turbulence_batch(8, 128, 43, dev)
def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(model(x), y)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
