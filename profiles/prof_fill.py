"""Write-only ceiling for the env frame output: torch fill_ / zero_ (memset) of the same 252 MB buffer vs env_step_kernel."""
import torch
x = torch.empty(65536 * 960, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
nbytes = x.numel() * 4
bufs = [x, y]
i = [0]
def fill(): i[0] ^= 1; bufs[i[0]].fill_(1.0)
def zero(): i[0] ^= 1; bufs[i[0]].zero_()
def copy(): y.copy_(x)
for name, fn, moved in (("fill_", fill, nbytes), ("zero_ (memset)", zero, nbytes), ("copy_ (read+write)", copy, 2 * nbytes)):
    ms = timed(fn)
    print(f"{name:20s} {ms*1e3:8.1f} us  {moved/ms/1e6:8.1f} GB/s")
