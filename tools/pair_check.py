import sys, os
os.environ["PSW_DIAGNOSTICS"] = "1"
import torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panoswintransformerobjectdetection_b200 import ops, _lib
lib = _lib.load()
torch.manual_seed(0)
dev = 'cuda:0'
def check(M, N, K, gelu, res, od, force):
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16(); b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).to(od) if res else None
    lib.psw_diag_linear_mode((1 << 27) if force else (1 << 26))
    y = ops.linear(x, w, b, residual=r, gelu=gelu, out_dtype=od)
    lib.psw_diag_linear_mode(0)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    if gelu: ref = torch.nn.functional.gelu(ref)
    if res: ref = ref + r.float()
    err = ((y.float() - ref).norm() / ref.norm()).item()
    print(f"M{M} N{N} K{K} gelu{gelu} res{res} {od} force{force}: rel {err:.2e}", flush=True)
    assert err < 6e-3, err
for force in (1, 0):
    check(512, 256, 128, False, False, torch.bfloat16, force)
    check(1024, 512, 384, True, False, torch.bfloat16, force)
    check(1000, 384, 1536, False, True, torch.float32, force)
    check(65536, 1152, 384, False, False, torch.bfloat16, force)
    check(16384, 768, 3072, False, True, torch.float32, force)
    check(333, 96, 64, False, True, torch.bfloat16, force)
print("pair ok")
