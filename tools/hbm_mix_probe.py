"""HBM bandwidth of a B200 as a function of the read : write mix (torch element-wise kernels over 2 GiB buffers, CUDA
events, best of 5).  The roofline denominator of bench.py is the COPY figure (1 : 1); the write-heavy kernels of the path
(qkv / fc1 GEMMs, the stem convolutions) sit on a lower ceiling, which this probe puts a number on."""
import json
import torch

dev = "cuda:0"
n = 1 << 29                                   # 2 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
c = torch.empty(n, device=dev)
d = torch.empty(n, device=dev)
a.normal_(); b.normal_(); c.normal_()


def best(fn, bytes_moved, reps=5):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return bytes_moved / (min(t) * 1e-3) / 1e9


res = {
    "write_only (fill)": best(lambda: d.fill_(1.0), 4 * n),
    "read_only (sum)": best(lambda: a.sum(), 4 * n),
    "copy 1r:1w": best(lambda: d.copy_(a), 8 * n),
    "add 2r:1w": best(lambda: torch.add(a, b, out=d), 12 * n),
    "addcmul 3r:1w": best(lambda: torch.addcmul(a, b, c, out=d), 16 * n),
}
# 1 read : 3 writes -- the byte mix of the stage-0 qkv GEMM: one fp32 read, three fp32 writes (three fills + a copy fused would
# need a custom kernel; approximate with a bf16 read -> fp32 + bf16 writes: 2 B read, 6 B written per element)
h = torch.empty(n, device=dev, dtype=torch.bfloat16)
h.normal_()
o1 = torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def one_to_three():
    o1.copy_(h)                                # 2 B read, 4 B written


res["bf16 -> fp32 cast 1r:2w"] = best(one_to_three, 6 * n)
for k, v in res.items():
    print(f"{k:28s} {v:8.0f} GB/s")
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
