import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artdeco_b200.mast3r import ops
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device=dev).manual_seed(0)
for K in (64, 256, 1024, 4096, 16384):
    a = torch.randn(256, K, device=dev, generator=g); w = torch.randn(256, K, device=dev, generator=g)
    ref = a.double() @ w.double().T
    out = torch.empty(256, 256, device=dev)
    ops.gemm(ops.split(a), ops.split(w), 256, 256, K, out=out)
    e3 = float((out.double() - ref).abs().max() / ref.abs().max())
    # exact-representable operands (bf16 values): isolates the accumulator
    ab, wb = a.bfloat16().float(), w.bfloat16().float()
    refb = ab.double() @ wb.double().T
    ops.gemm(ops.split(ab), ops.split(wb), 256, 256, K, out=out)
    eb = float((out.double() - refb).abs().max() / refb.abs().max())
    ef = float(((a @ w.T).double() - ref).abs().max() / ref.abs().max())
    print(f"K={K:6d}  bf16x3 err {e3:.2e}   exact-bf16-operands err {eb:.2e}   torch fp32 err {ef:.2e}")
