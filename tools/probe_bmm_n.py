"""Batch-1 torch.matmul [3,3]x[3,N] and [3,4]x[4,N]: FMA chain or rounded products, as a function of N."""
import torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def f32(x): return x.to(torch.float32)
def chain(A, X, fma):
    acc = f32(A[:, 0, None].double() * X[None, 0].double())
    for j in range(1, A.shape[1]):
        if fma: acc = f32(A[:, j, None].double() * X[None, j].double() + acc.double())
        else: acc = f32(acc.double() + f32(A[:, j, None].double() * X[None, j].double()).double())
    return acc
for k in (3, 4):
    A = torch.randn(3, k, device=dev, generator=g)
    res = []
    for N in (2048, 6144, 24576, 65536, 122880, 131072, 131073, 147456, 163840, 196608, 229376, 262144, 262145, 294912, 327680, 491520):
        X = torch.randn(k, N, device=dev, generator=g)
        ref = torch.matmul(A[None], X[None])[0]
        res.append((N, round(float((chain(A, X, True) == ref).float().mean()), 3), round(float((chain(A, X, False) == ref).float().mean()), 3)))
    print("k=%d (N, fma, mulacc):" % k, res)
