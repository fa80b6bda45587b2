"""Which rounding sequence does torch.matmul use for the tiny-k products of warp.py on this GPU?"""
import itertools, sys, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def f32(x): return x.to(torch.float32)
def fma(a, b, c): return f32(a.double() * b.double() + c.double())
def mul(a, b): return f32(a.double() * b.double())
def cands(A, X):
    """A [m,k], X [k,N] float32 -> dict name -> [m,N] for several accumulation orders."""
    m, k = A.shape
    out = {}
    for name, order in (("asc", range(k)), ("desc", range(k - 1, -1, -1))):
        order = list(order)
        acc = mul(A[:, order[0], None], X[None, order[0]])
        for j in order[1:]:
            acc = fma(A[:, j, None], X[None, j], acc)
        out["fma_" + name] = acc
        acc = mul(A[:, order[0], None], X[None, order[0]])
        for j in order[1:]:
            acc = f32(acc.double() + mul(A[:, j, None], X[None, j]).double())
        out["mulacc_" + name] = acc
    if k == 4:
        lo = fma(A[:, 1, None], X[None, 1], mul(A[:, 0, None], X[None, 0]))
        hi = fma(A[:, 3, None], X[None, 3], mul(A[:, 2, None], X[None, 2]))
        out["pair"] = f32(lo.double() + hi.double())
        acc = fma(A[:, 2, None], X[None, 2], mul(A[:, 0, None], X[None, 0]))
        acc2 = fma(A[:, 3, None], X[None, 3], mul(A[:, 1, None], X[None, 1]))
        out["evenodd"] = f32(acc.double() + acc2.double())
    out["exact"] = f32((A.double() @ X.double()))
    return out
g = torch.Generator(device=dev).manual_seed(0)
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (192, 640)
import numpy as np
K = np.array([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
invK = torch.from_numpy(np.linalg.pinv(K)).to(dev)
ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev), torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=dev)], 0)
for B in (1, 2, 12):
    ref = torch.matmul(invK[None, :3, :3].repeat(B, 1, 1), pix[None].repeat(B, 1, 1))[0]
    c = cands(invK[:3, :3], pix)
    print("rays B=%d" % B, {k: round(float((v == ref).float().mean()), 4) for k, v in c.items()})
P = (torch.randn(3, 4, device=dev, generator=g))
cam = torch.cat([torch.randn(3, H * W, device=dev, generator=g) * 5, torch.ones(1, H * W, device=dev)], 0)
for B in (1, 2, 12):
    ref = torch.matmul(P[None].repeat(B, 1, 1), cam[None].repeat(B, 1, 1))[0]
    c = cands(P, cam)
    print("proj B=%d" % B, {k: round(float((v == ref).float().mean()), 4) for k, v in c.items()})
Kt = torch.from_numpy(K).to(dev); T = torch.eye(4, device=dev); T[:3, :] += 0.01 * torch.randn(3, 4, device=dev, generator=g)
for B in (1, 2, 12):
    ref = torch.matmul(Kt[None].repeat(B, 1, 1), T[None].repeat(B, 1, 1))[0]
    c = cands(Kt, T)
    print("K@T  B=%d" % B, {k: round(float((v == ref).float().mean()), 4) for k, v in c.items()})
