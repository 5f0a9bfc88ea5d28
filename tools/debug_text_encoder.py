import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ops_ref as R
from t2v_b200 import prims as P

def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-9)).item()

torch.manual_seed(0)
for (B, L, C, H) in [(3, 16, 64, 1), (3, 24, 128, 2), (3, 77, 1024, 16), (3, 16, 1024, 16), (3, 77, 64, 1)]:
    D = C // H
    dev = "cuda"
    ids = torch.randint(0, 50, (B, L), device=dev)
    tok = torch.randn(50, C, device=dev); pos = torch.randn(L, C, device=dev)
    x = P.embed_tokens(ids, tok, pos); xr = R.embed_tokens(ids, tok, pos)
    print((B, L, C, H), "embed", rel(x, xr))
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    n, _ = P.layernorm_fwd(x, g, b, 1e-5); nr, _ = R.layernorm_fwd(x, g, b, 1e-5)
    print("  ln", rel(n, nr))
    w = (torch.randn(3 * C, C, device=dev) / C ** 0.5).bfloat16().view(3 * C, 1, 1, C); bias = torch.randn(3 * C, device=dev)
    qkv = P.conv_fwd(n.view(1, 1, B * L, C), w, bias).view(B, L, 3 * C)
    qkvr = R.conv_fwd(n.view(1, 1, B * L, C), w, bias).view(B, L, 3 * C)
    print("  qkv", rel(qkv, qkvr))
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    ld = (L + 7) // 8 * 8
    s = torch.zeros((B, H, L, ld), device=dev); sr = torch.zeros((B, H, L, ld), device=dev)
    args = (q, (1, q.stride(1), q.stride(0), D), k, (1, k.stride(1), k.stride(0), D))
    P.bgemm(*args, s, (ld, H * L * ld, L * ld), L, L, D, B, H, D ** -0.5, 1)
    R.bgemm(*args, sr, (ld, H * L * ld, L * ld), L, L, D, B, H, D ** -0.5, 1)
    print("  scores", rel(s[..., :L], sr[..., :L]))
    p = P.softmax_fwd(s, L, ld, causal_period=L); pr = R.softmax_fwd(s, L, ld, causal_period=L)
    print("  softmax", rel(p, pr))
    a = torch.zeros((B, L, C), device=dev, dtype=torch.bfloat16); ar = torch.zeros_like(a)
    args = (p, (1, ld, H * L * ld, L * ld), v, (0, v.stride(1), v.stride(0), D))
    P.bgemm(*args, a, (C, L * C, D), L, D, L, B, H, 1.0, 0)
    R.bgemm(*args, ar, (C, L * C, D), L, D, L, B, H, 1.0, 0)
    print("  pv", rel(a, ar))
    h = P.gelu_bf16(qkv.reshape(-1, 3 * C)); hr = R.gelu_bf16(qkv.reshape(-1, 3 * C))
    print("  gelu", rel(h, hr))
