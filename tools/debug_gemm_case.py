import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gemm_gpu as T
from t2v_b200 import native as nat
cases = T.CONV_CASES if len(sys.argv) < 2 else [T.CONV_CASES[int(sys.argv[1])]]
for case in cases:
    N, H, W, Ci, Co, KH, KW, s, pads = case
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Co, KH, KW, Ci, device="cuda", generator=g) / (KH * KW * Ci) ** 0.5).bfloat16()
    bias = torch.randn(Co, device="cuda", generator=g)
    ref = T.conv_ref(x, w, s, pads)
    Ho, Wo = ref.shape[1], ref.shape[2]
    rowbias = torch.randn(N, Co, device="cuda", generator=g)
    res = torch.randn(N, Ho, Wo, Co, device="cuda", generator=g).bfloat16()
    for name, args, dt, want in (("plain", (None, None, None, 1.0, 0, 1), torch.bfloat16, ref),
                                 ("bias", (bias.data_ptr(), None, None, 1.0, 0, 1), torch.bfloat16, ref + bias),
                                 ("bias+rb", (bias.data_ptr(), rowbias.data_ptr(), None, 1.0, 0, 1), torch.bfloat16, ref + bias + rowbias[:, None, None, :]),
                                 ("res", (None, None, res.data_ptr(), 1.0, 0, 1), torch.bfloat16, ref + res.float()),
                                 ("all fp32", (bias.data_ptr(), rowbias.data_ptr(), res.data_ptr(), 0.5, 1, 1), torch.float32,
                                  0.5 * ref + bias + rowbias[:, None, None, :] + res.float())):
        print(case, name, "...", end=" ", flush=True)
        y = torch.full((N, Ho, Wo, Co), float("nan"), device="cuda", dtype=dt)
        epi = nat.Epilogue(*args)
        nat.check(nat.lib().t2v_conv_fwd(T.P(x), T.P(w), T.P(y), N, H, W, Ci, Co, KH, KW, s, *pads, ctypes.byref(epi), T.stream()))
        torch.cuda.synchronize()
        print("err", T.rel_err(y, want), flush=True)
print("done")
