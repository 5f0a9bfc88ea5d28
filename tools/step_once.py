#!/usr/bin/env python
"""One eager cfg-2 finetune step (forward + backward + global-norm clip + fused AdamW) between cudaProfilerStart/Stop (after a
warm-up step), for
  ncu --profile-from-start off [--metrics ... | --set full -k regex:...] python tools/step_once.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from t2v_b200 import step as S  # noqa: E402
from t2v_b200.optim import FusedAdamW  # noqa: E402

dev = torch.device("cuda", 0)
unet = bench.build_unet(dev, "--small" in sys.argv)
stepper = S.DataParallelStep(unet, S.ddpm_alphas_cumprod(device=dev), passes=1, use_graph=False)
stepper.attach_optimizer(FusedAdamW(stepper.arena, [dict(params=list(unet.parameters()))], lr=5e-6, weight_decay=1e-2, max_grad_norm=1.0))
inputs = [x.to(dev) for x in bench.synthetic_inputs(1, bench.CFG2, 1234)]
stepper(*inputs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = stepper(*inputs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
