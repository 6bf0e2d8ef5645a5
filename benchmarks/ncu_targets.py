#!/usr/bin/env python
"""One launch of each kernel that is new in round 2, for `ncu --set full -k regex:...` (profiles/README.md):
the fused S = 256 attention backward at the c1 shape, the 3xTF32 GEMM (MLP-up forward shape of c1), the streaming
LayerNorm backward (c3 rows) and the vectorised softmax-CE at V = 50257."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from numpitron_b200 import ops  # noqa: E402
from numpitron_b200 import runtime as rt  # noqa: E402

BF16 = torch.bfloat16
rt.set_precision("bf16")
B, S, h, Dh, d = 64, 256, 6, 64, 384
qkv = torch.randn(B, S, h * 3 * Dh, device="cuda").to(BF16)
do = torch.randn(B, S, h * Dh, device="cuda").to(BF16)
div = float(np.float32(d**0.5))
for _ in range(2):
    out, lse = ops.attention_fwd(qkv, B, S, h, Dh, div)
    ops.attention_bwd(qkv, out, do, lse, B, S, h, Dh, div)
T = B * S
a, w = torch.randn(T, d, device="cuda"), torch.randn(d, 4 * d, device="cuda") * 0.02
c = rt.empty((T, 4 * d))
for _ in range(2):
    ops.gemm(a, w, T, 4 * d, d, lda=d, ldb=4 * d, C_out=c, ldc=4 * d, mode=2)
T3, d3 = 32768, 768
x = (1.5 + 0.03 * torch.randn(T3, d3, device="cuda")).to(BF16)
g, b = torch.rand(d3, device="cuda"), torch.zeros(d3, device="cuda")
dy = torch.randn(T3, d3, device="cuda").to(BF16)
_, mean, var = ops.layernorm_fwd(x, g, b, 1e-5)
for _ in range(2):
    ops.layernorm_bwd(x, g, mean, var, dy, 1e-5, bf16_only=True, want_colsum=True)
Tv, V = 4096, 50257
logits = torch.randn(Tv, V, device="cuda") * 3
labels = torch.randint(0, V, (Tv,), device="cuda").to(torch.int32).to(torch.uint16)
ops.softmax_ce_fused(logits, labels, 0, V)
torch.cuda.synchronize()
