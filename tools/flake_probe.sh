#!/bin/bash
# repeats a test many times in fresh and shared processes to look for nondeterminism
for i in 1 2 3; do python -m pytest tests/test_gpu_scatter.py tests/test_overlay_cpu.py tests/test_gpu_composed.py -m gpu -q 2>&1 | tail -2; done
python -m pytest tests/test_gpu_composed.py::test_mlp_module_forward -m gpu -q --count 1 2>&1 | tail -1
python - <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch.nn.functional as F
import ptgnn_b200 as P
bad = 0
for rep in range(300):
    torch.manual_seed(0)
    for kw in (dict(hidden_layers=[48, 20], use_biases=True), dict(hidden_layers=1), dict(hidden_layers=0), dict(hidden_layers=[7], use_biases=True, activation=torch.nn.Tanh())):
        mlp = P.MLP(36, 24, **kw).eval()
        x = torch.randn(500, 36)
        ref = x
        lins = mlp.linears
        for i, l in enumerate(lins):
            ref = F.linear(ref, l.weight, l.bias)
            if i + 1 < len(lins):
                ref = mlp.activation(ref)
        with torch.no_grad():
            got = mlp.cuda()(x.cuda())
        err = ((got.cpu() - ref.detach()).abs() / ref.detach().abs().clamp(min=1)).max().item()
        if err > 1e-5:
            bad += 1
            print("rep", rep, kw, "err %.3e" % err)
    if rep % 50 == 0:
        junk = [torch.full((1 << 20,), float(rep), device="cuda") for _ in range(16)]; del junk
print("bad", bad, "of 1200")
PY
