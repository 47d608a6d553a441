"""GPU bring-up probe for the fused aggregation kernel: each case runs in its own process (a trapped kernel poisons the CUDA
context) and prints compact diagnostics.  `python tools/fused_probe.py` (all cases) or `... --case NAME` (one, in-process)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def agg_case(n, counts, H, agg, use_target, bf16, seed=0):
    import torch
    import ptgnn_b200 as P
    from helpers import random_adjacency
    from oracle import ptgnn_oracle as O

    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    adj = random_adjacency(gen, n, counts)
    h = torch.randn(n, H, generator=gen)
    if bf16:
        h = h.to(torch.bfloat16)
    layer = P.MlpMessagePassingLayer(H, 128, 128, len(counts), agg, message_activation=None, use_target_state_as_message_input=use_target,
                                     use_layer_norm=False, use_dense_layer=False).cuda().eval()
    sd = {k: v.clone().cpu() for k, v in layer.state_dict().items()}
    p = "_MlpMessagePassingLayer__"
    ws = [[sd[f"{p}edge_message_transformation_layers.{t}._MLP__mlp_modules.1.weight"]] for t in range(len(counts))]
    ref = O.mlp_layer_forward(h.float(), adj, [torch.empty(c, 0) for c in counts], ws, agg, use_target_state_as_message_input=use_target,
                              message_activation=None)
    with torch.no_grad():
        got = layer(h.cuda(), [(s.cuda(), t.cuda()) for s, t in adj]).float().cpu()
    torch.cuda.synchronize()
    err = (got - ref).abs() / ref.abs().clamp(min=1)
    rel = ((got - ref).norm() / ref.norm().clamp(min=1e-30)).item()
    print(f"  max scaled err {err.max().item():.3e}  rel L2 {rel:.3e}  bad rows(>1e-4) {(err.max(dim=1).values > 1e-4).sum().item()}/{n}"
          f"  bad cols {(err.max(dim=0).values > 1e-4).sum().item()}/128")
    if err.max().item() > (2e-2 if bf16 else 1e-5):
        r = int(err.max(dim=1).values.argmax())
        print("  worst row", r, "got", got[r, :6].tolist(), "ref", ref[r, :6].tolist())
        deg = torch.zeros(n)
        for s_, t_ in adj:
            deg += torch.bincount(t_, minlength=n).float()
        bad = err.max(dim=1).values > 1e-4
        print("  mean in-degree of bad rows", deg[bad].mean().item() if bad.any() else 0, "all", deg.mean().item(),
              " first bad rows", torch.nonzero(bad).flatten()[:10].tolist())
        ratio = (got / ref.where(ref.abs() > 1e-3, torch.ones(()))).flatten()
        print("  median got/ref", ratio.median().item())
        return 1
    return 0


CASES = {
    "plan": None,
    "f32_small_1type": lambda: agg_case(300, [2000], 128, "sum", False, False),
    "f32_tiny": lambda: agg_case(40, [30], 128, "sum", False, False),
    "f32_multi": lambda: agg_case(3000, [9000, 0, 5000, 130, 1], 128, "sum", False, False),
    "f32_max": lambda: agg_case(3000, [9000, 0, 5000, 130, 1], 128, "max", False, False),
    "f32_mean_k64": lambda: agg_case(1000, [3000, 700], 64, "mean", False, False),
    "f32_target": lambda: agg_case(3000, [9000, 5000, 130], 128, "sum", True, False),
    "f32_dense_groups": lambda: agg_case(300, [40000], 128, "sum", False, False),
    "bf16_multi": lambda: agg_case(3000, [9000, 0, 5000, 130, 1], 128, "sum", False, True),
    "bf16_target_max": lambda: agg_case(3000, [9000, 5000, 130], 128, "max", True, True),
    "bf16_k256": lambda: agg_case(2000, [6000, 900], 256, "sum", False, True),
}


def plan_case():
    import numpy as np
    import torch
    import ptgnn_b200 as P
    from helpers import random_adjacency
    from oracle import ptgnn_oracle as O

    rc = 0
    for n, counts in [(1000, [3000, 0, 1500, 700]), (5000, [20000, 1, 130]), (17, [5])]:
        adj = random_adjacency(torch.Generator().manual_seed(n), n, counts)
        plan = P.EdgePlan([(s.cuda(), t.cuda()) for s, t in adj], n)
        plan.block_plan()
        _, group_off, src_f, tl_f, B = plan._block
        ref = O.block_plan(adj, n, B)
        E = sum(counts)
        ok = (np.array_equal(group_off.cpu().numpy(), ref["group_off"]) and np.array_equal(src_f.cpu().numpy()[:E], ref["src_f"])
              and np.array_equal(tl_f.cpu().numpy()[:E], ref["tl_f"]))
        plan.validate()
        print(f"  n={n} B={B} block plan {'ok' if ok else 'MISMATCH'}")
        rc |= 0 if ok else 1
    return rc


if __name__ == "__main__":
    if "--case" in sys.argv:
        name = sys.argv[sys.argv.index("--case") + 1]
        sys.exit(plan_case() if name == "plan" else CASES[name]())
    failed = []
    for name in CASES:
        print(f"== {name}", flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, "--case", name], timeout=300, capture_output=True, text=True)
            out = (r.stdout + r.stderr).strip().splitlines()
            print("\n".join(out[-12:]), flush=True)
            if r.returncode != 0:
                failed.append(name)
        except subprocess.TimeoutExpired:
            print("  TIMEOUT", flush=True)
            failed.append(name)
    print("FAILED:", failed)
