"""Shared test plumbing: rebuild the inputs of a golden fixture from its seed and load its outputs."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

from hept_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFG_KEYS = ("block_size", "n_hashes", "num_regions", "num_heads", "h_dim", "num_w_per_dist", "coords_dim", "n_layers")
CASES = ("tiny_example", "tiny_src", "small_batched", "small_src", "pileup_small", "tracking6k_seed42")
# regime B: the trained weights of the reference's checkpoint, layers 0 and 2 (make_ckpt_fixture.py)
CKPT_CASES = ("ckpt_l0", "ckpt_l2")
ALL_CASES = CASES + CKPT_CASES


_CKPT = None


def ckpt_tensors() -> Dict[str, torch.Tensor]:
    """Parameters of the reference's shipped checkpoint, layers 0 and 2 + feat_encoder + regions (ckpt_layers.npz)."""
    global _CKPT
    if _CKPT is None:
        z = np.load(os.path.join(GOLDEN, "ckpt_layers.npz"))
        _CKPT = {k: torch.from_numpy(z[k]) for k in z.files}
    return _CKPT


def ckpt_params(layer: int) -> Dict[str, torch.Tensor]:
    """The attention module's parameters of checkpoint layer ``layer`` under the names module_params() uses."""
    t = ckpt_tensors()
    p = f"attns.{layer}."
    return {"w_rpe.weight": t[p + "w_rpe.weight"], "w_rpe.bias": t[p + "w_rpe.bias"],
            "out_linear.weight": t[p + "attn.out_linear.weight"], "out_linear.bias": t[p + "attn.out_linear.bias"],
            "e2lsh.alpha": t[p + "attn.e2lsh.alpha"], "regions": t["regions"]}


def ckpt_qkv(layer: int, n: int, seed: int):
    """q, k, v as the checkpoint's own layers produce them from N(0, 0.5^2) features: feat_encoder -> norm1 -> w_q / w_k /
    w_v (example/transformer.py:113,157-158), evaluated in float64 and rounded once to float32 (host-independent bits)."""
    t = {k: v.double() for k, v in ckpt_tensors().items()}
    g = torch.Generator().manual_seed(seed + 104729)
    feats = (torch.randn(n, t["feat_encoder.0.weight"].shape[1], generator=g) * 0.5).double()
    h = torch.relu(feats @ t["feat_encoder.0.weight"].T + t["feat_encoder.0.bias"])
    h = h @ t["feat_encoder.2.weight"].T + t["feat_encoder.2.bias"]
    p = f"attns.{layer}."
    xn = torch.nn.functional.layer_norm(h, (h.shape[1],), t[p + "norm1.weight"], t[p + "norm1.bias"])
    return tuple((xn @ t[p + w].T).float().contiguous() for w in ("w_q.weight", "w_k.weight", "w_v.weight"))


def load_case(name: str):
    """-> (cfg, inputs, params, grad_out, golden dict of torch tensors)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = {k: int(z["meta_" + k]) for k in CFG_KEYS}
    seed = int(z["meta_seed"])
    flavour = str(z["meta_flavour"])
    gold = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("meta_")}
    n = gold["out"].shape[0]
    if "meta_ckpt_layer" in z.files:
        layer = int(z["meta_ckpt_layer"])
        params = ckpt_params(layer)
        q, k, v = ckpt_qkv(layer, n, seed)
    else:
        params = synthetic.module_params(cfg, seed)
        q, k, v = synthetic.qkv(n, cfg, seed)
    for nm, t in (("q", q), ("k", k), ("v", v), ("alpha", params["e2lsh.alpha"])):
        want = float(z["meta_chk_" + nm])
        got = float(t.double().sum())
        # the float64 sum is taken by torch on whatever host runs the test: its reduction order (vector width, threads)
        # moves the last bits, a drifted generator moves the leading ones
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), \
            f"synthetic generator drifted for {name}:{nm} ({got} vs {want}); regenerate the fixtures"
    inputs: Dict[str, torch.Tensor] = {"query": q, "key": k, "value": v, "coords": gold["coords"]}
    if flavour == "example":
        inputs["combined_shifts"] = gold["combined_shifts"]
    else:
        inputs["raw_size"] = int(z["meta_sizes"][0])
        inputs["regions_h"] = params["regions"].permute(1, 0, 2).reshape(2, -1)
        inputs["region_indices"] = [gold["region_eta"], gold["region_phi"]]
    g = torch.Generator().manual_seed(seed + 5)
    grad_out = torch.randn(n, cfg["h_dim"], generator=g)
    meta = {"flavour": flavour, "seed": seed, "sizes": [int(s) for s in z["meta_sizes"]]}
    return cfg, inputs, params, grad_out, gold, meta


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Relative Frobenius error of a against b, in float64."""
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
