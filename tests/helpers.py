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


def load_case(name: str):
    """-> (cfg, inputs, params, grad_out, golden dict of torch tensors)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = {k: int(z["meta_" + k]) for k in CFG_KEYS}
    seed = int(z["meta_seed"])
    flavour = str(z["meta_flavour"])
    gold = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("meta_")}
    params = synthetic.module_params(cfg, seed)
    n = gold["out"].shape[0]
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
