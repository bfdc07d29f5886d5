"""Regime-B fixtures: the TRAINED weights of the reference's shipped checkpoint (regime B).

    python tests/golden/make_ckpt_fixture.py     # writes tests/golden/ckpt_layers.npz, ckpt_l0.npz, ckpt_l2.npz

With ``example/ckpt/tracking-60k-model.pt`` the per-coordinate scale ``sqrt(2 w)`` of ``prep_qk`` (example/hept.py:21-28)
reaches 5 800 (layer 0) / 660 (layer 2): ``|q^|^2 ~ 1e5..1e8`` against useful scores of O(-10), the regime in which the
reference's ``q.k - |q|^2/2 - |k|^2/2`` cancels catastrophically — the reason the kernels re-centre every block and split
operands three ways.  The checkpoint cannot travel to the GPU box, so this script (run in the build container, where
/root/reference is mounted) commits
  * ``ckpt_layers.npz``  the parameters of layers 0 and 2 (+ ``feat_encoder``, ``regions``) as float32 arrays, and
  * ``ckpt_l0.npz`` / ``ckpt_l2.npz``  outputs of the UNMODIFIED reference module (example/hept.py) run with those weights on
    seeded inputs: q, k, v come from pushing N(0, 0.5^2) features through the checkpoint's feat_encoder -> norm1 -> w_q / w_k /
    w_v (tests/helpers.py::ckpt_qkv evaluates that in float64 and rounds once, so the same bits come out on any host).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as MG  # noqa: E402
from hept_b200 import synthetic  # noqa: E402

CKPT = os.path.join(MG.REF, "example", "ckpt", "tracking-60k-model.pt")
LAYERS = (0, 2)


def write_layers():
    sd = torch.load(CKPT, map_location="cpu", weights_only=True)
    out = {"regions": sd["regions"].numpy()}
    for k, v in sd.items():
        if k.startswith("feat_encoder.") or any(k.startswith(f"attns.{i}.") for i in LAYERS):
            out[k] = v.float().numpy()
    path = os.path.join(HERE, "ckpt_layers.npz")
    np.savez_compressed(path, **out)
    print(f"ckpt_layers: {os.path.getsize(path) / 1e6:.2f} MB, {len(out)} tensors")


def ckpt_case(layer: int, n_raw: int, seed: int):
    from tests import helpers

    ref_hept, ref_utils, ref_tr = MG.load_reference_example()
    cfg = dict(synthetic.TRACKING)
    params = helpers.ckpt_params(layer)
    coords_raw, batch = synthetic.batched_cloud([n_raw], cfg["coords_dim"], seed)
    helper = {"block_size": cfg["block_size"], "regions": params["regions"], "num_heads": cfg["num_heads"]}
    x_dummy = torch.arange(n_raw, dtype=torch.float32)[:, None]
    x_pad, kw, unpad = ref_tr.prepare_input(x_dummy, coords_raw, batch, helper)
    n = x_pad.shape[0]
    q, k, v = helpers.ckpt_qkv(layer, n, seed)
    mod = ref_hept.HEPTAttention(cfg["h_dim"] + cfg["coords_dim"], **cfg)
    mod.load_state_dict({kk: params[kk] for kk in ("out_linear.weight", "out_linear.bias", "e2lsh.alpha")}, strict=True)
    w_rpe = MG.WRpe(params["w_rpe.weight"])
    grad_out = torch.randn(n, cfg["h_dim"], generator=torch.Generator().manual_seed(seed + 5))
    res = MG.run_module(mod, w_rpe, q, k, v, kw, grad_out)
    res["combined_shifts"] = kw["combined_shifts"]
    res["pad_seq"] = x_pad[:, 0].long()
    res["unpad_seq"] = unpad
    res["coords"] = kw["coords"]
    meta = dict(flavour="example", sizes=np.asarray([n_raw]), seed=seed, ckpt_layer=layer,
                chk_q=MG.checksum(q), chk_k=MG.checksum(k), chk_v=MG.checksum(v), chk_coords=MG.checksum(coords_raw),
                chk_alpha=MG.checksum(params["e2lsh.alpha"]), **cfg)
    rows = np.sort(np.random.RandomState(seed).choice(n, size=96, replace=False))
    MG.save(f"ckpt_l{layer}", res, meta, rows=rows, full=False)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    write_layers()
    ckpt_case(0, 2937, seed=50)
    ckpt_case(2, 2937, seed=52)


if __name__ == "__main__":
    main()
