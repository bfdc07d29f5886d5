"""Generate golden fixtures by running the UNMODIFIED reference (Graph-COM/HEPT) in this container.

    python tests/golden/make_golden.py          # writes tests/golden/*.npz

The reference is imported from /root/reference (read-only); it cannot travel
to the GPU box, so its outputs are committed here as compressed ``.npz``
fixtures.  Inputs are regenerated from seeds by ``hept_b200.synthetic`` (the
fixtures carry checksums of the inputs so generator drift is detected rather
than silently compared against).  ``Tensor.argsort`` is wrapped only to RECORD
the permutations the reference computed; its result is passed through
untouched.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("HEPT_REFERENCE", "/root/reference")

from hept_b200 import synthetic  # noqa: E402


# ---------------------------------------------------------------- reference loaders
def load_reference_example():
    """example/hept.py, hept_utils.py, transformer.py (PyG's MLP is stubbed: not on the hot path)."""
    sys.path.insert(0, os.path.join(REF, "example"))
    if "torch_geometric" not in sys.modules:
        pyg = types.ModuleType("torch_geometric")
        pyg_nn = types.ModuleType("torch_geometric.nn")

        class MLP(torch.nn.Module):  # placeholder so `from torch_geometric.nn import MLP` resolves
            def __init__(self, *a, **k):
                super().__init__()

        pyg_nn.MLP = MLP
        pyg.nn = pyg_nn
        sys.modules["torch_geometric"] = pyg
        sys.modules["torch_geometric.nn"] = pyg_nn
    import hept as ref_hept  # type: ignore
    import hept_utils as ref_utils  # type: ignore
    import transformer as ref_tr  # type: ignore

    return ref_hept, ref_utils, ref_tr


def load_reference_src():
    """src/models/attention/hept.py without importing its siblings (they need fast_transformers / PyG)."""
    src = os.path.join(REF, "src")
    for name, path in (("models", "models"), ("models.attention", "models/attention"),
                       ("models.model_utils", "models/model_utils")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(src, path)]
            sys.modules[name] = m

    def by_path(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(src, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    hu = by_path("models.model_utils.hash_utils", "models/model_utils/hash_utils.py")
    at = by_path("models.attention.hept", "models/attention/hept.py")
    return at, hu


class ArgsortRecorder:
    def __enter__(self):
        self.seen = []
        self._orig = torch.Tensor.argsort
        rec = self

        def spy(t, *a, **k):
            out = rec._orig(t, *a, **k)
            rec.seen.append((t.detach().clone(), out.detach().clone()))
            return out

        torch.Tensor.argsort = spy
        return self

    def __exit__(self, *exc):
        torch.Tensor.argsort = self._orig


# ---------------------------------------------------------------- cases
def checksum(t: torch.Tensor) -> float:
    return float(t.double().sum())


class WRpe(torch.nn.Module):
    """Stand-in for the nn.Linear the reference reads ``.weight`` from (example/hept.py:48-54)."""

    def __init__(self, w):
        super().__init__()
        self.weight = torch.nn.Parameter(w.clone())


def run_module(mod, w_rpe, q, k, v, kwargs, grad_out):
    q, k, v = (x.clone().requires_grad_(True) for x in (q, k, v))
    for p in mod.parameters():
        p.grad = None
    w_rpe.weight.grad = None
    with ArgsortRecorder() as rec:
        # hand the module non-leaf tensors, as its caller (Linear outputs) does: the src/
        # flavour writes into a view of ``value`` (src/models/attention/hept.py:91)
        out = mod(q + 0, k + 0, v + 0, w_rpe=w_rpe, pe=None, **kwargs)
    out.backward(grad_out)
    (qk_keys, q_pos), (kk_keys, k_pos) = rec.seen[-2], rec.seen[-1]
    return dict(out=out.detach(), dq=q.grad, dk=k.grad, dv=v.grad, dw_rpe=w_rpe.weight.grad,
                dout_w=mod.out_linear.weight.grad, dout_b=mod.out_linear.bias.grad,
                q_keys=qk_keys, k_keys=kk_keys, q_pos=q_pos, k_pos=k_pos)


def save(name, res, meta, rows=None, full=True):
    """Store results; big gradients are stored as a row subset + column sums unless ``full``."""
    out = {}
    for k, v in meta.items():
        out["meta_" + k] = np.asarray(v)
    for k, v in res.items():
        a = v.numpy()
        if k.endswith("_pos"):
            a = a.astype(np.int32)
        if not full and k in ("dq", "dk", "dv"):
            out[k + "_rows"] = a[rows]
            out[k + "_colsum"] = v.double().sum(0).numpy()
            out[k + "_sqsum"] = np.asarray(float((v.double() ** 2).sum()))
            continue
        out[k] = a
    if rows is not None:
        out["rows"] = rows
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def example_case(name, cfg, sizes, seed, full):
    ref_hept, ref_utils, ref_tr = load_reference_example()
    torch.manual_seed(seed)
    coords_raw, batch = synthetic.batched_cloud(sizes, cfg["coords_dim"], seed)
    params = synthetic.module_params(cfg, seed)
    helper = {"block_size": cfg["block_size"], "regions": params["regions"], "num_heads": cfg["num_heads"]}
    x_dummy = torch.arange(coords_raw.shape[0], dtype=torch.float32)[:, None]
    x_pad, kw, unpad = ref_tr.prepare_input(x_dummy, coords_raw, batch, helper)
    n = x_pad.shape[0]
    q, k, v = synthetic.qkv(n, cfg, seed)
    mod = ref_hept.HEPTAttention(cfg["h_dim"] + cfg["coords_dim"], **cfg)
    mod.load_state_dict({"out_linear.weight": params["out_linear.weight"], "out_linear.bias": params["out_linear.bias"],
                         "e2lsh.alpha": params["e2lsh.alpha"]}, strict=True)
    w_rpe = WRpe(params["w_rpe.weight"])
    g = torch.Generator().manual_seed(seed + 5)
    grad_out = torch.randn(n, cfg["h_dim"], generator=g)
    res = run_module(mod, w_rpe, q, k, v, kw, grad_out)
    res["combined_shifts"] = kw["combined_shifts"]
    res["pad_seq"] = x_pad[:, 0].long()
    res["unpad_seq"] = unpad
    res["coords"] = kw["coords"]
    meta = dict(flavour="example", sizes=np.asarray(sizes), seed=seed,
                chk_q=checksum(q), chk_k=checksum(k), chk_v=checksum(v), chk_coords=checksum(coords_raw),
                chk_alpha=checksum(params["e2lsh.alpha"]), **{k_: v_ for k_, v_ in cfg.items()})
    rows = np.sort(np.random.RandomState(seed).choice(n, size=min(n, 96), replace=False))
    save(name, res, meta, rows=rows, full=full)


def src_case(name, cfg, n_raw, seed, full):
    at, hu = load_reference_src()
    torch.manual_seed(seed)
    coords_raw = synthetic.point_cloud(n_raw, cfg["coords_dim"], seed)
    params = synthetic.module_params(cfg, seed)
    b = cfg["block_size"]
    # HEPT branch of src/models/baselines/transformer.py:43-57, executed with the reference's own helpers.
    coords = hu.pad_to_multiple(coords_raw, b, dims=0, value=float("inf"))
    regions_h = params["regions"].permute(1, 0, 2).reshape(2, -1)
    eta = hu.quantile_partition(torch.argsort(coords[..., 0], dim=-1), regions_h[0][:, None])
    phi = hu.quantile_partition(torch.argsort(coords[..., 1], dim=-1), regions_h[1][:, None])
    coords = coords.clone()
    coords[n_raw:] = 0.0
    n = coords.shape[0]
    kw = dict(coords=coords, raw_size=n_raw, regions_h=regions_h, region_indices=[eta, phi])
    q, k, v = synthetic.qkv(n, cfg, seed)
    mod = at.HEPTAttention(cfg["h_dim"] + cfg["coords_dim"], **cfg)
    sd = mod.state_dict()
    sd.update({"out_linear.weight": params["out_linear.weight"], "out_linear.bias": params["out_linear.bias"],
               "e2lsh.alpha": params["e2lsh.alpha"]})
    mod.load_state_dict(sd, strict=True)
    w_rpe = WRpe(params["w_rpe.weight"])
    g = torch.Generator().manual_seed(seed + 5)
    grad_out = torch.randn(n, cfg["h_dim"], generator=g)
    res = run_module(mod, w_rpe, q, k, v.clone(), kw, grad_out)
    res["region_eta"], res["region_phi"], res["coords"] = eta, phi, coords
    meta = dict(flavour="src", sizes=np.asarray([n_raw]), seed=seed,
                chk_q=checksum(q), chk_k=checksum(k), chk_v=checksum(v), chk_coords=checksum(coords_raw),
                chk_alpha=checksum(params["e2lsh.alpha"]), **{k_: v_ for k_, v_ in cfg.items()})
    rows = np.sort(np.random.RandomState(seed).choice(n, size=min(n, 96), replace=False))
    save(name, res, meta, rows=rows, full=full)


TINY = dict(block_size=10, n_hashes=2, num_regions=6, num_heads=2, h_dim=8, num_w_per_dist=3, coords_dim=6, n_layers=1)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    example_case("tiny_example", TINY, [37, 25, 6], seed=3, full=True)
    src_case("tiny_src", TINY, 53, seed=4, full=True)
    example_case("small_batched", synthetic.TRACKING, [700, 480, 57], seed=11, full=False)
    src_case("small_src", synthetic.TRACKING, 1237, seed=12, full=False)
    example_case("pileup_small", synthetic.PILEUP, [1500], seed=13, full=False)
    # BASELINE.json configs[0]: one tracking-6k-shaped cloud through example/ on CPU, seed 42
    example_case("tracking6k_seed42", synthetic.TRACKING, synthetic.event_sizes("tracking-6k"), seed=42, full=False)


if __name__ == "__main__":
    main()
