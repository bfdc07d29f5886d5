"""hept_b200.prepare (the product-side mirror of the reference's prepare_input) against the reference's
golden outputs.  Device-agnostic torch code: checked on CPU here, exercised on the GPU by the parity tests."""
import pytest
import torch

from hept_b200 import prepare, synthetic
from tests.helpers import load_case


@pytest.mark.parametrize("name", ["tiny_example", "small_batched", "tracking6k_seed42", "pileup_small"])
def test_batched_prepare_matches_reference(name):
    cfg, inputs, params, grad_out, gold, meta = load_case(name)
    coords_raw, batch = synthetic.batched_cloud(meta["sizes"], cfg["coords_dim"], meta["seed"])
    x = torch.arange(coords_raw.shape[0], dtype=torch.float32)[:, None]
    helper = {"block_size": cfg["block_size"], "regions": params["regions"], "num_heads": cfg["num_heads"]}
    xp, kw, real = prepare.prepare_input(x, coords_raw, batch, helper)
    assert torch.equal(real, gold["unpad_seq"])
    assert kw["combined_shifts"].dtype == torch.int64
    assert torch.equal(kw["combined_shifts"][..., real], gold["combined_shifts"][..., real])
    assert torch.equal(xp[real, 0].long(), gold["pad_seq"][gold["unpad_seq"]])
    # padding rows repeat real points picked through an argsort of the (table 0, head 0) code; the reference's
    # argsort is not stable, so inside a tie of that code it may pick another point than the stable sort here:
    # the (0, 0) code of every padding row and the event it comes from must agree, the point itself need not
    assert torch.equal(kw["combined_shifts"][0, 0], gold["combined_shifts"][0, 0])
    pad_src = xp[~real, 0].long()
    assert torch.equal(batch[pad_src], batch[gold["pad_seq"][~gold["unpad_seq"]]])
    assert kw["coords"].shape == gold["coords"].shape


def test_batched_prepare_with_events_smaller_than_a_block():
    """an event shorter than block_size borrows its padding from the previous event (or wraps
    to the end of the batch for event 0) — reproduced, not fixed."""
    cfg = dict(synthetic.TRACKING)
    sizes = [130, 57, 311]
    coords_raw, batch = synthetic.batched_cloud(sizes, 6, 5)
    params = synthetic.module_params(cfg, 5)
    helper = {"block_size": 100, "regions": params["regions"], "num_heads": 8}
    x = torch.arange(sum(sizes), dtype=torch.float32)[:, None]
    xp, kw, real = prepare.prepare_input(x, coords_raw, batch, helper)
    assert xp.shape[0] == 200 + 100 + 400
    src = xp[:, 0].long()
    ev1_pad = src[200 + 57 : 300]
    assert bool((batch[ev1_pad] == 0).all())          # event 1's 43 padding rows are copies of event-0 points
    from oracle import hept_oracle as O

    _, kw_o, real_o = O.prepare_batched(x, coords_raw, batch, params["regions"], 100, 8)
    assert torch.equal(kw["combined_shifts"][..., real], kw_o["combined_shifts"][..., real])
    assert torch.equal(kw["combined_shifts"][0, 0], kw_o["combined_shifts"][0, 0]) and torch.equal(real, real_o)


@pytest.mark.parametrize("name", ["tiny_src", "small_src"])
def test_single_event_prepare_matches_reference(name):
    cfg, inputs, params, grad_out, gold, meta = load_case(name)
    coords_raw = synthetic.point_cloud(meta["sizes"][0], cfg["coords_dim"], meta["seed"])
    x = torch.ones(coords_raw.shape[0], 3)
    xp, kw = prepare.prepare_input_single(x, coords_raw, {"block_size": cfg["block_size"], "regions": params["regions"]})
    assert kw["raw_size"] == meta["sizes"][0] and xp.shape[0] == gold["coords"].shape[0]
    assert bool((xp[kw["raw_size"]:] == 0).all())
    assert torch.equal(kw["coords"], gold["coords"])
    # padding rows all carry +inf coordinates: their order inside that tie (hence their region index) is the
    # reference's unstable-argsort choice and is irrelevant downstream (their sort key is +inf regardless)
    raw = kw["raw_size"]
    assert torch.equal(kw["region_indices"][0][:, :raw], gold["region_eta"][:, :raw])
    assert torch.equal(kw["region_indices"][1][:, :raw], gold["region_phi"][:, :raw])
    assert kw["region_indices"][0].shape == gold["region_eta"].shape
    assert torch.equal(kw["regions_h"], inputs["regions_h"])
