"""hept_b200.prepare — the reference's prepare_input on the library's CUDA kernels (hept_prepare_batched /
hept_prepare_single, called through the C ABI) — against the reference's golden outputs and against the oracle.
Index work: everything is compared with torch.equal."""
import pytest
import torch

from hept_b200 import synthetic
from oracle import hept_oracle as O
from tests.helpers import load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _helper(cfg, params):
    return {"block_size": cfg["block_size"], "regions": params["regions"].to(DEV), "num_heads": cfg["num_heads"]}


@pytest.mark.parametrize("name", ["tiny_example", "small_batched", "tracking6k_seed42", "pileup_small", "ckpt_l0", "ckpt_l2"])
def test_batched_prepare_matches_reference(name):
    from hept_b200 import prepare

    cfg, inputs, params, grad_out, gold, meta = load_case(name)
    coords_raw, batch = synthetic.batched_cloud(meta["sizes"], cfg["coords_dim"], meta["seed"])
    x = torch.arange(coords_raw.shape[0], dtype=torch.float32)[:, None]
    xp, kw, real = prepare.prepare_input(x.to(DEV), coords_raw.to(DEV), batch.to(DEV), _helper(cfg, params))
    xp, real = xp.cpu(), real.cpu()
    shifts = kw["combined_shifts"].cpu()
    assert real.dtype == torch.bool and torch.equal(real, gold["unpad_seq"])
    assert shifts.dtype == torch.int64 and shifts.shape == gold["combined_shifts"].shape
    assert torch.equal(shifts[..., real], gold["combined_shifts"][..., real])
    assert torch.equal(kw["combined_shifts32"].cpu().long(), shifts)
    assert torch.equal(xp[real, 0].long(), gold["pad_seq"][gold["unpad_seq"]])
    # padding rows repeat real points picked through an argsort of the (table 0, head 0) code; the reference's
    # argsort is not stable, so inside a tie of that code it may pick another point than the stable sort here:
    # the (0, 0) code of every padding row and the event it comes from must agree, the point itself need not
    assert torch.equal(shifts[0, 0], gold["combined_shifts"][0, 0])
    pad_src = xp[~real, 0].long()
    assert torch.equal(batch[pad_src], batch[gold["pad_seq"][~gold["unpad_seq"]]])
    assert torch.equal(kw["coords"].cpu()[real], gold["coords"][real])
    assert torch.equal(kw["coords"].cpu(), coords_raw[xp[:, 0].long()])


@pytest.mark.parametrize("sizes", [[130, 57, 311], [57], [100, 200], [21000, 15000, 11300, 9000, 3000, 700, 130, 57], [61237]])
def test_batched_prepare_against_the_oracle(sizes):
    """Against the oracle's restatement of prepare_input (pinned to the reference by tests/test_oracle_golden.py), incl.
    SURVEY.md 7.3-7: an event shorter than block_size borrows its padding from the previous event (or wraps to the end of
    the batch for event 0) — reproduced, not fixed — and BASELINE.json configs[3] unscaled."""
    from hept_b200 import prepare

    cfg = dict(synthetic.TRACKING)
    coords_raw, batch = synthetic.batched_cloud(sizes, 6, 5)
    params = synthetic.module_params(cfg, 5)
    x = torch.arange(sum(sizes), dtype=torch.float32)[:, None]
    xp, kw, real = prepare.prepare_input(x.to(DEV), coords_raw.to(DEV), batch.to(DEV), _helper(cfg, params))
    xo, kw_o, real_o = O.prepare_batched(x, coords_raw, batch, params["regions"], 100, 8)
    xp, real, shifts = xp.cpu(), real.cpu(), kw["combined_shifts"].cpu()
    assert torch.equal(real, real_o)
    assert torch.equal(shifts[..., real], kw_o["combined_shifts"][..., real])
    assert torch.equal(shifts[0, 0], kw_o["combined_shifts"][0, 0])
    # the oracle's argsort of the (0,0) code is torch's default (unstable) one: same events, same codes, maybe other points
    assert torch.equal(batch[xp[:, 0].long()], batch[xo[:, 0].long()])
    if sizes == [130, 57, 311]:
        assert xp.shape[0] == 200 + 100 + 400
        ev1_pad = xp[:, 0].long()[200 + 57: 300]
        assert bool((batch[ev1_pad] == 0).all())          # event 1's 43 padding rows are copies of event-0 points
    # with the oracle's sorts pinned to the stable tie-break everything is equal, padding rows included
    xs, kw_s, real_s = O.prepare_batched(x, coords_raw, batch, params["regions"], 100, 8, stable=True)
    assert torch.equal(xp, xs) and torch.equal(shifts, kw_s["combined_shifts"]) and torch.equal(kw["coords"].cpu(), kw_s["coords"])
    # passing the sizes avoids the read-back of bincount and changes nothing
    xp2, kw2, real2 = prepare.prepare_input(x.to(DEV), coords_raw.to(DEV), batch.to(DEV), _helper(cfg, params), sizes=sizes)
    assert torch.equal(kw2["combined_shifts"].cpu(), shifts) and torch.equal(xp2.cpu(), xp)


@pytest.mark.parametrize("name", ["tiny_src", "small_src"])
def test_single_event_prepare_matches_reference(name):
    from hept_b200 import prepare

    cfg, inputs, params, grad_out, gold, meta = load_case(name)
    coords_raw = synthetic.point_cloud(meta["sizes"][0], cfg["coords_dim"], meta["seed"])
    x = torch.ones(coords_raw.shape[0], 3)
    xp, kw = prepare.prepare_input_single(x.to(DEV), coords_raw.to(DEV), {"block_size": cfg["block_size"],
                                                                          "regions": params["regions"].to(DEV)})
    assert kw["raw_size"] == meta["sizes"][0] and xp.shape[0] == gold["coords"].shape[0]
    assert bool((xp[kw["raw_size"]:] == 0).all())
    assert torch.equal(kw["coords"].cpu(), gold["coords"])
    # padding rows all carry +inf coordinates: their order inside that tie (hence their region index) is the
    # reference's unstable-argsort choice and is irrelevant downstream (their sort key is +inf regardless)
    raw = kw["raw_size"]
    assert torch.equal(kw["region_indices"][0].cpu()[:, :raw], gold["region_eta"][:, :raw])
    assert torch.equal(kw["region_indices"][1].cpu()[:, :raw], gold["region_phi"][:, :raw])
    assert kw["region_indices"][0].shape == gold["region_eta"].shape
    assert torch.equal(kw["regions_h"].cpu(), inputs["regions_h"])


@pytest.mark.parametrize("n_raw", [9950, 60000, 61237])
def test_single_event_prepare_against_the_oracle(n_raw):
    from hept_b200 import prepare

    cfg = dict(synthetic.PILEUP)
    coords_raw = synthetic.point_cloud(n_raw, 4, 31)
    params = synthetic.module_params(cfg, 31)
    x = torch.zeros(n_raw, 3)
    xp, kw = prepare.prepare_input_single(x.to(DEV), coords_raw.to(DEV), {"block_size": 100, "regions": params["regions"].to(DEV)})
    xo, kw_o = O.prepare_single_event(x, coords_raw, params["regions"], 100)
    assert xp.shape == xo.shape and torch.equal(kw["coords"].cpu(), kw_o["coords"])
    for a in (0, 1):
        assert torch.equal(kw["region_indices"][a].cpu()[:, :n_raw], kw_o["region_indices"][a][:, :n_raw])


def test_no_cpu_path():
    from hept_b200 import prepare

    cfg = dict(synthetic.TRACKING)
    coords_raw, batch = synthetic.batched_cloud([130], 6, 5)
    params = synthetic.module_params(cfg, 5)
    with pytest.raises(RuntimeError):
        prepare.prepare_input(torch.zeros(130, 1), coords_raw, batch, {"block_size": 100, "regions": params["regions"], "num_heads": 8})
