"""CPU oracle for the HEPT LSH-bucketed attention path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  ``hept_b200`` never does: the product path
is the sm_100a library behind ``include/hept_b200.h`` and fails loudly when
that library is missing.

What it is: a restatement, in plain torch CPU ops, of the algorithm in the
reference repo (Graph-COM/HEPT).  Every function cites the reference
file:line it follows (paths relative to the reference root).  The reference's
arithmetic lives in torch (bmm / argsort / gather / einsum / exp), so the
restatement uses the same ATen kernels in the same order; it is dtype-generic
so the same code evaluated in float64 measures the fp32 rounding noise of the
reference itself.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md 4).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``tests/golden/make_golden.py`` imports the unmodified reference from
/root/reference in the build container, runs it on seeded inputs and commits
the results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the
oracle against those fixtures (bit-exact for keys/permutations under the
stable tie-break, exact-equal floats for outputs since both run the same ATen
CPU kernels), and, when /root/reference is present, against the live reference.

Conventions fixed here:
  * tie-break of the sort = stable ascending (lowest original index first);
  * eager fp32, TF32 off, no torch.compile;
  * parameters shared through state_dict (alpha, w_rpe.weight, out_linear.*).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# a3  prep_qk                      example/hept.py:21-28, src/models/attention/hept.py:36-43
# --------------------------------------------------------------------------
def coord_scale(w_rpe_weight: Tensor, num_heads: int, dim_per_head: int, num_w_per_dist: int) -> Tensor:
    """Per-head coordinate scale ``sqrt(2 * [qw0, qw0, qw1, ...])`` -> (H, C).

    ``w_rpe_weight`` is (H*D, R*K); viewed (H, D, R, K) exactly like the
    reference's rearrange "(h d) (r k) -> h d r k" (example/hept.py:48-54).
    qw[h, r] = sum_k exp(min(sum_d w[h, d, r, k], 50)) (example/hept.py:22);
    eta and phi share weight 0 (example/hept.py:23).
    """
    hd, rk = w_rpe_weight.shape
    assert hd == num_heads * dim_per_head and rk % num_w_per_dist == 0
    w = w_rpe_weight.view(num_heads, dim_per_head, rk // num_w_per_dist, num_w_per_dist)
    qw = w.sum(dim=1).clamp(max=50).exp().sum(dim=-1)            # (H, R)
    widened = torch.cat([qw[:, :1], qw], dim=-1)                  # (H, C=R+1)
    return torch.sqrt(2 * widened)


def augment_qk(q: Tensor, k: Tensor, scale: Tensor, coords: Tensor) -> Tuple[Tensor, Tensor]:
    """q_hat = [q | scale*coords], k_hat likewise; (N,H,D)->(H,N,E).  example/hept.py:25-27,57-58."""
    n, h, _ = q.shape
    sc = scale[None] * coords[:, None]                            # (N, H, C)
    q_hat = torch.cat([q, sc], dim=-1).permute(1, 0, 2)
    k_hat = torch.cat([k, sc], dim=-1).permute(1, 0, 2)
    return q_hat, k_hat


# --------------------------------------------------------------------------
# a4/a5  E2LSH.forward + lsh_mapping    example/hept_utils.py:45-47, 64-71
# --------------------------------------------------------------------------
def e2lsh_project(x_hat: Tensor, alpha: Tensor) -> Tensor:
    """(H,N,E) x (H,E,T) -> (T,H,N); one bmm then a permute (example/hept_utils.py:46-47)."""
    return torch.bmm(x_hat, alpha.to(x_hat.dtype)).permute(2, 0, 1)


def hash_span(q_proj: Tensor, k_proj: Tensor) -> Tensor:
    """hash_shift[t,h] = max(max q, max k) - min(min q, min k) -> (T,H,1).  example/hept_utils.py:68-70."""
    hi = torch.maximum(q_proj.amax(dim=-1, keepdim=True), k_proj.amax(dim=-1, keepdim=True))
    lo = torch.minimum(q_proj.amin(dim=-1, keepdim=True), k_proj.amin(dim=-1, keepdim=True))
    return hi - lo


# --------------------------------------------------------------------------
# a6 / a6'   AND-construction: shift the projection by region / batch bits
# --------------------------------------------------------------------------
def keys_from_packed_shifts(proj: Tensor, combined_shifts: Tensor, span: Tensor) -> Tensor:
    """example/hept.py:63-65: key = proj + int64_shift * span (convert, multiply, add: three roundings)."""
    return proj + combined_shifts * span


def keys_from_region_indices(
    proj: Tensor, region_eta: Tensor, region_phi: Tensor, regions_h: Tensor, span: Tensor, raw_size: int
) -> Tensor:
    """src/models/attention/hept.py:46-56,93-101.

    ``proj`` (T,H,N); ``region_*`` (T*H, N) float; ``regions_h`` (2, T*H);
    ``span`` (T,H,1).  Padding rows (>= raw_size) get +inf before the shift is
    added (src/.../hept.py:95-96), so they sort last.
    """
    t, h, n = proj.shape
    span_flat = span.reshape(t * h, 1)
    eta_part = region_eta * span_flat
    phi_part = region_phi * span_flat * (torch.ceil(regions_h[0][:, None]) + 1)
    shift = (phi_part + eta_part).view(t, h, n)
    proj = proj.clone()
    proj[..., raw_size:] = float("inf")
    return proj + shift


def stable_argsort(keys: Tensor) -> Tensor:
    """example/hept.py:67-68 with the tie-break pinned to stable ascending."""
    return torch.argsort(keys, dim=-1, stable=True)


# --------------------------------------------------------------------------
# a8  sort_to_buckets / batched_index_select    example/hept_utils.py:74-92
# --------------------------------------------------------------------------
def gather_blocks(x: Tensor, perm: Tensor, block_size: int) -> Tensor:
    """x (H,N,F), perm (T,H,N) -> (T,H,nb,B,F): row gather in sorted order, then cut into blocks."""
    t, h, n = perm.shape
    f = x.shape[-1]
    idx = perm[..., None].expand(t, h, n, f)
    rows = x[None].expand(t, h, n, f).gather(2, idx)
    if n % block_size != 0:
        raise ValueError(f"N={n} is not a multiple of block_size={block_size}")
    return rows.view(t, h, n // block_size, block_size, f)


# --------------------------------------------------------------------------
# a9  qkv_res                                   example/hept.py:7-18
# --------------------------------------------------------------------------
def block_kernel_attention(sq: Tensor, sk: Tensor, sv: Tensor) -> Tuple[Tensor, Tensor]:
    """Per block: S = q.k - |q|^2/2 - |k|^2/2; P = exp(min(S,0)); denom = sum_j P + 1e-20; so = P v."""
    q_half = -0.5 * (sq * sq).sum(dim=-1, keepdim=True)
    k_half = -0.5 * (sk * sk).sum(dim=-1, keepdim=True)
    s = torch.matmul(sq, sk.transpose(-1, -2))
    p = (s + q_half + k_half.transpose(-1, -2)).clamp(max=0.0).exp()
    denom = p.sum(dim=-1, keepdim=True) + 1e-20
    so = torch.matmul(p, sv)
    return denom, so


# --------------------------------------------------------------------------
# a10/a11/a12  invert_permutation, unsort_from_buckets, OR-combine
#              example/hept_utils.py:50-61,95-97; example/hept.py:76-79
# --------------------------------------------------------------------------
def inverse_permutation(perm: Tensor) -> Tensor:
    n = perm.shape[-1]
    inv = torch.empty_like(perm)
    src = torch.arange(n, dtype=perm.dtype, device=perm.device).expand_as(perm)
    inv.scatter_(-1, perm, src)
    return inv


def ungather_blocks(sx: Tensor, inv: Tensor) -> Tensor:
    t, h, nb, b, f = sx.shape
    flat = sx.reshape(t, h, nb * b, f)
    return flat.gather(2, inv[..., None].expand(t, h, nb * b, f))


def or_combine(o: Tensor, denom: Tensor) -> Tensor:
    """(T,H,N,D),(T,H,N,1) -> (H,N,D): sum over tables of numerators / sum over tables of normalisers."""
    return o.sum(dim=0) / denom.sum(dim=0)


# --------------------------------------------------------------------------
# a2  HEPTAttention.forward, both flavours
# --------------------------------------------------------------------------
def attention_core(
    query: Tensor,
    key: Tensor,
    value: Tensor,
    *,
    w_rpe_weight: Tensor,
    alpha: Tensor,
    coords: Tensor,
    block_size: int,
    num_heads: int,
    dim_per_head: int,
    num_w_per_dist: int,
    combined_shifts: Optional[Tensor] = None,
    raw_size: Optional[int] = None,
    regions_h: Optional[Tensor] = None,
    region_indices: Optional[Sequence[Tensor]] = None,
    q_positions: Optional[Tensor] = None,
    k_positions: Optional[Tensor] = None,
    trace: Optional[Dict[str, Tensor]] = None,
) -> Tensor:
    """Everything in HEPTAttention.forward up to (not including) out_linear -> (N, H*D).

    ``combined_shifts`` given  -> example/ flavour (example/hept.py:43-79).
    ``raw_size`` given         -> src/ flavour (src/models/attention/hept.py:71-115);
                                  rows >= raw_size of q_hat, k_hat, v are zeroed
                                  (:89-91) — on copies; the reference's write-through
                                  into the caller's ``value`` is not reproduced.
    ``q_positions/k_positions`` override the sort (stage-wise tests inject the
    reference's permutations so rounding in the keys cannot cascade).
    ``trace`` (dict) receives the intermediates.
    """
    n = query.shape[0]
    h, d = num_heads, dim_per_head
    q = query.reshape(n, h, d)
    k = key.reshape(n, h, d)
    v = value.reshape(n, h, d).permute(1, 0, 2)
    scale = coord_scale(w_rpe_weight, h, d, num_w_per_dist)
    q_hat, k_hat = augment_qk(q, k, scale, coords)

    src_flavour = raw_size is not None
    if src_flavour:
        keep = (torch.arange(n, device=q_hat.device) < raw_size).to(q_hat.dtype)[None, :, None]
        q_hat, k_hat, v = q_hat * keep, k_hat * keep, v * keep
        # multiplying by 0 would turn an inf coordinate into nan; the reference
        # zeroes the padded coords before the call (src/.../transformer.py:57).

    with torch.no_grad():
        q_proj = e2lsh_project(q_hat, alpha)
        k_proj = e2lsh_project(k_hat, alpha)
        span = hash_span(q_proj, k_proj)
        if src_flavour:
            q_keys = keys_from_region_indices(q_proj, region_indices[0], region_indices[1], regions_h, span, raw_size)
            k_keys = keys_from_region_indices(k_proj, region_indices[0], region_indices[1], regions_h, span, raw_size)
        else:
            q_keys = keys_from_packed_shifts(q_proj, combined_shifts, span)
            k_keys = keys_from_packed_shifts(k_proj, combined_shifts, span)
        q_pos = stable_argsort(q_keys) if q_positions is None else q_positions
        k_pos = stable_argsort(k_keys) if k_positions is None else k_positions

    sq = gather_blocks(q_hat, q_pos, block_size)
    sk = gather_blocks(k_hat, k_pos, block_size)
    sv = gather_blocks(v, k_pos, block_size)
    denom, so = block_kernel_attention(sq, sk, sv)
    inv = inverse_permutation(q_pos)
    o = ungather_blocks(so, inv)
    lg = ungather_blocks(denom, inv)
    out = or_combine(o, lg)                                      # (H, N, D)
    if trace is not None:
        trace.update(
            scale=scale, q_hat=q_hat, k_hat=k_hat, q_proj=q_proj, k_proj=k_proj, span=span,
            q_keys=q_keys, k_keys=k_keys, q_pos=q_pos, k_pos=k_pos, denom=lg, numer=o,
        )
    return out.permute(1, 0, 2).reshape(n, h * d)


def attention_forward(query, key, value, *, out_weight: Tensor, out_bias: Tensor, **kw) -> Tensor:
    """Full module forward: attention_core then out_linear (example/hept.py:80)."""
    pre = attention_core(query, key, value, **kw)
    return torch.nn.functional.linear(pre, out_weight, out_bias)


# --------------------------------------------------------------------------
# a13..a17  per-forward preparation
# --------------------------------------------------------------------------
def quantile_regions(sorted_idx: Tensor, num_regions: Tensor) -> Tensor:
    """example/hept_utils.py:6-14.  sorted_idx (n,), num_regions (R,1) float -> (R,n) float.

    region of the point with rank r = floor(r / ceil(n / num_regions)) + 1.
    """
    n = sorted_idx.shape[-1]
    width = torch.ceil(n / num_regions)
    rank_of = torch.argsort(sorted_idx, dim=-1)
    by_rank = torch.arange(n)[None] // width + 1
    return by_rank[:, rank_of]


def pack_bits(low: Tensor, high: Tensor) -> Tensor:
    """example/transformer.py:10-13: (high << bits(low)) | low, bit width chosen per row from max(low)."""
    top = low.max(dim=1, keepdim=True).values
    width = torch.ceil(torch.log2(top + 1)).long()
    return (high << width) | low


def pad_plan(block_size: int, order_key: Tensor, sizes: Tensor, stable: bool = False) -> Tuple[Tensor, Tensor]:
    """example/transformer.py:16-32 (``pad_and_unpad``), restated without in-place index shuffling.

    Every event is padded to a multiple of ``block_size`` by repeating real
    points taken from the argsort of ``order_key`` (table 0 / head 0 packed
    code): the reference fills event i's pad slots with
    ``argsort(order_key)[cum_raw[i] - block_size + j]``, j < pad_i.  For an
    event shorter than ``block_size`` that index reaches into the previous
    event; for event 0 it is negative and wraps like any
    Python index.  Reproduced as is.
    ``stable``: break ties of ``order_key`` by ascending index (the convention the product pins, like for the hash sort)
    instead of whatever torch's default argsort does (what the reference calls).
    Returns (gather index into the raw points (N_pad,), bool mask of real rows).
    """
    sizes = sizes.long()
    padded = (sizes + block_size - 1) // block_size * block_size
    pads = padded - sizes
    total = int(padded.sum())
    order = order_key.argsort(stable=True) if stable else order_key.argsort()
    n_raw = int(sizes.sum())
    take = torch.empty(total, dtype=torch.long)
    real = torch.ones(total, dtype=torch.bool)
    raw_end = sizes.cumsum(0)
    pad_end = padded.cumsum(0)
    for i in range(len(sizes)):
        r0 = int(raw_end[i] - sizes[i])
        p0 = int(pad_end[i] - padded[i])
        s, p = int(sizes[i]), int(pads[i])
        take[p0 : p0 + s] = torch.arange(r0, r0 + s)
        src = int(raw_end[i]) - block_size + torch.arange(p)
        src = torch.where(src < 0, src + n_raw, src)
        take[p0 + s : p0 + s + p] = order[src]
        real[p0 + s : p0 + s + p] = False
    return take, real


def prepare_batched(x: Tensor, coords: Tensor, batch: Tensor, regions: Tensor, block_size: int, num_heads: int,
                    stable: bool = False):
    """example/transformer.py:35-63.  regions (T,2,H) -> kwargs {combined_shifts (T,H,Np) int64, coords (Np,C)}."""
    t, two, h = regions.shape
    reg = regions.permute(1, 0, 2).reshape(two, t * h)             # "c a h -> a (c h)"
    sizes = torch.bincount(batch)
    eta_parts, phi_parts = [], []
    start = 0
    for s in sizes.tolist():
        c = coords[start : start + s]
        eta_parts.append(quantile_regions(torch.argsort(c[:, 0], dim=-1, stable=stable), reg[0][:, None]))
        phi_parts.append(quantile_regions(torch.argsort(c[:, 1], dim=-1, stable=stable), reg[1][:, None]))
        start += s
    eta = torch.cat(eta_parts, dim=-1).long()
    phi = torch.cat(phi_parts, dim=-1).long()
    code = pack_bits(eta, phi)
    code = pack_bits(code, batch[None])
    code = code.view(t, h, -1)
    take, real = pad_plan(block_size, code[0, 0], sizes, stable)
    return x[take], {"combined_shifts": code[..., take], "coords": coords[take]}, real


def prepare_single_event(x: Tensor, coords: Tensor, regions: Tensor, block_size: int):
    """HEPT branch of src/models/baselines/transformer.py:43-57 (one event per step)."""
    n = x.shape[0]
    pad = (-n) % block_size
    t, two, h = regions.shape
    regions_h = regions.permute(1, 0, 2).reshape(two, t * h)
    if pad:
        x = torch.cat([x, x.new_zeros(pad, x.shape[1])])
        coords = torch.cat([coords, coords.new_full((pad, coords.shape[1]), float("inf"))])
    else:
        coords = coords.clone()
    eta = quantile_regions(torch.argsort(coords[:, 0], dim=-1), regions_h[0][:, None])
    phi = quantile_regions(torch.argsort(coords[:, 1], dim=-1), regions_h[1][:, None])
    coords[n:] = 0.0
    return x, {"coords": coords, "raw_size": n, "regions_h": regions_h, "region_indices": [eta, phi]}


# --------------------------------------------------------------------------
# SURVEY.md 8(f)-4  InfoNCE loss of the tracking task      src/utils/losses.py:8-74, src/utils/metrics.py:8-16
# --------------------------------------------------------------------------
def segment_reduce_sorted(src: Tensor, index: Tensor, reduce: str) -> Tensor:
    """``deterministic_scatter`` (losses.py:66-74): sort by index, then torch_scatter.segment_csr over the runs of equal
    indices — restated with a sequential index_add_ over the sorted values (torch_scatter is not installed here;
    segment_csr(src, indptr, "sum" | "mean") is the sum / mean of src[indptr[g]:indptr[g+1]]).  Returns one entry per DISTINCT
    index, in ascending index order."""
    sorted_arg = torch.argsort(index)
    sorted_index = index[sorted_arg]
    sorted_src = src[sorted_arg]
    _, inverse, counts = torch.unique_consecutive(sorted_index, return_inverse=True, return_counts=True)
    out = sorted_src.new_zeros(counts.numel()).index_add_(0, inverse, sorted_src)
    return out / counts.to(src.dtype) if reduce == "mean" else out


def pair_mask(cluster_ids: Tensor, point_pairs: Tensor, recons: Tensor, pts: Tensor, pt_thres: float = 0.9) -> Tensor:
    """Positive pairs: same cluster (losses.py:15) and both points reconstructable with pt above the threshold
    (``pair_filter``, metrics.py:8-16)."""
    a, b = point_pairs[0], point_pairs[1]
    return (cluster_ids[a] == cluster_ids[b]) & (recons[a] != 0) & (recons[b] != 0) & (pts[a] > pt_thres) & (pts[b] > pt_thres)


def infonce_loss(x: Tensor, point_pairs: Tensor, cluster_ids: Tensor, recons: Tensor, pts: Tensor, tau: float,
                 dist_metric: str, compact_like_reference: bool = True) -> Tensor:
    """InfoNCELoss.forward (losses.py:14-39) + calc_info_nce (losses.py:41-53).

    ``compact_like_reference``: the reference indexes the compacted per-point sums (one entry per point owning a negative
    pair) with raw point numbers (losses.py:48-51); False scatters them to point numbers first (what the product does).  The
    two agree whenever every point up to the largest first index owns a negative pair."""
    a, b = point_pairs[0], point_pairs[1]
    pos = pair_mask(cluster_ids, point_pairs, recons, pts)
    neg = ~pos
    if dist_metric == "cosine":
        sim = torch.nn.functional.cosine_similarity(x[a], x[b], dim=-1)
    else:
        dist = torch.linalg.norm(x[a] - x[b], ord=2, dim=-1)
        sim = torch.exp(-dist / (2 * 0.75 ** 2)) if dist_metric == "l2_rbf" else 1.0 / (dist + 1.0)
    scaled = sim / tau
    e = torch.exp(scaled - scaled.max())
    group = a[neg]
    den = segment_reduce_sorted(e[neg], group, "sum").clamp(min=0)
    if not compact_like_reference:
        full = den.new_zeros(x.shape[0])
        full[torch.unique(group)] = den
        den = full
    den = den[a[pos]]
    per_pair = -torch.log(e[pos] / (e[pos] + den))
    labels = torch.unique(cluster_ids[a[pos]], return_inverse=True)[1]
    return segment_reduce_sorted(per_pair, labels, "mean").mean()


def knn_metrics(embeddings: Tensor, cluster_ids: Tensor, mask: Tensor, dist_metric: str, K: int = 19):
    """``acc_and_pr_at_k`` + ``calc_scores`` (src/utils/metrics.py:23-93): for every masked point the K nearest neighbours
    (after dropping the nearest, the point itself) among ALL points; k = cluster size - 1; accuracy = matches in the first k /
    k, precision = matches / K, recall = matches / k; points with k == 0 are skipped; means over the rest."""
    if "l2" in dist_metric:
        dist = torch.cdist(embeddings[mask], embeddings, p=2.0)
    else:
        dist = 1 - torch.nn.functional.cosine_similarity(embeddings[mask].unsqueeze(1), embeddings.unsqueeze(0), dim=-1)
    uniq, counts = torch.unique(cluster_ids, return_counts=True)
    size_of = counts[torch.searchsorted(uniq, cluster_ids)]
    k_list = size_of[mask] - 1
    assert int(k_list.max()) <= K, f"K is too small, max k is {int(k_list.max())}"
    idx = dist.topk(K + 1, dim=1, largest=False, sorted=True)[1][:, 1:]
    matches = cluster_ids[idx] == cluster_ids[mask][:, None]
    keep = k_list > 0
    k = k_list[keep].double()
    in_first_k = (torch.arange(K)[None, :] < k_list[keep][:, None]) & matches[keep]
    acc = (in_first_k.sum(1).double() / k).mean()
    prec = (matches[keep].sum(1).double() / K).mean()
    rec = (matches[keep].sum(1).double() / k).mean()
    return float(acc), float(prec), float(rec)


# --------------------------------------------------------------------------
# convenience: run fwd+bwd and hand back everything a parity test compares
# --------------------------------------------------------------------------
def forward_backward(inputs: Dict[str, Tensor], params: Dict[str, Tensor], cfg: Dict[str, int], grad_out: Tensor,
                     dtype=torch.float32, positions=None):
    """Evaluate the module in ``dtype`` and return out + grads wrt q, k, v, w_rpe.weight, out_linear.*."""
    cast = lambda x: x.detach().to(dtype).clone()
    q, k, v = (cast(inputs[n]).requires_grad_(True) for n in ("query", "key", "value"))
    w = cast(params["w_rpe.weight"]).requires_grad_(True)
    ow = cast(params["out_linear.weight"]).requires_grad_(True)
    ob = cast(params["out_linear.bias"]).requires_grad_(True)
    kw = dict(
        w_rpe_weight=w, alpha=cast(params["e2lsh.alpha"]), coords=cast(inputs["coords"]),
        block_size=cfg["block_size"], num_heads=cfg["num_heads"], dim_per_head=cfg["h_dim"],
        num_w_per_dist=cfg["num_w_per_dist"],
    )
    if "combined_shifts" in inputs:
        kw["combined_shifts"] = inputs["combined_shifts"]
    else:
        kw.update(raw_size=int(inputs["raw_size"]), regions_h=cast(inputs["regions_h"]),
                  region_indices=[cast(r) for r in inputs["region_indices"]])
    if positions is not None:
        kw["q_positions"], kw["k_positions"] = positions
    trace: Dict[str, Tensor] = {}
    out = attention_forward(q, k, v, out_weight=ow, out_bias=ob, trace=trace, **kw)
    out.backward(grad_out.to(dtype))
    return {
        "out": out.detach(), "dq": q.grad, "dk": k.grad, "dv": v.grad, "dw_rpe": w.grad,
        "dout_w": ow.grad, "dout_b": ob.grad,
        "q_pos": trace["q_pos"], "k_pos": trace["k_pos"], "q_keys": trace["q_keys"], "k_keys": trace["k_keys"],
        "span": trace["span"], "scale": trace["scale"].detach(),
    }
