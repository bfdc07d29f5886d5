"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares (no compute calls —
there is no GPU here), argument validation in the library, the module's constructor / state_dict contract, the
"no CPU fallback" rule, and the event-sharding logic over a 2-process gloo group."""
import ctypes
import os
import re

import pytest
import torch

from hept_b200 import HEPTAttention, _lib, ops, sharding, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "hept_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hept_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hept_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signatures and header disagree"
    assert lib.hept_abi_version() == 1


def test_supported_shapes_and_workspace_queries():
    lib = _lib.load()
    assert lib.hept_shape_supported(24, 6, 100) == 1      # tracking
    assert lib.hept_shape_supported(24, 4, 100) == 1      # pileup
    assert lib.hept_shape_supported(8, 6, 10) == 1        # test shape
    assert lib.hept_shape_supported(32, 6, 100) == 0
    s = _lib.Shape(60000, 8, 24, 6, 3, 100, 60000)
    fwd = lib.hept_attention_fwd_workspace_bytes(ctypes.byref(s))
    bwd = lib.hept_attention_bwd_workspace_bytes(ctypes.byref(s))
    assert 180e6 < fwd < 400e6 and 450e6 < bwd < 700e6
    assert lib.hept_argsort_workspace_bytes(48, 60000) > 4 * 48 * 60000 * 4


def test_library_rejects_bad_arguments_without_touching_a_gpu():
    lib = _lib.load()
    bad = _lib.Shape(6050, 8, 24, 6, 3, 100, 6050)        # N not a multiple of block_size
    one = ctypes.c_void_p(8)                              # non-null dummy pointers: validation fails first
    rc = lib.hept_block_attention_fwd(ctypes.byref(bad), one, one, one, one, one, one, one, one, None)
    assert rc == _lib.HEPT_EINVAL and b"multiple of block_size" in lib.hept_last_error()
    rc = lib.hept_segmented_argsort(None, 4, 100, one, one, 0, None)
    assert rc == _lib.HEPT_EINVAL
    ok = _lib.Shape(6100, 8, 24, 6, 3, 100, 6100)
    rc = lib.hept_block_attention_bwd(ctypes.byref(ok), *([one] * 14), 16, None)
    assert rc == _lib.HEPT_EWORKSPACE
    odd = _lib.Shape(6000, 8, 16, 6, 3, 100, 6000)        # D=16 not compiled in: no slow fallback
    rc = lib.hept_block_attention_fwd(ctypes.byref(odd), one, one, one, one, one, one, one, one, None)
    assert rc == _lib.HEPT_EUNSUPPORTED
    with pytest.raises(ValueError):
        _lib.check(_lib.HEPT_EINVAL, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.HEPT_EUNSUPPORTED, "x")


def test_module_surface_matches_reference():
    cfg = dict(synthetic.TRACKING)
    m = HEPTAttention(cfg["h_dim"] + cfg["coords_dim"], **cfg)         # extra keys (num_regions, pe_type...) ignored
    assert sorted(m.state_dict()) == ["e2lsh.alpha", "out_linear.bias", "out_linear.weight"]
    assert tuple(m.e2lsh.alpha.shape) == (8, 30, 3) and not m.e2lsh.alpha.requires_grad
    assert tuple(m.out_linear.weight.shape) == (24, 192)
    p = synthetic.module_params(cfg, 0)
    m.load_state_dict({k: p[k] for k in ("out_linear.weight", "out_linear.bias", "e2lsh.alpha")}, strict=True)
    # a src/ checkpoint also carries the unused e2lsh.beta (hash_utils.py:344): still strict-loadable
    sd = dict(m.state_dict())
    sd["e2lsh.beta"] = torch.rand(1, 3)
    m.load_state_dict(sd, strict=True)
    assert "e2lsh.beta" in m.state_dict()
    m2 = HEPTAttention(30, e2lsh_beta=True, **cfg)
    assert "e2lsh.beta" in m2.state_dict()
    with pytest.raises(RuntimeError):
        m.load_state_dict({"out_linear.weight": p["out_linear.weight"]}, strict=True)


@pytest.mark.skipif(not os.path.exists("/root/reference/example/ckpt/tracking-60k-model.pt"),
                    reason="reference checkpoint only exists in the build container")
def test_reference_checkpoint_loads_strictly():
    sd = torch.load("/root/reference/example/ckpt/tracking-60k-model.pt", map_location="cpu")
    for layer in range(4):
        pre = f"attns.{layer}.attn."
        sub = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
        m = HEPTAttention(30, **synthetic.TRACKING)
        m.load_state_dict(sub, strict=True)
        assert torch.equal(m.e2lsh.alpha, sub["e2lsh.alpha"])


def test_no_cpu_fallback():
    cfg = dict(synthetic.TRACKING)
    m = HEPTAttention(30, **cfg)
    x = torch.zeros(100, 192)
    w_rpe = torch.nn.Linear(50, 192)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x, x, x, w_rpe=w_rpe, coords=torch.zeros(100, 6), combined_shifts=torch.zeros(3, 8, 100, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.segmented_argsort(torch.zeros(4, 10))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hept_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_round_robin_event_assignment():
    owned = [sharding.events_of_rank(11, r, 4) for r in range(4)]
    assert sorted(sum(owned, [])) == list(range(11))
    assert owned[1] == [1, 5, 9]
    with pytest.raises(ValueError):
        sharding.events_of_rank(4, 4, 4)


def _free_port():
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    frozen = torch.nn.Parameter(torch.ones(4), requires_grad=False)      # like e2lsh.alpha: never has a grad
    events = sharding.events_of_rank(6, rank, world)
    x = torch.stack([torch.full((5,), float(e + 1)) for e in events])
    lin(x).sum().backward()
    nbytes = sharding.allreduce_gradients([lin.weight, lin.bias, frozen, None], average=False)
    # plain lists, not tensors: a tensor in a queue is a handle to the sender's shared memory, gone if the sender exits first
    out.put((rank, lin.weight.grad.tolist(), lin.bias.grad.tolist(), nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    """world_size 2 over gloo: gradients summed across ranks equal the single-process gradient over all events."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((out.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    x = torch.stack([torch.full((5,), float(e + 1)) for e in range(6)])
    lin(x).sum().backward()
    for rank, gw, gb, nbytes in got:
        assert torch.allclose(torch.tensor(gw), lin.weight.grad) and torch.allclose(torch.tensor(gb), lin.bias.grad)
        assert nbytes == (15 + 3) * 4


def _bucket_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    unused = torch.nn.Parameter(torch.ones(4))                           # like w_rpe.bias: trainable, never used
    bucket = sharding.GradBucket([lin.weight, lin.bias, unused])
    res = []
    for step in range(2):                                                # two steps: zero() must really reset the views
        bucket.zero()
        events = sharding.events_of_rank(6, rank, world)
        x = torch.stack([torch.full((5,), float(e + 1 + step)) for e in events])
        lin(x).sum().backward()
        assert bucket.attached()
        bucket.allreduce(average=True)
        res.append((lin.weight.grad.tolist(), lin.bias.grad.tolist(), unused.grad.tolist()))
    out.put((rank, res, bucket.nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_bucket_two_ranks_gloo():
    """GradBucket: .grad tensors are views of one flat buffer, all-reduced in place (world_size 2 over gloo); the average
    over the ranks equals the single-process gradient over all events divided by the world size."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((out.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(2):
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        x = torch.stack([torch.full((5,), float(e + 1 + step)) for e in range(6)])
        lin(x).sum().backward()
        for rank, res, nbytes in got:
            gw, gb, gu = res[step]
            assert torch.allclose(torch.tensor(gw), lin.weight.grad / 2) and torch.allclose(torch.tensor(gb), lin.bias.grad / 2)
            assert all(v == 0 for v in gu) and nbytes == (15 + 3 + 4) * 4


def test_trace_variant_of_the_library_builds():
    """The pipeline timeline probe (csrc/trace.cuh, `make TRACE=1`) is compiled out of the product library; this keeps the
    instrumented variant building (it exports the two setters tools/pipeline_trace.py binds)."""
    import ctypes
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "hept_b200", "csrc")
    subprocess.run(["make", "-C", csrc, f"-j{os.cpu_count() or 4}", "TRACE=1"], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(root, "hept_b200", "libhept_sm100_trace.so"))
    for name in ("hept_debug_trace_fwd", "hept_debug_trace_bwd", "hept_attention_fwd"):
        assert hasattr(lib, name), name
    from hept_b200 import _lib

    assert not hasattr(_lib.load(), "hept_debug_trace_fwd"), "the product library must not carry the probe"
