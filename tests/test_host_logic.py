"""Host-side control logic (bit-exact with the reference: routing, shards, batch order, swap pairs, placement),
the plan extractor on the plugin models, and the C-ABI surface.  CPU only."""
import re
from pathlib import Path

import pytest
import torch

from util import plugin

REPO = Path(__file__).resolve().parent.parent


def test_routing_and_k():
    from mdgan_b200 import routing

    assert [routing.num_generated_batches(n) for n in (1, 2, 4, 8, 20, 21, 54, 55)] == [2, 2, 2, 2, 2, 3, 3, 4]
    # server.py:238-239: t_n = stack([K[n % k], K[(n + 1) % k]])
    assert [routing.route(n, 2) for n in range(4)] == [(0, 1), (1, 0), (0, 1), (1, 0)]
    assert [routing.route(n, 3) for n in range(4)] == [(0, 1), (1, 2), (2, 0), (0, 1)]
    assert [routing.feedback_slot(n, 2) for n in range(5)] == [0, 1, 0, 1, 0]
    assert routing.num_workers(9) == 8
    with pytest.raises(ValueError):
        routing.num_workers(1)


def test_split_dataset_matches_reference_calls():
    from mdgan_b200 import routing

    g = torch.Generator()
    g.manual_seed(0)
    expect = torch.chunk(torch.randperm(1000, generator=g), 4)      # server.py:46-64,151-154
    got = routing.split_dataset(1000, 4, True)
    assert len(got) == 4 and all(torch.equal(a, b) for a, b in zip(got, expect))
    got = routing.split_dataset(10, 3, False)
    assert [t.tolist() for t in got] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]   # torch.chunk semantics, ragged tail
    before = torch.get_rng_state()
    routing.split_dataset(100, 2, True)
    assert torch.equal(before, torch.get_rng_state()), "the split uses a private generator (global stream untouched)"


def test_swap_schedule_and_pairs_bit_exact():
    from mdgan_b200 import routing

    assert not routing.swap_due(0, 1, 4) and routing.swap_due(1, 1, 4) and not routing.swap_due(3, 2, 4)
    assert not routing.swap_due(5, 1, 1), "a single worker never swaps (server.py:315)"
    torch.manual_seed(3)
    torch.randn(4)                                   # the server's stream is shared with the noise draws
    expect = torch.randperm(8, dtype=torch.int).view(-1, 2) + 1
    torch.manual_seed(3)
    torch.randn(4)
    got = routing.draw_swap_pairs(8)
    assert got.dtype == torch.int32 and torch.equal(got, expect)
    partners = routing.partners_from_pairs(got)
    assert sorted(partners) == list(range(1, 9)) and all(partners[partners[r]] == r for r in partners)
    with pytest.raises(ValueError):
        routing.draw_swap_pairs(3)


def test_real_batch_stream_order_and_ragged():
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing

    ds = SyntheticImages((1, 4, 4), 48)
    shard = routing.split_dataset(48, 2, True)[1]
    g = torch.Generator()
    g.manual_seed(0)
    loader = torch.utils.data.DataLoader(torch.utils.data.Subset(ds, shard), batch_size=8, shuffle=True, generator=g)
    stream = routing.RealBatchStream(ds, shard, 8)
    first_epoch = [b[0] for b in loader]
    for ref in first_epoch:
        assert torch.equal(stream.next(), ref)
    second = [b[0] for b in loader]                   # re-iterating the same loader continues its generator
    assert torch.equal(stream.next(), second[0])
    ragged = routing.RealBatchStream(ds, shard[:20], 8)
    ragged.next(), ragged.next()
    with pytest.raises(ValueError):
        ragged.next()                                # 20 = 8 + 8 + 4: the reference would crash in BCELoss here


def test_real_batch_stream_with_loader_workers_keeps_the_order():
    """MDGAN_LOADER_WORKERS / num_workers only moves the sample transforms to loader processes: the batches are the
    reference's batches, bit for bit, across the epoch boundary too."""
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing

    ds = SyntheticImages((3, 8, 8), 48)
    shard = routing.split_dataset(len(ds), 2, True)[1]
    a = routing.RealBatchStream(ds, shard, 8, num_workers=0)
    b = routing.RealBatchStream(ds, shard, 8, num_workers=2)
    for _ in range(7):  # 3 batches per epoch: crosses two epoch boundaries
        assert torch.equal(a.next(), b.next())
    del b


def test_placement_and_rank_parsing():
    from mdgan_b200 import routing

    assert routing.workers_of_process(0, 1, 4) == [0, 1, 2, 3]
    assert [routing.workers_of_process(p, 8, 8) for p in range(8)] == [[i] for i in range(8)]
    assert [routing.workers_of_process(p, 3, 8) for p in range(3)] == [[0, 1, 2], [3, 4, 5], [6, 7]]
    assert all(routing.process_of_worker(n, 3, 8) == p for p in range(3) for n in routing.workers_of_process(p, 3, 8))
    with pytest.raises(ValueError):
        routing.workers_of_process(0, 5, 4)
    assert routing.parse_ranks("0..4") == [0, 1, 2, 3, 4] and routing.parse_ranks("0,2,5") == [0, 2, 5]
    assert routing.parse_ranks("7") == [7]
    with pytest.raises(ValueError):
        routing.parse_ranks("a-b")


@pytest.mark.parametrize("name,g_kinds,d_kinds", [
    ("CIFAR10", ["dense_up", "up", "up", "up"], ["down", "down", "down", "head"]),
    ("CelebA", ["dense_up", "up", "up", "up", "up"], ["down", "down", "down", "down", "head"]),
    ("MNIST_DCGAN", ["dense_up", "up", "up"], ["down", "down", "head"]),
])
def test_plan_extraction(name, g_kinds, d_kinds):
    from mdgan_b200.plan import extract_plan

    mod = plugin(name)
    torch.manual_seed(0)
    G, D = mod.Generator(), mod.Discriminator()
    sd_before = {k: v.clone() for k, v in D.state_dict().items()}
    rng = torch.get_rng_state()
    pg = extract_plan(G, "generator", (mod.Z_DIM, 1, 1))
    pd = extract_plan(D, "discriminator", tuple(mod.SHAPE))
    assert torch.equal(rng, torch.get_rng_state()), "the dry run must not consume the global RNG"
    assert all(torch.equal(v, sd_before[k]) for k, v in D.state_dict().items()), "nor touch BatchNorm buffers"
    assert [l.kind for l in pg.layers] == g_kinds and [l.kind for l in pd.layers] == d_kinds
    assert pg.out_shape == tuple(mod.SHAPE)
    keys_g, keys_d = set(G.state_dict()), set(D.state_dict())
    for l in pg.layers:
        assert l.weight in keys_g and (l.bn is None or {l.bn.weight, l.bn.running_var} <= keys_g)
    for l in pd.layers:
        assert l.weight in keys_d
    if name == "CelebA":  # functional-style forward: slope 0.01 first layer, biased convs in front of BatchNorm
        assert pd.layers[0].slope == pytest.approx(0.01) and pd.layers[1].slope == pytest.approx(0.2)
        assert pd.layers[1].bias == "cv2.bias" and pd.layers[2].bias == "cv3.bias" and pd.layers[3].bias is None
    assert pd.layers[-1].act == "sigmoid" and pg.layers[-1].act == "tanh"


def test_plan_refuses_unsupported_models():
    from mdgan_b200.plan import UnsupportedModelError, extract_plan

    mlp = plugin("MNIST")
    with pytest.raises(UnsupportedModelError):
        extract_plan(mlp.Discriminator(), "discriminator", (1, 28, 28))
    with pytest.raises(UnsupportedModelError):
        extract_plan(mlp.Generator(), "generator", (100, 1, 1))

    class Odd(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.c = torch.nn.Conv2d(3, 8, 3, 1, 1)

        def forward(self, x):
            return torch.sigmoid(self.c(x)).mean((1, 2, 3))

    with pytest.raises(UnsupportedModelError):
        extract_plan(Odd(), "discriminator", (3, 32, 32))


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads (no GPU needed) and exports exactly what include/mdgan_b200.h declares."""
    from mdgan_b200 import _lib

    header = (REPO / "include" / "mdgan_b200.h").read_text()
    declared = set(re.findall(r"^(?:int|void|long long)\s+(mdgan_\w+)\s*\(", header, flags=re.M))
    assert declared, "header parse failed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()                       # resolves every symbol, raises if one is missing
    assert lib.mdgan_abi_version() == 2     # a host-only call
    assert lib.mdgan_wgrad_splits(128, 8, 8, 128, 64, 0) >= 1 and lib.mdgan_bn_workspace_floats(2, 4096, 128) > 0


def test_no_cpu_fallback():
    from mdgan_b200 import ops
    from mdgan_b200.engine import CudaNetFactory

    with pytest.raises(RuntimeError):
        CudaNetFactory(torch.device("cpu"))
    with pytest.raises(Exception):
        ops.pack_down(torch.zeros(64, 32, 4, 4))      # CPU tensor: refused, not silently computed


def test_bootstrap_cli_keeps_reference_flags():
    import bootstrap

    p = bootstrap.build_parser()
    a = p.parse_args([])
    ref_defaults = dict(backend="nccl", world_size=2, dataset="cifar", ranks="0,1,2", epochs=10, swap_interval=1,
                        local_epochs=10, model="cifar", batch_size=32, log_interval=50, generator_lr=0.001,
                        discriminator_lr=0.004, device="cpu", master_addr="localhost", master_port="1234", iid=1,
                        seed=1, beta_1=0.0, beta_2=0.999)     # /root/reference/src/bootstrap.py:30-51
    for k, v in ref_defaults.items():
        assert getattr(a, k) == v, k
    with pytest.raises(RuntimeError):
        bootstrap.main(["--world_size", "3", "--ranks", "0..2", "--dataset", "CIFAR10", "--device", "cpu"])
    with pytest.raises(ValueError):
        bootstrap.main(["--world_size", "4", "--ranks", "0..3", "--dataset", "CIFAR10", "--device", "cuda"])


def test_torch_custom_ops_cover_the_c_abi():
    """The C-ABI launchers are registered as torch.ops.mdgan_b200.* (CUDA key only: no CPU kernel to fall back to),
    one op per device-side launcher of the header, argument for argument."""
    from mdgan_b200 import _lib, torch_ops

    host_only = {"mdgan_abi_version", "mdgan_check_device", "mdgan_wgrad_splits", "mdgan_pack_job_words",
                 "mdgan_bn_workspace_floats", "mdgan_thin_wgrad_slices", "mdgan_conv_rows_per_tile", "mdgan_conv_stat_phases"}
    assert {f"mdgan_{n}" for n in torch_ops.SPECS} == set(_lib.SIGNATURES) - host_only
    for name, spec in torch_ops.SPECS.items():
        op = getattr(torch.ops.mdgan_b200, name).default
        assert len(op._schema.arguments) + 1 == len(_lib.SIGNATURES[f"mdgan_{name}"][1])   # + the stream
        assert any(a.alias_info is not None and a.alias_info.is_write for a in op._schema.arguments), name
    with pytest.raises(NotImplementedError):
        torch.ops.mdgan_b200.pad_rows(torch.zeros(4, 4), torch.zeros(4, 8), 4, 4, 8, 0)
