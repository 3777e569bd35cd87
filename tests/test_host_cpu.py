"""CPU tests of the host-side mirror of the reference interface: flags, registry, variable names/shapes,
C-ABI surface (symbols only: no compute without a GPU), data-parallel bucket logic on gloo (world size 2)."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flags_have_reference_names_and_defaults():
    from learnablepoolingmethods_b200.flags import FLAGS, ensure_parsed
    ensure_parsed()
    # frame_level_models.py:35-36,2197-2216 ; video_level_models.py:26-45
    assert FLAGS.iterations == 30
    assert FLAGS.netvlad_cluster_size == 256 and FLAGS.netvlad_hidden_size == 1024
    assert FLAGS.netvlad_add_batch_norm is True and FLAGS.netvlad_relu is False
    assert FLAGS.gating is True and FLAGS.gating_remove_diag is False
    assert FLAGS.moe_num_mixtures == 2 and FLAGS.moe_l2 == 1e-8 and FLAGS.moe_low_rank_gating == -1
    assert FLAGS.moe_prob_gating is False and FLAGS.sample_random_frames is True


def test_registry_lookup_like_train_py():
    from learnablepoolingmethods_b200 import frame_level_models, models, utils, video_level_models
    for name in ("NetVladV1", "NetVladV2"):
        cls = utils.find_class_by_name(name, [frame_level_models, video_level_models])
        assert issubclass(cls, models.BaseModel)
        import inspect
        sig = inspect.signature(cls().create_model)
        assert list(sig.parameters)[:9] == ["model_input", "vocab_size", "num_frames", "iterations", "add_batch_norm",
                                            "sample_random_frames", "cluster_size", "hidden_size", "is_training"]
        assert sig.parameters["is_training"].default is True
    assert utils.find_class_by_name("MoeModel", [frame_level_models, video_level_models]) is video_level_models.MoeModel
    with pytest.raises(StopIteration):
        utils.find_class_by_name("NoSuchModel", [frame_level_models, video_level_models])


@pytest.mark.parametrize("model", ["NetVladV1", "NetVladV2"])
def test_variable_names_and_shapes_match_tf_naming(model):
    """State-dict keys = the reference's TF variable names; the oracle lists them independently."""
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from learnablepoolingmethods_b200.variables import VariableStore
    from oracle import netvlad_oracle as O
    store = VariableStore("cpu")
    NetVladEngine(NetVladConfig(model=model, iterations=32, cluster_size=16, hidden_size=24, vocab_size=40), store)
    specs = O.param_specs(model, iterations=32, cluster_size=16, hidden_size=24, vocab_size=40)
    assert set(store.vars) == set(specs)
    for k, (shape, kind, arg) in specs.items():
        assert tuple(store.vars[k].shape) == tuple(shape), k
    assert store.vars["audio_VLAD/cluster_weights" if model == "NetVladV1" else "audio_VLAD/cluster_centers"].shape[1] == 4
    # initialisers: BN/LN affine at (1, 0), biases at 0, cluster weights ~ N(0, 1/sqrt(D))
    assert float(store.vars["gating_bn/gamma"].min()) == 1.0 and float(store.vars["experts/biases"].abs().max()) == 0.0
    sd = store.state_dict(prefix="tower/")
    assert all(k.startswith("tower/") for k in sd)
    store.load_state_dict(sd, prefix="tower/")


def test_same_seed_same_weights_on_every_rank():
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from learnablepoolingmethods_b200.variables import VariableStore
    a, b = VariableStore("cpu", seed=1810), VariableStore("cpu", seed=1810)
    cfg = NetVladConfig(iterations=16, cluster_size=8, hidden_size=16, vocab_size=10)
    NetVladEngine(cfg, a); NetVladEngine(cfg, b)
    assert all(torch.equal(a.vars[k], b.vars[k]) for k in a.vars)


def test_learning_rate_schedule_matches_reference():
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from learnablepoolingmethods_b200.trainer import Trainer
    from learnablepoolingmethods_b200.variables import VariableStore
    from oracle import netvlad_oracle as O
    eng = NetVladEngine(NetVladConfig(iterations=16, cluster_size=8, hidden_size=16, vocab_size=10), VariableStore("cpu"))
    tr = Trainer(eng, base_learning_rate=2e-4, learning_rate_decay=0.85, learning_rate_decay_examples=4000000, batch_size=80)
    for step in (0, 1, 49999, 50000, 100001):
        tr.global_step = step
        assert tr.learning_rate() == O.learning_rate(2e-4, 0.85, 4000000, step, 80, 1)


def test_cabi_library_exports_every_declared_symbol():
    """include/lpm_b200.h is the boundary: every function it declares must be exported (no compute calls here)."""
    from learnablepoolingmethods_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "lpm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(lpm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.lpm_version.restype = ctypes.c_int
    assert lib.lpm_version() >= 100
    lib.lpm_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.lpm_last_error(), bytes)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "learnablepoolingmethods_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r"#.*", "", src).replace("the oracle", ""), fn


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from learnablepoolingmethods_b200.dp import BucketedAllReduce
    n = 1000
    g = torch.arange(n, dtype=torch.float32) * (rank + 1)
    red = BucketedAllReduce(g, bucket_elems=256)
    launched_at = []
    for end in (100, 300, 512, 900, 1000):       # gradients become final in flat order
        red.mark_done(end)
        launched_at.append(len(red.launched))
    red.flush()
    red.wait()
    expect = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(g, expect)
    # a skipped range (summed another way, dp.FactorGather) is left untouched and never crosses a bucket
    g2 = torch.arange(n, dtype=torch.float32) * (rank + 1)
    red2 = BucketedAllReduce(g2, bucket_elems=256, skip=((100, 600),))
    red2.mark_done(650)
    n_early = len(red2.launched)
    red2.flush()
    red2.wait()
    expect2 = expect.clone()
    expect2[100:600] = torch.arange(100, 600, dtype=torch.float32) * (rank + 1)
    ok = ok and torch.equal(g2, expect2) and n_early == 1 and red2.launched == [(0, 100), (600, 856), (856, 1000)]
    q.put((rank, ok, launched_at, list(red.launched)))
    dist.destroy_process_group()


def test_bucketed_allreduce_sum_world2_gloo():
    """utils.py:205-211 sums the tower gradients: all-reduce SUM; buckets launch only when fully produced."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, launched_at, launched in res:
        assert ok, rank
        assert launched_at == [0, 1, 2, 3, 4]      # buckets complete at 256, 512, 768 and the 232-element tail at 1000
        assert launched == [(0, 256), (256, 512), (512, 768), (768, 1000)]


def _cpu_store(model="NetVladV1"):
    from learnablepoolingmethods_b200 import variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    cfg = NetVladConfig(model=model, iterations=16, cluster_size=32, hidden_size=32, vocab_size=50, rgb_dim=64, audio_dim=16,
                        rgb_heads=4, audio_heads=2)
    store = variables.VariableStore("cpu", seed=1)
    NetVladEngine.build_variables(type("E", (), {"cfg": cfg, "store": store, "_dense_vars": NetVladEngine._dense_vars,
                                                  "_ln_vars": NetVladEngine._ln_vars, "wc_suffix": NetVladEngine.wc_suffix})())
    return store


def _worker_split_allreduce(rank, world, port, q):
    """The data-parallel graph step's exchange on gloo: the head span of the flat gradient is reduced first (it travels
    under the modalities' backward on the GPU), the rest afterwards; together they are the plain SUM all-reduce."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from learnablepoolingmethods_b200.trainer import FlatState, head_gradient_span
    store = _cpu_store()
    tr = store.trainable()
    head = [n for n in tr if not n.startswith(("video_", "audio_", "input_bn"))]
    body = [n for n in tr if n not in head]
    flat = FlatState(store, [n for n in head if n != "hidden1_weights"] + body, {}, factored=("hidden1_weights",))
    end, ok = head_gradient_span(flat)
    flat.g.copy_(torch.arange(flat.g.numel(), dtype=torch.float32) * (rank + 1))
    h0 = dist.all_reduce(flat.g[:end], op=dist.ReduceOp.SUM, async_op=True)
    h1 = dist.all_reduce(flat.g[end:], op=dist.ReduceOp.SUM, async_op=True)
    h0.wait(); h1.wait()
    want = torch.arange(flat.g.numel(), dtype=torch.float32) * sum(r + 1 for r in range(world))
    q.put((rank, ok, end, bool(torch.equal(flat.g, want))))
    dist.destroy_process_group()


def test_head_tensor_range_for_ranged_optimizer_steps():
    """trainer.head_tensor_range: the leading tensors / chunks of the flat layout that lpm_adam_clip_step_range may update
    as soon as the head of the backward is done (per-tensor clipping, utils.py:181-188)."""
    from learnablepoolingmethods_b200.trainer import FlatState, head_tensor_range
    store = _cpu_store("WillowModelReg")
    tr = store.trainable()
    head = [n for n in tr if not n.startswith(("video_", "audio_", "input_bn")) and n != "hidden1_weights"]
    body = [n for n in tr if n.startswith(("video_", "audio_", "input_bn"))]
    flat = FlatState(store, head + body, {}, factored=("hidden1_weights",))
    nt, nc = head_tensor_range(flat)
    assert nt == len(head) and flat.order[:nt] == head
    assert nc == flat.chunk_begin_host[nt] and 0 < nc < flat.chunk_begin_host[-1]
    # the chunks of the range are exactly the chunks of its tensors (table rows carry absolute tensor ids)
    table = flat.table.cpu()
    assert int(table[:nc, 0].max()) == nt - 1 and int(table[nc:, 0].min()) == nt
    assert flat.chunk_begin.cpu().tolist() == flat.chunk_begin_host
    # a layout that does not start with the head's tensors has no such range
    store2 = _cpu_store("WillowModelReg")
    flat2 = FlatState(store2, body[:2] + head + body[2:], {}, factored=("hidden1_weights",))
    assert head_tensor_range(flat2) is None


def test_head_gradient_span_and_split_allreduce_world2_gloo():
    import torch.multiprocessing as mp
    from learnablepoolingmethods_b200.trainer import FlatState, head_gradient_span
    # layout rules on one process: the backward's order (head first) is splittable, hidden1_weights (factored) has no
    # gradient storage, and an interleaved order is rejected (one all-reduce at the end instead)
    store = _cpu_store("WillowModelReg")
    tr = store.trainable()
    head = [n for n in tr if not n.startswith(("video_", "audio_", "input_bn")) and n != "hidden1_weights"]
    body = [n for n in tr if n.startswith(("video_", "audio_", "input_bn"))]
    flat = FlatState(store, head + body, {}, factored=("hidden1_weights",))
    end, ok = head_gradient_span(flat)
    assert ok and end == flat.end_offset(head[-1]) and 0 < end < flat.g_total
    assert "hidden1_weights" not in flat.grad_views and flat.g.numel() == flat.g_total
    store2 = _cpu_store("WillowModelReg")
    flat2 = FlatState(store2, body[:2] + head + body[2:], {}, factored=("hidden1_weights",))
    assert head_gradient_span(flat2)[1] is False
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_split_allreduce, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert len({r[2] for r in res}) == 1                     # every rank splits at the same offset
    for rank, ok, end, equal in res:
        assert ok and equal, rank


def _worker_sharded_exchange(rank, world, port, q):
    """ShardedHiddenUpdate's exchange on gloo with CPU tensors: descriptor column slices by all-to-all, dLoss/dhidden by
    all-gather, the shard's tower-summed gradient as ONE product over world*B rows, the clip norm of the whole tensor by
    a scalar all-reduce, and the all-gather that makes every rank's copy of the updated rows complete again."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from learnablepoolingmethods_b200.trainer import pack_descriptor_slices, shard_row_range
    B, Kd, H = 3, 16 * world, 5
    towers = []
    for r in range(world):                                  # every rank can rebuild every tower's factors (the reference sum)
        g = torch.Generator().manual_seed(77 + r)
        towers.append((torch.randn(B, Kd, generator=g), torch.randn(B, H, generator=g)))
    vlad, dact = towers[rank]
    r0, r1 = shard_row_range(Kd, world, rank)
    send = pack_descriptor_slices(vlad, world)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    g_all = torch.empty(world * B, H)
    dist.all_gather_into_tensor(g_all, dact.contiguous())
    dw_shard = recv.view(world * B, r1 - r0).t() @ g_all                     # [rows_r, H]
    dense = sum(a.t() @ g for a, g in towers)                                # utils.py:205-211: SUM over towers
    ok = torch.allclose(dw_shard, dense[r0:r1], rtol=1e-5, atol=1e-5)
    sumsq = (dw_shard * dw_shard).sum().reshape(1)
    dist.all_reduce(sumsq, op=dist.ReduceOp.SUM)
    ok = ok and abs(float(sumsq.sqrt()) - float(dense.norm())) < 1e-4 * float(dense.norm())
    # shard-local update, then the gather of the updated rows (sync_master / the fp16 operand gather)
    w = torch.zeros(Kd, H)
    w[r0:r1] = dense[r0:r1] * 0.5
    dist.all_gather_into_tensor(w, w[r0:r1].clone())
    ok = ok and torch.allclose(w, dense * 0.5, rtol=1e-6, atol=1e-6)
    q.put((rank, bool(ok), (r0, r1)))
    dist.destroy_process_group()


def test_sharded_hidden_exchange_world2_gloo():
    import torch.multiprocessing as mp
    from learnablepoolingmethods_b200.trainer import pack_descriptor_slices, shard_row_range
    assert [shard_row_range(64, 4, r) for r in range(4)] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    v = torch.arange(2 * 8, dtype=torch.float32).view(2, 8)
    s = pack_descriptor_slices(v, 2)
    assert s.shape == (2, 2, 4) and torch.equal(s[1], v[:, 4:])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_sharded_exchange, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[2] for r in res) == [(0, 16), (16, 32)]
    for rank, ok, _ in res:
        assert ok, rank
