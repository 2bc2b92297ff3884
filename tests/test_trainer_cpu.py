"""Host logic of savqa_b200.train.EncoderTrainer on CPU (kernels replaced by tests/fake_ops.py): the flat-buffer layout,
the binding of the modules' weight packs / gradient sinks to it, the bf16 mirror kept by the optimizer, and the
data-parallel exchange (world size 2 over gloo).  Reference: the same model stepped by plain autograd + torch.optim.Adam
(what main_itp_ddp_tar_super_node.py:203-206, 363-366 does)."""
import copy
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fake_ops

HERE = os.path.dirname(os.path.abspath(__file__))


def _setup(seed=0, batch=4):
    from savqa_b200 import synthetic
    cfg = synthetic.TINY
    model = synthetic.build_model(cfg, vocab_rows=1200, seed=seed)
    b = synthetic.make_batch(cfg, batch, seed=5, vocab_rows=1200)
    return cfg, model, b


def _reference_steps(model, batches, lr, dec_mask=True):
    """Plain autograd through the (unbound) modules + dense torch.optim.Adam over every parameter that gets a gradient."""
    from savqa_b200 import AttModel_x3 as A
    opt = None
    losses = []
    for b in batches:
        model.zero_grad(set_to_none=True)
        logits = model.encoder_step(b["vis_fea"], b["vis_fea_mask"], b["q_ipt"], b["q_ipt_mask"], b["q_ipt_graph"], b["syb_ipt"],
                                    b["macro_node_mask"], b["macro_graph_ipt"], dec_mask)
        loss = A.answer_loss(*logits, b["answer"])
        loss.backward()
        if opt is None:
            opt = torch.optim.Adam([p for p in model.parameters() if p.grad is not None], lr=lr)
        opt.step()
        losses.append(float(loss))
    return losses


def _max_param_diff(m1, m2):
    worst = 0.0
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        if a.dtype.is_floating_point:
            worst = max(worst, float((a - b).abs().max()))
    return worst


@pytest.mark.parametrize("rowsparse,steps", [(True, 3), (False, 2)])
def test_bound_trainer_matches_autograd_adam(monkeypatch, rowsparse, steps):
    fake_ops.install(monkeypatch)
    from savqa_b200 import train
    cfg, model, b = _setup()
    ref = copy.deepcopy(model)
    lr = 1e-3
    # different batches per step: word rows come and go, so the row-sparse tables exercise the deferred (catch-up) Adam path
    from savqa_b200 import synthetic
    bs = [b] + [synthetic.make_batch(cfg, 4, seed=6 + i, vocab_rows=1200) for i in range(steps - 1)]
    ref_losses = _reference_steps(ref, bs, lr)
    tr = train.EncoderTrainer(model, lr=lr, rowsparse=rowsparse)
    losses = [float(tr.step(x)) for x in bs]
    tr.flush_tables()  # every word row up to date: the tables now hold what dense Adam holds
    # every attention / feed-forward / head pack that can be bound is bound, and its views alias the flat buffers
    att = model.att_vis_grid.enc_self_attention_0
    assert att._packs["qkv"].bound and att._packs["kv"].bound and att.normalization._sink.bound
    assert att._packs["qkv"].gw.data_ptr() == att.Q_proj[0].weight.grad.data_ptr()
    assert att._packs["kv"].w.data_ptr() == att._packs["qkv"].w[att.num_units:].data_ptr()
    assert model.att_vis_grid.enc_feed_forward_0._packs["w1"].bound
    assert not model.att_vis_grid._pk["mlp"].bound  # K = 300 is not a multiple of 8: stays on the staging path
    assert model.att_syb._pk["mlp2"].bound and model._pk["cls"][0].bound
    # the decoder's cross-attention layers keep [Wk_0; Wv_0; ...; Wk_5; Wv_5] as ONE block of the flat buffers (one K/V GEMM for
    # the six layers); every layer's own packs are slices of it, and its Wq / bq are bound on their own
    br = model.att_vis_grid
    C, L = br.hidden_size, br.num_blocks
    pall = br._pk["kv_all"]
    assert pall.bound and pall.w.shape == (2 * L * C, C) and pall.gb.shape == (2 * L * C,)
    for i in range(L):
        lay = getattr(br, "dec_vanilla_attention_%d" % i)
        assert lay._kv_external and lay._kv_index == i and lay._packs["q"].bound and not lay._packs["qkv"].bound
        assert lay._packs["kv"].w.data_ptr() == pall.w[2 * C * i:].data_ptr() and lay._packs["kv"].gw.data_ptr() == pall.gw[2 * C * i:].data_ptr()
        assert lay._packs["kv"].gw.data_ptr() == lay.K_proj[0].weight.grad.data_ptr()
        assert lay._packs["v"].gb.data_ptr() == lay.V_proj[0].bias.grad.data_ptr()
    assert losses == pytest.approx(ref_losses, rel=1e-5)
    # bf16 operands make the two runs differ only through rounding noise inside identical arithmetic: same stand-in kernels
    assert _max_param_diff(model, ref) < 2e-5
    # the mirror the GEMMs read is the bf16 image of the updated parameters
    assert torch.equal(tr.flat_bf16, tr.flat_param.to(torch.bfloat16))
    tr.release()
    assert not att._packs["qkv"].bound and not att.normalization._sink.bound


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_ops as F
    import savqa_b200.ops as real
    for name in dir(F):
        fn = getattr(F, name)
        if callable(fn) and hasattr(real, name) and not name.startswith("_") and name != "install":
            setattr(real, name, fn)
    from savqa_b200 import train
    torch.set_num_threads(2)  # same BLAS threading (= summation order) as the single-rank run it is compared with
    cfg, model, b = _setup(batch=4)
    shard = {k: v[rank::world].contiguous() for k, v in b.items()}
    tr = train.EncoderTrainer(model, lr=1e-3, eps=1e-2, rowsparse=True)
    tr.step(shard)
    # the bucketed reducer ran from inside the backward pass: 6 + 6 encoder blocks and the heads reported their buckets
    want = 1 + 2 * -(-6 // train.BUCKET_BLOCKS)
    assert tr.reducer is not None and len(tr.reducer.done) >= want, len(tr.reducer.done)
    if rank == 0:
        torch.save({k: v.clone() for k, v in model.state_dict().items()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_world2_gloo(monkeypatch, tmp_path):
    """Two ranks on half batches each == one rank on the whole batch (mean loss => averaged gradients; row-sparse word-table
    gradients exchanged by all-gather)."""
    out = str(tmp_path / "dp.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    fake_ops.install(monkeypatch)
    from savqa_b200 import train
    nthreads = torch.get_num_threads()
    torch.set_num_threads(2)
    cfg, model, b = _setup(batch=4)
    # eps well above the fp32 summation-order noise of the gradients: Adam's first step is lr * g / (|g| + eps), which would
    # otherwise turn a sign flip of a noise-level gradient into a 2 * lr difference
    tr = train.EncoderTrainer(model, lr=1e-3, eps=1e-2, rowsparse=True)
    tr.step(b)
    torch.set_num_threads(nthreads)
    dp = torch.load(out)
    worst = max(float((dp[k] - v).abs().max()) for k, v in model.state_dict().items() if v.dtype.is_floating_point)
    assert worst < 1e-6



def test_grad_reducer_launch_order_follows_the_backward_pass():
    """Under graph capture the reducer launches the bucket all-reduces in finish(), ordered by the stage of the backward pass that
    completes them -- heads, decoders, encoder blocks last to first, the two branches alternating -- not in the order autograd
    happened to report them (the whole visual branch first): collectives of one communicator run in launch order."""
    from savqa_b200.train import GradReducer
    a, b = 111, 222  # ids of the two branch models
    reported = [(a, "dec")] + [(a, "enc", i) for i in range(5, -1, -1)] + [(0, "heads")] + [(b, "dec")] + [(b, "enc", i) for i in range(5, -1, -1)]
    pending = [(GradReducer._stage(k), n, k) for n, k in enumerate(reported)]
    order = [k for _, _, k in sorted(pending, key=lambda t: (t[0], t[1]))]
    assert order[0] == (0, "heads") and order[1:3] == [(a, "dec"), (b, "dec")]
    assert order[3:] == [(br, "enc", i) for i in range(5, -1, -1) for br in (a, b)]


def test_deferred_row_adam_host_semantics(monkeypatch):
    """The deferred row-wise Adam protocol (catch-up before the read, update after the backward, flush) equals dense
    torch.optim.Adam -- here with the CPU stand-ins of the kernels; tests/test_gpu_kernels.py runs the same case on the B200."""
    import parity_cases as PC
    PC.deferred_adam_case(fake_ops, "cpu", steps=10)


@pytest.mark.parametrize("kind", ["full", "compact"])
def test_full_step_trainer_matches_autograd_adam(monkeypatch, kind):
    """The train script's whole step -- 16-argument AttModel.forward with MIL_NCE, loss + (-mil_nce_obj), backward, Adam
    (main_itp_ddp_tar_super_node.py:321-366) -- through the bound trainer (three row-sparse word tables, MIL_NCE's Linear layers in
    the flat buffers), from collate_fn's dense batch and from the compact hand-off: same losses / parameters as plain autograd over
    the unbound modules + dense torch.optim.Adam."""
    fake_ops.install(monkeypatch)
    from savqa_b200 import AttModel_x3 as A, collate, synthetic, train
    cfg, model, b0 = _setup()
    bs = [collate.expand_batch(collate.compact_batch(x)) for x in
          [b0] + [synthetic.make_batch(cfg, 4, seed=16 + i, vocab_rows=1200) for i in range(2)]]  # bf16-representable features
    ref = copy.deepcopy(model)
    lr = 1e-3
    opt, ref_losses = None, []
    for b in bs:
        ref.zero_grad(set_to_none=True)
        e = torch.empty((4, 0))
        lc, lv, ls, obj, _ = ref(b["vis_fea"], b["vis_fea_mask"], b["q_ipt"], b["q_ipt_mask"], b["q_ipt_graph"], b["macro_node_ipt"],
                                 b["macro_node_mask"], b["macro_graph_ipt"], b["macro_obj_loc_ipt"], b["micro_positive_obj_ipt"],
                                 b["micro_negative_obj_ipt"], b["micro_obj_mask"], e, e, e, e, decMask=True, mcb=False)
        loss = A.answer_loss(lc, lv, ls, b["answer"]) - obj
        loss.backward()
        if opt is None:
            # eps well above the rounding noise of the gradients: with the default 1e-8 a 1-ulp parameter difference (deferred vs
            # dense Adam differ in the order of a few fp32 operations) flips the sign-like update of noise-level gradients
            opt = torch.optim.Adam([p for p in ref.parameters() if p.grad is not None], lr=lr, eps=1e-2)
        opt.step()
        ref_losses.append(float(loss))
    tr = train.EncoderTrainer(model, lr=lr, eps=1e-2, rowsparse=True, step=kind)
    feed = bs if kind == "full" else [collate.compact_batch(x) for x in bs]
    losses = [float(tr.step(x)) for x in feed]
    tr.flush_tables()
    assert len(tr.tables) == 3 and model.MIL_NCE._pk["vis"].bound and model.MIL_NCE._pk["ipt"].bound
    assert losses == pytest.approx(ref_losses, rel=2e-5)
    assert _max_param_diff(model, ref) < 2e-5
    assert model.MIL_NCE.marco_mlp[0].weight.grad is None  # detached at AttModel_x3.py:354: never in the flat buffers
    tr.release()


def test_fused_decoder_matches_per_module_chain(monkeypatch):
    """functional.DecoderFn (the decoder as fused GEMM + LayerNorm launches, forward and hand-written backward) against the
    per-module chain (TokenSelfAttentionFn -> GraphAttentionFn -> FeedForwardFn per layer, itself pinned to the oracle): output,
    gradient of the start token, every decoder parameter gradient in the flat buffer, and d(memory) through the fused K/V block."""
    fake_ops.install(monkeypatch)
    from savqa_b200 import functional as Fn, ops, synthetic, train
    cfg, model, b = _setup(batch=5)
    tr = train.EncoderTrainer(model, lr=1e-3)
    tr.prepare(b)  # binds the packs / LayerNorm sinks to the flat buffers
    br = model.att_syb
    C, L, H = br.hidden_size, br.num_blocks, br.num_heads
    B, T = 5, 9
    g = torch.Generator().manual_seed(3)
    mem0 = torch.randn(B, T, C, generator=g)
    mem0[1, T - 2:] = 0  # padded memory tokens (key-masked)
    dec_mask = (torch.rand(B, 1, T, generator=g) < 0.8).float()
    dec_mask[:, 0, 0] = 1
    dout = torch.randn(B, 1, C, generator=g)
    x0 = (br.dec_emb.lookup_table[2] * (C ** 0.5) + br.dec_positional_encoding.lookup_table[0]).detach().reshape(1, 1, C).expand(B, 1, C).contiguous()
    layers = [(getattr(br, 'dec_self_attention_%d' % i), getattr(br, 'dec_vanilla_attention_%d' % i), getattr(br, 'dec_feed_forward_%d' % i))
              for i in range(L)]

    def run(fused):
        tr.flat_grad.zero_()
        x = x0.clone().requires_grad_(True)
        mem = mem0.clone().requires_grad_(True)
        if fused:
            pall = br._pk["kv_all"]
            on, mb = ops.row_nonzero(mem.detach().reshape(B * T, C))
            kv_all = torch.empty(B * T, 2 * L * C, dtype=torch.bfloat16)
            ops.gemm(mb, pall.w, B * T, 2 * L * C, C, bias=pall.bias, relu=True, out_bf16=kv_all)
            h = Fn.MemoryHolder()
            h.kv_all, h.pack_all, h.mem_bf16, h.shape = kv_all, pall, mb, mem.shape
            memj = Fn.MemoryJoinFn.apply(mem, h)
            memj._savqa_side = Fn.Side(memj, mb, on)
            y, _, _ = Fn.DecoderFn.apply(x, memj, dec_mask, dict(layers=layers, heads=H, kv_holder=h))
        else:
            y = x
            for sa, ca, ff in layers:
                y = ff(ca(sa(y, y, y), mem, mem, dec_mask))
        (y * dout).sum().backward()
        return y.detach(), x.grad.clone(), mem.grad.clone(), tr.flat_grad.clone()

    monkeypatch.setattr(Fn, "WGRAD_SIDE_STREAM", False)
    y1, dx1, dm1, fg1 = run(True)
    y0, dx0, dm0, fg0 = run(False)
    rel = lambda a_, b_: float((a_ - b_).norm() / (b_.norm() + 1e-30))  # noqa: E731
    assert rel(y1, y0) < 1e-5, rel(y1, y0)
    assert rel(dx1, dx0) < 2e-3 and rel(dm1, dm0) < 2e-3, (rel(dx1, dx0), rel(dm1, dm0))
    names = {id(p): k for k, p in model.named_parameters()}
    off = tr.views._off
    for p in tr.dense:
        k = names[id(p)]
        if k.startswith("att_syb.dec_") and not k.startswith(("att_syb.dec_emb", "att_syb.dec_pos")):
            o, n = off[id(p)]
            a_, b_ = fg1[o:o + n], fg0[o:o + n]
            if float(b_.abs().sum()) == 0.0:
                assert float(a_.abs().sum()) == 0.0, k
            else:
                assert rel(a_, b_) < 5e-3, (k, rel(a_, b_))
    tr.release()
