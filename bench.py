#!/usr/bin/env python
"""Benchmark of the B200-native graph-guided encoder path of SA-VQA (BASELINE.json metric:
"SA-VQA encoder train samples/s at 1/2/4/8 B200; attn % of tensor-core peak").

    python bench.py --gpus N --steps K --warmup W                 # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...       # the reference's CPU path (oracle port) on the host cores
    python bench.py --impl stock-gpu [--stock-dtype fp32|bf16]    # the restated reference on the SAME B200 with stock ATen / cuBLAS
    python bench.py --workload inference                          # BASELINE configs[1]: batch-256 inference, V=100, M=279
    python bench.py --dense-tables                                # word tables in the dense flat buffers (dense grads + dense Adam)

A "step" (default workload) is one full training step of the hot path over one GQA-shaped synthetic batch of 128 samples per
GPU (BASELINE.json configs[2]/[3]): both branch models (visual T=56, symbolic T=128), classifier heads, label-smoothed loss,
backward, gradient all-reduce (N > 1), Adam.  Prints ONE JSON line (rank 0).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "SA-VQA encoder train samples/s"
METRIC_INFER = "SA-VQA encoder inference samples/s"
UNIT = "samples/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def measured_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this round (profiles/r2_ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when no
    capture of the current kernel is committed -- never a literal."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        d = json.load(open(path))
        e = d.get(kernel_key)
        return (float(e["dram_bytes"]), e.get("source", path)) if e else (None, None)
    except Exception:
        return None, None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                          str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """Samples from here on count (the sampler is started early: nvidia-smi needs ~1 s before its first line)."""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        t0 = getattr(self, "t_mark", 0.0)
        for ts, r in self.rows:
            if ts < t0:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _time_alone(fn, flush, iters):
    import torch
    for _ in range(3):
        fn()
    us = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        us.append(e0.elapsed_time(e1) * 1e3)
    return sum(us) / len(us)


def dominant_kernel_roofline(batch, cfg, peaks, iters=10):
    """Live CUDA-event timing of the dominant kernel of the step -- the CTA-pair tensor-core GEMM (csrc/gemm2_tcgen05.cu,
    ~60 % of the step's kernel time, 98 % of its FLOPs) -- on the forward GEMM shapes of both branch models' encoder blocks
    (fused QKV projection, feedforward conv1 and conv2), each launched alone on the current stream with the L2 flushed
    between launches.  achieved = algorithmic 2*M*N*K FLOPs per launch / mean launch duration; peak = the measured BURST
    bf16 figure (a kernel timed alone)."""
    import torch
    from savqa_b200 import ops
    C, Hd = cfg["hidden"], 4 * cfg["hidden"]
    BF = torch.bfloat16
    ops.clear_stream_sm_limits()  # the training leg left this stream with the visual branch's share of the SMs (56 of 148)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    total_flops = total_us = 0.0
    n_launch = 0
    shapes = []
    for T in (cfg["V"] + cfg["Q"], cfg["M"] + cfg["Q"]):
        M = batch * T
        x = torch.randn(M, Hd, device="cuda").to(BF)
        for name, N, K, fuse_res in (("qkv", 3 * C, C, False), ("conv1", Hd, C, False), ("conv2", C, Hd, True)):
            w = torch.randn(N, K, device="cuda").to(BF)
            bias = torch.randn(N, device="cuda")
            if fuse_res:
                res, out = torch.randn(M, N, device="cuda"), torch.empty(M, N, device="cuda")
                fn = lambda: ops.gemm(x[:, :K], w, M, N, K, bias=bias, res=res, out_f32=out)  # noqa: E731
            else:
                out = torch.empty(M, N, device="cuda", dtype=BF)
                fn = lambda: ops.gemm(x[:, :K], w, M, N, K, bias=bias, relu=True, out_bf16=out)  # noqa: E731
            mean_us = _time_alone(fn, flush, iters)
            total_flops += 2.0 * M * N * K
            total_us += mean_us
            n_launch += 1
            shapes.append(f"{name} M={M} N={N} K={K}: {mean_us:.1f} us")
    achieved = total_flops / total_us / 1e6  # TFLOP/s
    traffic, src = measured_traffic("gemm2_conv1_M16384")
    return {"bound": "tensor", "kernel": "gemm2_bf16_kernel (tcgen05 cta_group::2, TMA, TMEM)", "achieved": achieved, "peak": peaks["tf_burst"],
            "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"], "traffic": traffic,
            "note": f"FLOP-weighted over {n_launch} forward GEMM shapes of the encoder blocks, each timed alone with CUDA events on the "
                    f"launching stream, L2 flushed between launches; peak = {peaks['source']} burst bf16 figure; traffic = "
                    + (f"ncu dram bytes of one conv1 M=16384 launch ({src})" if traffic is not None else "null (no ncu capture of the current kernel committed)"),
            "shapes": shapes}


def attention_roofline(batch, cfg, peaks, iters=10):
    """BASELINE.json's second figure, "attn % of tensor-core peak": the fused graph-masked attention core (csrc/attn_tcgen05.cu,
    attn_bwd_tcgen05.cu) timed alone with CUDA events at the step's two shapes (T = V+Q and M+Q keys per sample, 8 heads, d = 64,
    bit-packed graph, forward statistics reused by the backward), L2 flushed between launches.  Dense-equivalent FLOPs
    (4 T^2 d forward, 10 T^2 d backward per sample and head) against the measured burst bf16 peak, and algorithmic bytes against
    the measured HBM peak -- at these sizes (AI 28..130 flop/B, ridge 212) the core is memory/latency bound, so the tensor
    fraction is small by construction."""
    import torch
    from savqa_b200 import ops
    C, H = cfg["hidden"], cfg["heads"]
    d = C // H
    BF = torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    tot_flops = tot_us = 0.0
    for T in (cfg["V"] + cfg["Q"], cfg["M"] + cfg["Q"]):
        N, M = batch, batch * T
        qkv = torch.randn(M, 3 * C, device="cuda").relu().to(BF)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        graph = (torch.rand(N, T, T, device="cuda") < 0.3).float()
        graph[:, torch.arange(T), torch.arange(T)] = 1
        on = torch.ones(M, device="cuda")
        bits = ops.pack_graph_bits(graph)
        stats = torch.empty(H * N * T * 4, device="cuda")
        dout = torch.randn(M, C, device="cuda")
        dqkv = torch.empty(M, 3 * C, device="cuda", dtype=BF)
        fwd = lambda: ops.graph_attention_fwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, False, 0, graph_bits=bits, stats=stats)  # noqa: E731
        o, _ = fwd()
        bwd = lambda: ops.graph_attention_bwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, dout, dqkv[:, :C], dqkv[:, C:2 * C],  # noqa: E731
                                              dqkv[:, 2 * C:], graph_bits=bits, stats=stats, fwd_out=o)
        for name, fn, flops, nbytes in (("fwd", fwd, 4.0 * N * H * T * T * d, M * 3 * C * 2 + N * T * T / 8 + M * C * 4),
                                        ("bwd", bwd, 10.0 * N * H * T * T * d, M * 3 * C * 2 * 2 + N * T * T / 8 + 2 * M * C * 4)):
            mean_us = _time_alone(fn, flush, iters)
            tot_flops += flops
            tot_us += mean_us
            out[f"{name}_T{T}"] = {"us": round(mean_us, 1), "tflops": round(flops / mean_us / 1e6, 1),
                                   "pct_tensor_peak": round(100 * flops / mean_us / 1e6 / peaks["tf_burst"], 2),
                                   "gbs": round(nbytes / mean_us / 1e3, 1), "pct_hbm_peak": round(100 * nbytes / mean_us / 1e3 / peaks["hbm_gbs"], 1)}
    out["pct_tensor_peak"] = round(100 * tot_flops / tot_us / 1e6 / peaks["tf_burst"], 2)
    out["note"] = ("fused adjacency-masked attention core, forward and backward at the step's two shapes, each launch timed alone (CUDA events, "
                   "L2 flushed); dense-equivalent FLOPs vs the measured burst bf16 peak; HBM/latency bound at these sizes")
    return out


def hbm_kernel_rooflines(batch, cfg, peaks, iters=10):
    """north_star: "achieved HBM GB/s for LayerNorm and gather".  Residual+LayerNorm forward / backward at the symbolic branch's
    row count and the word-row gather, each timed alone (CUDA events, L2 flushed), algorithmic bytes over the
    measured HBM peak; the ncu dram counters of the same launches are under profiles/ (r2_ncu_hbm_kernels.txt)."""
    import torch
    from savqa_b200 import ops
    C = cfg["hidden"]
    rows = batch * (cfg["M"] + cfg["Q"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    x, res = torch.randn(rows, C, device="cuda"), torch.randn(rows, C, device="cuda")
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dy = torch.randn(rows, C, device="cuda")
    out = {}
    us = _time_alone(lambda: ops.layernorm_fwd(x, res, gamma, beta, 1e-8, True, True, True), flush, iters)
    nbytes = rows * C * (4 * 4 + 2)  # read x, res; write pre, y (fp32) + y (bf16)
    out["res_ln_fwd"] = {"rows": rows, "us": round(us, 1), "gbs": round(nbytes / us / 1e3, 1), "frac_hbm_peak": round(nbytes / us / 1e3 / peaks["hbm_gbs"], 3)}
    _, pre, _, _ = ops.layernorm_fwd(x, res, gamma, beta, 1e-8, True, False, False)
    us = _time_alone(lambda: ops.layernorm_bwd(dy, pre, gamma, 1e-8, dg, db, want_bf16=True), flush, iters)
    nbytes = rows * C * (3 * 4 + 2)  # read dy, pre; write dx (fp32) + dx (bf16)
    out["ln_bwd"] = {"rows": rows, "us": round(us, 1), "gbs": round(nbytes / us / 1e3, 1), "frac_hbm_peak": round(nbytes / us / 1e3 / peaks["hbm_gbs"], 3)}
    table = torch.randn(407000, 300, device="cuda")
    n = 1 << 18  # a large gather: the step's own gathers (2560 rows) are launch-latency, not bandwidth
    idx = torch.randint(0, 407000, (n,), device="cuda")
    us = _time_alone(lambda: ops.gather_rows(table, idx, want_f32=True, want_bf16=True), flush, iters)
    nbytes = n * (300 * 4 * 2 + 304 * 2 + 8)
    out["gather_rows"] = {"rows": n, "us": round(us, 1), "gbs": round(nbytes / us / 1e3, 1), "frac_hbm_peak": round(nbytes / us / 1e3 / peaks["hbm_gbs"], 3)}
    out["note"] = "each kernel timed alone (CUDA events, L2 flushed); algorithmic bytes over the measured HBM copy bandwidth"
    return out


def cpu_reference_run(cfg, steps, warmup, batch_size, threads=None, optimizer=True):
    """The reference's own CPU implementation of the path = the oracle port (the reference is Python and cannot be shipped to the
    box; oracle/savqa_oracle.py restates it op for op and is pinned to it by tests/golden): forward, backward and -- like our
    step -- torch.optim.Adam over the parameters that receive gradients."""
    import torch
    from oracle import savqa_oracle as O
    from savqa_b200 import synthetic
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    model = synthetic.build_model(cfg, vocab_rows=20000)  # gather / row-update cost is row-count independent; keeps host RAM small
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and not k.startswith(("mcb.", "cls_mcb.")))
              for k, v in model.state_dict().items()}
    batch = synthetic.make_batch(cfg, batch_size, seed=0, vocab_rows=20000)
    times = []
    opt = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in params.values():
            p.grad = None
        loss, _, _, _ = O.full_step(params, batch, cfg["blocks"], cfg["heads"])  # 16-argument forward incl. MIL_NCE + loss (main...:321-360)
        loss.backward()
        if optimizer:
            if opt is None:
                opt = torch.optim.Adam([p for p in params.values() if p.grad is not None], lr=1e-4)
            opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch_size / sec, sec, threads


def stock_gpu_run(cfg, batch_size, steps, warmup, mode, dev, vocab_rows=None):
    """SURVEY.md 8(d), last row: the restated reference path on the SAME B200 with stock PyTorch ops (ATen / cuBLAS), fp32 or
    bf16 autocast, dense word-table gradients and torch.optim.Adam over every parameter that receives a gradient -- what the
    reference does.  (The restatement builds the masks with vectorised ops: it does not pay the reference's per-sample Python
    loop with a host sync per sample.)  Returns (samples/s, seconds per step)."""
    import torch
    from oracle import savqa_oracle as O
    from savqa_b200 import synthetic
    model = synthetic.build_model(cfg, vocab_rows=vocab_rows)
    params = {k: v.detach().to(dev).requires_grad_(v.dtype.is_floating_point and not k.startswith(("mcb.", "cls_mcb.")))
              for k, v in model.state_dict().items()}
    del model
    batch = {k: v.to(dev) for k, v in synthetic.make_batch(cfg, batch_size, seed=0, vocab_rows=vocab_rows).items()}
    opt = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize()
            ev[0].record()
        for p in params.values():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            loss, _, _, _ = O.full_step(params, batch, cfg["blocks"], cfg["heads"])
        loss.backward()
        if opt is None:
            opt = torch.optim.Adam([p for p in params.values() if p.grad is not None], lr=1e-4)
        opt.step()
    ev[1].record()
    torch.cuda.synchronize()
    sec = ev[0].elapsed_time(ev[1]) / 1e3 / steps
    del params, opt
    gc.collect()
    torch.cuda.empty_cache()
    return batch_size / sec, sec


def train_config(args, world):
    what = {"compact": "the train script's WHOLE step: 16-argument AttModel.forward (MIL_NCE -> symbolic node features, both branch models, "
                       "heads), loss + MIL-NCE term, backward, Adam; inputs = the loader's compact hand-off (bf16 region features, lengths, "
                       "bit-packed adjacency)",
            "full": "the train script's WHOLE step (16-argument AttModel.forward incl. MIL_NCE, loss + MIL-NCE term, backward, Adam); inputs = "
                    "collate_fn's dense fp32 / int32 batch",
            "encoder": "encoder step with MIL_NCE's output `syb_ipt` [B,M,2048] fp32 shipped from the host (round-1 workload)"}[args.step]
    return {"workload": f"configs[2]/[3]: AttModel_x3 training step -- {what}; GQA-shaped synthetic batch, V=36 regions + Q=20 tokens (T=56) "
                        "visual branch, M=108 nodes + Q=20 (T=128) symbolic branch, hidden 512, 8 heads, 6+6 blocks, hidden_size_mil 64, "
                        f"topN 1, 1845 classes, decMask=True, dropout {args.dropout:g}",
            "step": args.step,
            "per_gpu_batch": args.batch, "global_batch": args.batch * max(world, 1), "parallelism": f"dp{max(world, 1)}",
            "l2": "activations per step (~2 GB) exceed the 126 MB L2; no explicit flush",
            "word_tables": "dense flat buffers (dense gradients + dense Adam)" if args.dense_tables else
                           "row-sparse gradients + deferred row-wise Adam (== dense torch.optim.Adam after flush)"}


def run_training(args, rank, world, local_rank, dev, peaks, sampler, dense_tables, steps, warmup, extras=True):
    """Device-timed and end-to-end training throughput of one configuration; returns a dict of measurements."""
    import torch
    import torch.distributed as dist
    from savqa_b200 import synthetic, train
    cfg = synthetic.GQA_SHAPED
    model = synthetic.build_model(cfg, seed=0, dropout=args.dropout).to(dev)
    model.train()
    from savqa_b200 import collate
    keys = {"encoder": train.STEP_KEYS, "full": train.FULL_KEYS, "compact": train.COMPACT_KEYS}[args.step]
    host_batches = []
    for i in range(2):
        b = synthetic.make_batch(cfg, args.batch, seed=100 + rank * 16 + i)
        if args.step == "compact":
            b = collate.compact_batch(b)  # what a collate_fn replacement emits
        host_batches.append({k: b[k].contiguous().pin_memory() for k in keys})
    dev_batch = {k: v.to(dev) for k, v in host_batches[0].items()}
    trainer = train.EncoderTrainer(model, lr=1e-4, rowsparse=not dense_tables, step=args.step)
    trainer.prepare(dev_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.mode == "graph":
        trainer.capture(dev_batch, warmup=2)
        run_step = lambda: trainer.replay()  # noqa: E731
    else:
        run_step = lambda: trainer.step(dev_batch)  # noqa: E731
        trainer.step(dev_batch)
    launches_per_step = trainer.launches_per_step

    # ---- device-resident timing: `value` ----
    for _ in range(max(warmup, 3)):
        run_step()
    barrier()
    if sampler is not None:
        sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        loss = run_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    loss_val = float(loss)

    # ---- end to end through the public API with HOST (pinned) inputs: `e2e` ----
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())
    loss_host = torch.zeros(1).pin_memory()
    if args.mode == "graph":
        trainer.stage(host_batches[0])

        def e2e_step(i):
            # every step: its inputs have travelled host -> staging on the copy stream while the previous step computed; they
            # move into the graph's static buffers, the NEXT step's host batch starts travelling, the step runs, and the loss
            # comes back to the host
            trainer.commit()
            trainer.stage(host_batches[(i + 1) & 1])
            l = trainer.replay()
            loss_host.copy_(l.reshape(1), non_blocking=True)
    else:
        def e2e_step(i):
            b = {k: v.to(dev, non_blocking=True) for k, v in host_batches[i & 1].items()}
            l = trainer.step(b)
            loss_host.copy_(l.reshape(1), non_blocking=True)
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        e2e_step(i)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)

    # ---- ranks agree: same parameters on every rank after the same number of steps (bit for bit) ----
    ranks = None
    if world > 1:
        tt = torch.tensor([ms, e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])
        trainer.flush_tables()  # outside the timed region: every table row current, so the tables can be compared as a whole
        tabs = [t.weight.data.view(torch.int32).sum(dtype=torch.int64).double() for t in trainer.tables]  # exact checksum of the bits
        chk = torch.stack([trainer.flat_param.double().sum(), trainer.flat_param.double().abs().sum()] + tabs +
                          [torch.tensor(loss_val, device=dev, dtype=torch.float64)])
        got = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(got, chk)
        ranks = {"loss": [float(g[-1]) for g in got], "param_checksum": [float(g[0]) for g in got],
                 "table_checksums": [[float(x) for x in g[2:-1]] for g in got],
                 "param_checksums_equal": all(bool(torch.equal(got[0][:-1], g[:-1])) for g in got)}
        assert ranks["param_checksums_equal"], f"ranks diverged: {ranks}"
    res = dict(ms=ms, e2e_ms=e2e_ms, loss=loss_val, h2d=h2d, launches_per_step=launches_per_step, ranks=ranks)
    trainer.graph = None
    trainer.release()
    del trainer, model
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_inference(args, rank, world, dev, peaks, steps, warmup):
    """BASELINE configs[1]: AttModel_x3 inference, batch 256 per GPU, 100 region nodes (T=120) and 279 symbolic nodes (T=299),
    batch-sharded with no collective.  Device-timed (inputs resident) and end to end (pinned host inputs, logits read back)."""
    import torch
    import torch.distributed as dist
    from savqa_b200 import infer, synthetic
    cfg = synthetic.CFG2
    B = args.infer_batch
    model = synthetic.build_model(cfg, seed=0).to(dev).eval()
    from savqa_b200 import collate
    host = []
    for i in range(2):
        b = collate.compact_batch(synthetic.make_batch(cfg, B, seed=300 + rank * 16 + i))
        host.append({k: b[k].contiguous().pin_memory() for k in infer.COMPACT_KEYS})
    dev_batch = {k: v.to(dev) for k, v in host[0].items()}
    runner = infer.InferenceRunner(model, dec_mask=True, full=True, keys=infer.COMPACT_KEYS)
    runner.capture(dev_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        runner.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = runner.replay()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    logits_host = torch.zeros(B, cfg["ncls"]).pin_memory()
    runner.stage(host[0])

    def e2e(i):
        runner.commit()
        runner.stage(host[(i + 1) & 1])
        o = runner.replay()
        logits_host.copy_(o[0], non_blocking=True)
    for i in range(3):
        e2e(i)
    barrier()
    e0.record()
    for i in range(steps):
        e2e(i)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms, e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    flops = synthetic.step_flops(cfg, B, backward=False, full=True)
    res = dict(ms=ms, e2e_ms=e2e_ms, h2d=h2d, d2h=B * cfg["ncls"] * 4, flops=flops, launches=runner.launches_per_batch, batch=B,
               finite=bool(torch.isfinite(out[0]).all()))
    runner.graph = None
    del runner, model
    gc.collect()
    torch.cuda.empty_cache()
    return res


def leave(world):
    """Tear the process group down properly (the CUDA graphs that captured its collectives are gone by now); a watchdog ends
    the process if NCCL's teardown still blocks."""
    import torch
    import torch.distributed as dist
    if world <= 1:
        return
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Thread(target=lambda: (time.sleep(30), os._exit(0)), daemon=True).start()
    try:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    except Exception:
        pass
    os._exit(0)


_REAL_STDOUT = None


def _own_stdout() -> None:
    """stdout carries exactly ONE line, the JSON result: everything libraries write to file descriptor 1 (NCCL prints its version
    banner there) goes to stderr instead; emit() writes the result to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _own_stdout()
    if int(os.environ.get("WORLD_SIZE", "1") or 1) > 1:  # see savqa_b200/__init__.py; here too, ahead of any CUDA call of this process
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "stock-gpu"])
    ap.add_argument("--workload", default="train", choices=["train", "inference"])
    ap.add_argument("--batch", type=int, default=128, help="samples per GPU (training)")
    ap.add_argument("--infer-batch", type=int, default=256, help="samples per GPU (inference, BASELINE configs[1])")
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"])
    ap.add_argument("--step", default="compact", choices=["compact", "full", "encoder"],
                    help="compact: whole step (MIL_NCE included) from the compact loader hand-off; full: same from collate_fn's dense batch; "
                         "encoder: round-1 workload (syb_ipt shipped from the host)")
    ap.add_argument("--dropout", type=float, default=0.0, help="dropout_rate of the model (the launcher's production value is 0.5, submit.py:72-104)")
    ap.add_argument("--dense-tables", action="store_true", help="word tables in the dense flat buffers: dense gradients + dense Adam")
    ap.add_argument("--stock-dtype", default="both", choices=["fp32", "bf16", "both"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline measurement (no comparators / secondary lines)")
    ap.add_argument("--cpu-sample", type=int, default=128, help="samples per step of the bounded CPU-baseline run (one GQA-shaped batch)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    from savqa_b200 import synthetic
    cfg = synthetic.GQA_SHAPED
    config = train_config(args, world)

    # ------------------------------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        n = args.cpu_sample
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        if (steps + warmup) * 3.0 > 240:  # ~3 s per 128-sample step on the box's host cores: keep the whole run within minutes
            n = max(8, int(n * 240 / ((steps + warmup) * 3.0)) // 8 * 8)
        value, sec, threads = cpu_reference_run(cfg, steps, warmup, n)
        cfg_line = dict(config, per_gpu_batch=0, global_batch=n, parallelism="cpu (one process, all host threads)", word_tables="dense (20000-row tables)")
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 0, "requested_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg_line,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{n}-sample GQA-shaped batch per step, full fwd+bwd of the encoder step + torch.optim.Adam "
                                           "(oracle port, torch CPU fp32, 20000-row word tables)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch.distributed as dist
    from savqa_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peaks = load_peaks()

    # ------------------------------------------------------------------------------------------ stock PyTorch on the same GPU
    if args.impl == "stock-gpu":
        if rank != 0:
            return
        modes = ["fp32", "bf16"] if args.stock_dtype == "both" else [args.stock_dtype]
        res = {}
        for m in modes:
            v, sec = stock_gpu_run(cfg, args.batch, max(1, args.steps), max(1, args.warmup), m, dev)
            res[m] = {"value": v, "ms_per_step": sec * 1e3}
        best = max(res.values(), key=lambda r: r["value"])
        line = {"impl": "stock-gpu", "metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 / bf16-autocast", "data": "synthetic", "config": dict(config, word_tables="dense (reference semantics)"),
                "modes": res, "note": "restated reference path (oracle/savqa_oracle.py) on CUDA tensors: stock ATen / cuBLAS kernels, dense "
                                      "407000-row table gradients, torch.optim.Adam"}
        emit(line)
        return

    # ------------------------------------------------------------------------------------------ our arm (B200)
    _lib.require_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # early: nvidia-smi needs a second or two before its first line; only samples after mark() count

    if args.workload == "inference":
        r = run_inference(args, rank, world, dev, peaks, args.steps, args.warmup)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            ms_per = r["ms"] / args.steps
            achieved = r["flops"] / (ms_per / 1e3) / 1e12
            line = {"metric": METRIC_INFER, "value": r["batch"] * world * args.steps / (r["ms"] / 1e3), "unit": UNIT, "n_gpus": world,
                    "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": "configs[1]: AttModel_x3 inference (16-argument forward: MIL_NCE + both branch models + heads, no grad; compact "
                                           "loader hand-off), batch 256 per GPU, V=100 regions "
                                           "(T=120), M=279 symbolic nodes (T=299), 2048-d features, hidden 512, 8 heads, 6+6 blocks, decMask=True",
                               "per_gpu_batch": r["batch"], "global_batch": r["batch"] * world, "parallelism": f"batch-sharded x{world}, no collective",
                               "l2": "inputs per batch (~0.5 GB) exceed the 126 MB L2; no explicit flush"},
                    "clocks": clocks, "outputs_finite": r["finite"],
                    "e2e": {"value": r["batch"] * world * args.steps / (r["e2e_ms"] / 1e3), "unit": UNIT, "h2d_bytes_per_step": r["h2d"],
                            "d2h_bytes_per_step": r["d2h"]},
                    "gpu_launches": r["launches"] * args.steps,
                    "step_roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                                      "frac": achieved / peaks["tf_sustained"],
                                      "note": f"algorithmic dense-equivalent forward FLOPs {r['flops'] / 1e12:.3f} TFLOP/batch/GPU over the CUDA-event time"}}
            emit(line)
        leave(world)
        return

    r = run_training(args, rank, world, local_rank, dev, peaks, sampler if rank == 0 else None, args.dense_tables, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = r["ms"] / args.steps
    value = args.batch * world * args.steps / (r["ms"] / 1e3)
    e2e_value = args.batch * world * args.steps / (r["e2e_ms"] / 1e3)
    extras = not args.no_extras and world == 1  # comparators and secondary lines: single-GPU run only (N > 1 ranks would idle in NCCL)
    line = None
    if rank == 0:
        flops = synthetic.step_flops(cfg, args.batch, backward=True, full=args.step != "encoder")
        achieved = flops / (ms_per_step / 1e3) / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": config, "mode": args.mode, "loss": r["loss"],
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4},
                "gpu_launches": r["launches_per_step"] * args.steps,
                "step_roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                                  "frac": achieved / peaks["tf_sustained"],
                                  "note": f"whole step: algorithmic dense-equivalent FLOPs {flops / 1e12:.3f} TFLOP/step/GPU over the "
                                          f"CUDA-event step time, vs {peaks['source']} sustained bf16 peak"}}
        if r["ranks"] is not None:
            line["ranks"] = r["ranks"]
        line["roofline"] = dominant_kernel_roofline(args.batch, cfg, peaks)
        line["attn_roofline"] = attention_roofline(args.batch, cfg, peaks)
        if extras:
            line["hbm_kernels"] = hbm_kernel_rooflines(args.batch, cfg, peaks)
    if extras:
        # the same step with the word tables in the dense flat buffers (the reference's literal optimizer layout), beside the default
        short = max(5, min(args.steps, 10))
        other = run_training(args, rank, world, local_rank, dev, peaks, None, not args.dense_tables, short, 3)
        line["dense_tables" if not args.dense_tables else "rowsparse_tables"] = {
            "ms_per_step": other["ms"] / short, "value": args.batch * short / (other["ms"] / 1e3), "unit": UNIT, "steps": short, "loss": other["loss"]}
        inf = run_inference(args, rank, world, dev, peaks, short, 3)
        line["inference"] = {"metric": METRIC_INFER, "workload": "configs[1]: batch 256, V=100 (T=120), M=279 (T=299), no grad",
                             "value": inf["batch"] * short / (inf["ms"] / 1e3), "unit": UNIT, "ms_per_batch": inf["ms"] / short,
                             "e2e_value": inf["batch"] * short / (inf["e2e_ms"] / 1e3), "h2d_bytes_per_batch": inf["h2d"],
                             "frac_of_sustained_peak": inf["flops"] / (inf["ms"] / short / 1e3) / 1e12 / peaks["tf_sustained"]}
        try:
            sg = {}
            for m in ("fp32", "bf16"):
                v, sec = stock_gpu_run(cfg, args.batch, 4, 2, m, dev)
                sg[m] = {"value": v, "ms_per_step": sec * 1e3}
            line["stock_gpu_baseline"] = dict(sg, unit=UNIT, note="restated reference on this B200 with stock ATen / cuBLAS ops, dense table "
                                                                   "gradients, torch.optim.Adam (python bench.py --impl stock-gpu)")
        except Exception as e:  # the headline number must not die with a comparator
            line["stock_gpu_baseline"] = {"value": None, "note": f"failed: {e}"}
        if not args.no_cpu_baseline:
            try:
                v, sec, threads = cpu_reference_run(cfg, 3, 1, args.cpu_sample)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                        "sample": f"{args.cpu_sample}-sample GQA-shaped batch (the GPU step's batch), mean of 3 fwd+bwd+Adam encoder "
                                                  f"steps (16-argument forward incl. MIL_NCE) after one warm-up (oracle port, torch CPU fp32, {sec:.2f} s per step)"}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        emit(line)
    leave(world)


if __name__ == "__main__":
    main()
