#!/usr/bin/env python
"""Benchmark of the B200-native graph-guided encoder path of SA-VQA (BASELINE.json metric:
"SA-VQA encoder train samples/s at 1/2/4/8 B200; attn % of tensor-core peak").

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A "step" is one full training step of the hot path over one GQA-shaped synthetic batch of 128 samples per GPU
(BASELINE.json configs[2]/[3]): bf16 weight staging, both branch models (visual T=56, symbolic T=128), classifier
heads, label-smoothed loss, backward, gradient all-reduce (N > 1), Adam.  Prints ONE JSON line (rank 0).
"""
import argparse
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
UNIT = "samples/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


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


def dominant_kernel_roofline(batch, cfg, peaks, iters=10):
    """Live CUDA-event timing of the dominant kernel of the step -- the CTA-pair tensor-core GEMM (csrc/gemm2_tcgen05.cu,
    ~60 % of the step's kernel time, 98 % of its FLOPs) -- on the forward GEMM shapes of both branch models' encoder blocks
    (fused QKV projection, feedforward conv1 and conv2), each launched alone on the current stream with the L2 flushed
    between launches.  achieved = algorithmic 2*M*N*K FLOPs per launch / mean launch duration; peak = the measured BURST
    bf16 figure (a kernel timed alone).  `traffic` is the ncu dram__bytes_read+write of the conv1 launch (profiles/r1_07_ncu_gemm2_conv1.txt)."""
    import torch
    from savqa_b200 import ops
    C, Hd = cfg["hidden"], 4 * cfg["hidden"]
    BF = torch.bfloat16
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
            mean_us = sum(us) / len(us)
            total_flops += 2.0 * M * N * K
            total_us += mean_us
            n_launch += 1
            shapes.append(f"{name} M={M} N={N} K={K}: {mean_us:.1f} us")
    achieved = total_flops / total_us / 1e6  # TFLOP/s
    return {"bound": "tensor", "kernel": "gemm2_bf16_kernel (tcgen05 cta_group::2, TMA, TMEM)", "achieved": achieved, "peak": peaks["tf_burst"],
            "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"], "traffic": 29.9e6,
            "note": f"FLOP-weighted over {n_launch} forward GEMM shapes of the encoder blocks, each timed alone with CUDA events on the "
                    f"launching stream, L2 flushed between launches; peak = {peaks['source']} burst bf16 figure; traffic = ncu "
                    "dram bytes of the conv1 M=16384 launch (profiles/r1_07_ncu_gemm2_conv1.txt: 18.9 MB read + 11.0 MB written while ncu counts; the 67 MB bf16 output "
                    "is still in the 126 MB L2 when the kernel ends; algorithmic bytes 86 MB)",
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
            mean_us = sum(us) / len(us)
            tot_flops += flops
            tot_us += mean_us
            out[f"{name}_T{T}"] = {"us": round(mean_us, 1), "tflops": round(flops / mean_us / 1e6, 1),
                                   "pct_tensor_peak": round(100 * flops / mean_us / 1e6 / peaks["tf_burst"], 2),
                                   "gbs": round(nbytes / mean_us / 1e3, 1), "pct_hbm_peak": round(100 * nbytes / mean_us / 1e3 / peaks["hbm_gbs"], 1)}
    out["pct_tensor_peak"] = round(100 * tot_flops / tot_us / 1e6 / peaks["tf_burst"], 2)
    out["note"] = ("fused adjacency-masked attention core, forward and backward at the step's two shapes, each launch timed alone (CUDA events, "
                   "L2 flushed); dense-equivalent FLOPs vs the measured burst bf16 peak; HBM/latency bound at these sizes")
    return out


def cpu_reference_run(cfg, steps, warmup, batch_size, threads=None):
    """The reference's own CPU implementation of the path = the oracle port (the reference is Python and cannot be
    shipped to the box; oracle/savqa_oracle.py restates it op for op and is pinned to it by tests/golden)."""
    import torch
    from oracle import savqa_oracle as O
    from savqa_b200 import synthetic
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    model = synthetic.build_model(cfg, vocab_rows=20000)  # gather cost is row-count independent; keeps host RAM small
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and not k.startswith(("mcb.", "MIL_NCE.", "cls_mcb.")))
              for k, v in model.state_dict().items()}
    batch = synthetic.make_batch(cfg, batch_size, seed=0, vocab_rows=20000)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in params.values():
            p.grad = None
        loss, _, _, _ = O.encoder_step(params, batch, cfg["blocks"], cfg["heads"])
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch_size / sec, sec, threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="samples per GPU")
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"])
    ap.add_argument("--dense-tables", action="store_true", help="reference-faithful dense word-table gradients + dense Adam")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=128, help="samples per step of the bounded CPU-baseline run (one GQA-shaped batch)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    from savqa_b200 import synthetic
    cfg = synthetic.GQA_SHAPED
    config = {"workload": "configs[2]/[3]: AttModel_x3 encoder training step (fwd+bwd+classifier heads+loss+Adam), GQA-shaped synthetic "
                          "batch, V=36 regions + Q=20 tokens (T=56) visual branch, M=108 nodes + Q=20 (T=128) symbolic branch, hidden 512, "
                          "8 heads, 6+6 blocks, 1845 classes, decMask=True, dropout 0",
              "per_gpu_batch": args.batch, "global_batch": args.batch * max(world, 1), "parallelism": f"dp{max(world, 1)}",
              "l2": "activations per step (~2 GB) exceed the 126 MB L2; no explicit flush",
              "word_tables": "dense" if args.dense_tables else "row-sparse gradients + row-wise Adam"}

    # ------------------------------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        n = args.cpu_sample
        steps = max(1, min(args.steps, 5))
        value, sec, threads = cpu_reference_run(cfg, steps, min(args.warmup, 1), n)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{n}-sample GQA-shaped batch, full fwd+bwd of the encoder step (oracle port, torch CPU fp32)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------ our arm (B200)
    import torch.distributed as dist
    from savqa_b200 import _lib, train
    torch.cuda.set_device(local_rank)
    _lib.require_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    peaks = load_peaks()

    model = synthetic.build_model(cfg, seed=0).to(dev)
    model.train()
    host_batches = [synthetic.make_batch(cfg, args.batch, seed=100 + rank * 16 + i, pin=True) for i in range(2)]
    host_batches = [{k: b[k] for k in train.STEP_KEYS} for b in host_batches]
    dev_batch = {k: v.to(dev) for k, v in host_batches[0].items()}
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # early: nvidia-smi needs a second or two before its first line; only samples after mark() count
    trainer = train.EncoderTrainer(model, lr=1e-4, rowsparse=not args.dense_tables)
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
    trainer.step(dev_batch) if args.mode == "eager" else None
    launches_per_step = trainer.launches_per_step

    # ---- device-resident timing: `value` ----
    for _ in range(max(args.warmup, 3)):
        run_step()
    barrier()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss = run_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    loss_val = float(loss)

    # ---- end to end through the public API with HOST (pinned) inputs: `e2e` ----
    copy_stream = torch.cuda.Stream()
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
    for i in range(args.steps):
        e2e_step(i)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    # clocks / throttle reasons sampled under load: from the start of the device-timed region to the end of the end-to-end one
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks ----
    if world > 1:
        tt = torch.tensor([ms, e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])
    ms_per_step = ms / args.steps
    value = args.batch * world * args.steps / (ms / 1e3)
    e2e_value = args.batch * world * args.steps / (e2e_ms / 1e3)

    if rank == 0:
        flops = synthetic.step_flops(cfg, args.batch, backward=True)
        achieved = flops / (ms_per_step / 1e3) / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": config, "mode": args.mode, "loss": loss_val,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches_per_step * args.steps,
                "roofline": dominant_kernel_roofline(args.batch, cfg, peaks),
                "attn_roofline": attention_roofline(args.batch, cfg, peaks),
                "step_roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                                  "frac": achieved / peaks["tf_sustained"],
                                  "note": f"whole step: algorithmic dense-equivalent FLOPs {flops / 1e12:.3f} TFLOP/step/GPU over the "
                                          f"CUDA-event step time, vs {peaks['source']} sustained bf16 peak"}}
        if not args.no_cpu_baseline and world >= 1:
            try:
                v, sec, threads = cpu_reference_run(cfg, 3, 1, args.cpu_sample)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                        "sample": f"{args.cpu_sample}-sample GQA-shaped batch (the GPU step's batch), mean of 3 fwd+bwd encoder "
                                                  f"steps after one warm-up (oracle port, torch CPU fp32, {sec:.2f} s per step; no optimizer)"}
            except Exception as e:  # the GPU number must not die with the CPU leg
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing the NCCL communicator down: destroying it while CUDA graphs that captured its collectives are
        # alive can block for minutes.  The ranks agree that everyone is done, then exit.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
