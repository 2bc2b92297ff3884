"""On-hardware numerical check of the data-parallel path (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dp_check.py

  (i)  ranks that were built from DIFFERENT seeds hold bit-identical parameters after prepare() (rank 0's, what
       DistributedDataParallel's constructor does, main_itp_ddp_tar_super_node.py:203) and stay bit-identical after k steps
       (dense flat buffer AND the row-sparse word tables with their lazy row-wise Adam);
  (ii) N ranks x B samples give the gradients of ONE rank x N*B samples (the concatenated batch) up to the order of fp32 sums,
       and the post-Adam parameters follow.
tests/test_gpu_parity.py::test_dp_two_ranks_on_hardware launches this when the box has >= 2 GPUs."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"])
    ap.add_argument("--dense-tables", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    solo = dist.new_group([0])  # every rank calls it; rank 0 uses it as a world-1 "group" for the single-process comparison
    from savqa_b200 import synthetic, train
    cfg = dict(synthetic.GQA_SHAPED, ncls=256)
    V = 4000
    keys = train.STEP_KEYS
    batches = [synthetic.make_batch(cfg, args.batch, seed=100 + r, vocab_rows=V) for r in range(world)]
    mine = {k: batches[rank][k].to(dev) for k in keys}

    model = synthetic.build_model(cfg, vocab_rows=V, seed=rank).to(dev)  # different weights per rank until prepare() broadcasts
    tr = train.EncoderTrainer(model, lr=1e-3, rowsparse=not args.dense_tables)
    tr.prepare(mine)
    result = {"world": world, "mode": args.mode, "rowsparse": not args.dense_tables}

    def all_equal(t):
        got = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(got, t.contiguous())
        return all(torch.equal(got[0], g) for g in got[1:])

    result["params_equal_after_prepare"] = all_equal(tr.flat_param) and all(all_equal(t.weight.data) for t in tr.tables)
    if args.mode == "graph":
        tr.capture(mine, warmup=1)  # one eager step
        first_grad = None
        losses = [float(tr.replay()) for _ in range(args.steps - 1)]
    else:
        losses = [float(tr.step(mine)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    result["params_equal_after_steps"] = all_equal(tr.flat_param) and all(all_equal(t.weight.data) for t in tr.tables)
    result["mirror_tracks_params"] = bool(torch.equal(tr.flat_bf16, tr.flat_param.to(torch.bfloat16)))
    grad_dp = tr.flat_grad.clone()  # all-reduced (AVG) gradients of the LAST step

    ok = True
    if rank == 0:
        ref_model = synthetic.build_model(cfg, vocab_rows=V, seed=0).to(dev)
        cat = {k: torch.cat([b[k] for b in batches], 0).to(dev) for k in keys}
        ref = train.EncoderTrainer(ref_model, lr=1e-3, rowsparse=not args.dense_tables, process_group=solo)
        assert ref.world == 1
        for _ in range(args.steps):
            ref.step(cat)
        torch.cuda.synchronize()
        assert ref.flat_param.numel() == tr.flat_param.numel()
        gerr = float((grad_dp - ref.flat_grad).norm() / ref.flat_grad.norm())
        perr = float((tr.flat_param - ref.flat_param).abs().max())
        terr = max(float((a.weight.data - b.weight.data).abs().max()) for a, b in zip(tr.tables, ref.tables))
        result.update(grad_rel_err_vs_single_rank=gerr, param_max_abs_diff=perr, table_max_abs_diff=terr, lr=1e-3, steps=args.steps,
                      losses=losses)
        # same kernels on the same operands up to the order of fp32 sums (per-rank partial sums, atomics) and the bf16 roundings
        # those reorderings flip after the first update; Adam moves a parameter by at most ~lr per step
        ok = (result["params_equal_after_prepare"] and result["params_equal_after_steps"] and result["mirror_tracks_params"]
              and gerr < 5e-2 and perr <= 2.5 * 1e-3 * args.steps and terr <= 2.5 * 1e-3 * args.steps)
        result["ok"] = ok
        print("DP_CHECK " + json.dumps(result), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
