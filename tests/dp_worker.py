"""Worker of tests/test_gpu_dp.py (one process per GPU, launched by torch.distributed.run): after one data-parallel step

  * the all-reduced flat gradient bucket / N equals the mean of the oracle's per-shard gradients (every rank runs the
    oracle on every shard -- SURVEY section 4, last table row; per-rank BatchNorm statistics = DDP semantics),
  * the parameters after the fused RMSprop update are bit-identical on all ranks,
  * and the same holds when the whole step (exchange and optimizer included) is replayed from ONE CUDA graph.

Prints one JSON line per rank-0 check; exits non-zero on failure."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def l2_rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    import feature_level_style_transfer_for_tsc_b200 as T
    from feature_level_style_transfer_for_tsc_b200.train_step import StyleTransferModelSet, Trainer
    from oracle import os_cnn as O
    from oracle import step as OS

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T._lib.load()
    Ct, Lt, Kt, Cs, Ls, Ks, Bs = 3, 96, 3, 2, 80, 4, 8           # per-rank shard of Bs series per domain
    style_w = 50.0
    xt, yt = O.synthetic_batch(Bs * world, Ct, Lt, Kt, 0)
    xs, ys = O.synthetic_batch(Bs * world, Cs, Ls, Ks, 1)
    shard = lambda t, r: t[r * Bs:(r + 1) * Bs]
    ok = True
    for engine, dtype, tol in (("simt", torch.float32, 2e-4), ("tcgen05", torch.float64, 2.5e-1)):
        T.set_engine(engine)
        # the oracle's per-shard gradients (identical parameters on every shard: one seed)
        ref = {}
        for r in range(world):
            oms = OS.ModelSet(Ct, Lt, Kt, Cs, Ls, Ks, seed=0)
            if dtype == torch.float64:
                for name in ("fe_t", "cl_t", "fe_s", "du", "cl_s"):
                    setattr(oms, name, O.clone_state(getattr(oms, name), dtype))
            oms.set_requires_grad()
            out = OS.step_forward(oms, shard(xt, r).to(dtype), shard(yt, r), shard(xs, r).to(dtype), shard(ys, r), style_w, True)
            out["loss"].backward()
            for gname, k, p in oms.trainable():
                if p.grad is not None:
                    ref.setdefault((gname, k), []).append(p.grad.detach().double())
        ref = {key: sum(v) / world for key, v in ref.items()}
        for use_graph in (False, True):
            torch.manual_seed(1000 + rank)                        # ranks start from DIFFERENT parameters: broadcast must fix it
            model = StyleTransferModelSet(Ct, Lt, Kt, Cs, Ls, Ks).to(dev)
            if rank == 0:
                torch.manual_seed(0)
                model0 = StyleTransferModelSet(Ct, Lt, Kt, Cs, Ls, Ks).to(dev)
                model.load_state_dict(model0.state_dict())
            tr = Trainer(model, style_weight=style_w, use_graph=use_graph)
            tr.broadcast_parameters(0)
            ins = [shard(xt, rank).to(dev), shard(yt, rank).to(dev), shard(xs, rank).to(dev), shard(ys, rank).to(dev)]
            # one whole step -- forward, backward, the gradient exchange (classifier slices all-reduced during backward, the
            # rest after it) and the fused RMSprop -- eagerly or as ONE replayed CUDA graph with the NCCL collectives inside;
            # the bucket still holds the summed gradients afterwards (the optimizer only reads it)
            print(f"[rank {rank}] {engine} graph={use_graph}: step", file=sys.stderr, flush=True)
            tr.step(*ins)
            torch.cuda.synchronize()
            print(f"[rank {rank}] {engine} graph={use_graph}: step done", file=sys.stderr, flush=True)
            worst = 0.0
            worst_key = None
            for gname in ("fe_t", "cl_t", "fe_s", "du", "cl_s"):
                for k, p in getattr(model, gname).named_parameters():
                    if k.endswith("conv1d.bias") or (gname, k) not in ref:
                        continue
                    e = l2_rel((p.grad / world).detach().cpu(), ref[(gname, k)])
                    if e > worst:
                        worst, worst_key = e, f"{gname}.{k}"
            tr.release_graph()            # a graph with NCCL inside must not outlive its use (and never the process group)
            gathered = [torch.empty_like(tr.flat.flat_p) for _ in range(world)]
            dist.all_gather(gathered, tr.flat.flat_p)
            identical = all(torch.equal(gathered[0], g) for g in gathered[1:])
            bufs_same = True
            # (BatchNorm running statistics are per rank by design: they differ between shards)
            good = worst < tol and identical
            ok = ok and good
            if rank == 0:
                print(json.dumps(dict(check="dp_equality", engine=engine, cuda_graph=use_graph, world=world,
                                      grad_l2_rel_max_vs_mean_of_oracle_shards=worst, worst_tensor=worst_key, tol=tol,
                                      params_bit_identical_across_ranks=identical, ok=good)), flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    code = 1 if int(flag) else 0
    sys.stdout.flush()
    import threading
    t = threading.Timer(30.0, lambda: os._exit(code))      # never hang the launcher in the teardown
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()
    sys.exit(code)


if __name__ == "__main__":
    main()
