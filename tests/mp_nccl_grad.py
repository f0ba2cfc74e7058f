"""Launched by tests/test_gpu_multi.py under torchrun on 2 GPUs: the gradient of rows split 2-way over time-sharded
ranks, all-reduced by the LIBRARY's communicator (per flow, side stream), against the 1-GPU gradient of the same rows;
then nma_train_step on both ranks: identical variables afterwards."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from viforssms_b200 import feed
    from viforssms_b200.config import ar_config, param_layout
    from viforssms_b200.engine import NMAEngine
    from viforssms_b200.trainer import ARStepper, glorot_blob
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    T, rows = 20000, 48
    rs = np.random.RandomState(4)
    obs = rs.normal(8.0, 3.0, size=T); obs_bin = (rs.uniform(size=T) < 0.7).astype(np.float64)
    tt = rs.randint(1, 4, size=T).astype(np.float64)
    for tc in (3, 7, 0):
        st = ARStepper(T=T, rows=rows, device=dev, rank=rank, world=world, seed=3, series=(obs, obs_bin, tt), tensor_cores=tc)
        cfg = st.cfg
        assert st.eng._lib.nma_comm_world(st.eng._h) == world
        g = torch.Generator().manual_seed(9)
        # global rows: the first `rows` start in shard 0, the others in shard 1
        half = T // 2
        idx_all = np.concatenate([rs.choice(np.arange(0, half, cfg.B), rows, replace=False),
                                  rs.choice(np.arange(half, T, cfg.B), rows, replace=False)]).astype(np.int64)
        eps_all = torch.randn(2 * rows, cfg.L0, generator=g)
        th_all = torch.stack([torch.randn(2 * rows, generator=g) * 0.5 + 4.0, torch.randn(2 * rows, generator=g) * 0.1 + 0.5,
                              torch.randn(2 * rows, generator=g) * 0.2 + 1.0], dim=1).float()
        sl = slice(rank * rows, (rank + 1) * rows)
        assert st.t0 == rank * half
        params = st.blob[:st.n_nma].clone()
        out = st.eng.elbo_fwd_bwd(params, eps_all[sl].to(dev), th_all[sl].contiguous().to(dev),
                                  torch.from_numpy(idx_all[sl] - st.t0).to(dev))
        st.eng.comm_wait()
        torch.cuda.synchronize()
        g_sharded = out["grad_params"].clone()
        if rank == 0:
            full_cfg = ar_config(p=2 * rows, T=T)
            full = NMAEngine(full_cfg, dev, tensor_cores=tc)
            full.set_series(feed.ar_base_arrays(obs, obs_bin, tt, T, 3, 50, 10))
            ref = full.elbo_fwd_bwd(params, eps_all.to(dev), th_all.to(dev), torch.from_numpy(idx_all).to(dev))
            torch.cuda.synchronize()
            # same rows, same arithmetic; only the summation order differs (atomics, all-reduce)
            layout, n = param_layout(full_cfg)
            worst = 0.0
            gn = ref["grad_params"].norm().item()
            for name, (off, shape) in layout.items():
                k = int(np.prod(shape))
                a, b = g_sharded[off:off + k].double(), ref["grad_params"][off:off + k].double()
                worst = max(worst, (a - b).norm().item() / max(b.norm().item(), 1e-6 * gn))
            print("tc=%d: 2-rank all-reduced gradient vs 1-GPU gradient of the same rows: worst per-variable rel diff %.2e"
                  % (tc, worst), flush=True)
            assert worst < 2e-5, worst
            full.close()
        # the whole iteration: ranks start from the same variables and must end with the same variables
        for _ in range(3):
            st._step(st.idx_dev)
        torch.cuda.synchronize()
        mine = st.blob.clone()
        other = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        assert torch.equal(other[0], other[1]), "ranks diverged after nma_train_step"
        # ... and the same through a captured graph
        st.capture()
        for _ in range(3):
            st.step_resident()
        torch.cuda.synchronize()
        mine = st.blob.clone()
        dist.all_gather(other, mine)
        assert torch.equal(other[0], other[1]), "ranks diverged after graph replays"
        assert torch.isfinite(mine).all()
        st.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MP_NCCL_OK", flush=True)


if __name__ == "__main__":
    main()
