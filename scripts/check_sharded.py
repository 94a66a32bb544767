"""torchrun entry: trials sharded over WORLD_SIZE GPUs must reproduce the single-GPU run of the whole batch.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from vjf_b200.distributed import ShardedVJF, shard_bounds
from vjf_b200.model import VJF

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D, d, R, H, Bg, T = 60, 3, 20, [16], 200, 12
for lik in ("poisson", "gaussian"):
    torch.manual_seed(7)  # identical parameters and data on every rank
    m = VJF.make_model(D, d, 0, R, H, lik, lr=1e-3, max_trials=Bg)
    y = (torch.poisson(torch.full((T, Bg, D), 0.8)) if lik == "poisson" else torch.randn(T, Bg, D)).cuda()
    eps = torch.randn(T, 2, Bg, d).cuda()
    lo, hi = shard_bounds(Bg, world, rank)
    sh = ShardedVJF(m)
    mu, lv, losses = sh.run(y[:, lo:hi].contiguous(), eps=eps[:, :, lo:hi].contiguous())
    # reference: the same model on one GPU with the whole batch (fused persistent kernel)
    torch.manual_seed(7)
    ref = VJF.make_model(D, d, 0, R, H, lik, lr=1e-3, max_trials=Bg)
    rmu, rlv, rlosses = ref.run(y, eps=eps)
    torch.cuda.synchronize()
    e_mu = (mu - rmu[:, lo:hi]).abs().max().item(); e_lv = (lv - rlv[:, lo:hi]).abs().max().item()
    e_loss = ((losses - rlosses).abs() / (1 + rlosses.abs())).max().item()
    e_state = (m._flat - ref._flat).abs().max().item()
    flat = m._flat.clone(); others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    lockstep = all(torch.equal(o, others[0]) for o in others)
    print(f"[rank {rank}] {lik}: |mu| {e_mu:.2e} |logvar| {e_lv:.2e} loss rel {e_loss:.2e} state {e_state:.2e} replicas identical: {lockstep}", flush=True)
    assert e_mu < 2e-4 and e_lv < 2e-4 and e_loss < 2e-4 and e_state < 2e-3 and lockstep
dist.destroy_process_group()
if rank == 0:
    print("SHARDED_OK")
